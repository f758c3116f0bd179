"""`WanVideoGenerator` on the B200-native engine - the drop-in for the class of the same name in the reference's
`infinicube/videogen/inference.py` (constructor :42-50, checkpoint prefixes :101-128, buffer validation :130-162,
`generate` :164-236, `__call__` :238-240).  What has to stay identical for the stage-2 caller
(`guidance_buffer_generation.py:742-791`) does: argument names, defaults, return type, the order of the input checks
and the exception types.  Everything behind that surface is this package's own: the buffers go to the GPU as one
uint8 tensor each, the DiT / VAE / prompt encoder are the sm_100a kernels, nothing is offloaded.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from PIL import Image

from .pipeline import ModelConfig, WanVideoPipeline

DEFAULT_PROMPT = "The video is about a driving scene captured at daytime. The weather is clear."
DEFAULT_NEGATIVE_PROMPT = ("色调艳丽，过曝，静态，细节模糊不清，字幕，风格，作品，画作，画面，静止，整体发灰，最差质量，低质量，JPEG压缩残留，丑陋的，残缺的，"
                           "多余的手指，画得不好的手部，画得不好的脸部，畸形的，毁容的，形态畸形的肢体，手指融合，静止不动的画面，杂乱的背景，三条腿，背景人很多，倒着走")
_BASE_FILES = ("diffusion_pytorch_model*.safetensors", "models_t5_umt5-xxl-enc-bf16.pth", "Wan2.1_VAE.pth")


def load_state_dict(path: str, device="cpu") -> Dict[str, torch.Tensor]:
    """Stand-in for diffsynth.load_state_dict: a safetensors file or a torch checkpoint as a flat tensor dict."""
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device=str(device))
    sd = torch.load(path, map_location=device, weights_only=True)
    return sd.get("state_dict", sd) if isinstance(sd, dict) else sd


def save_video(frames, save_path: str, fps: int = 10, quality: int = 8) -> None:
    """Stand-in for diffsynth.save_video.  OpenCV's mp4 writer (imageio / ffmpeg are not in this image); `quality`
    is accepted for signature compatibility, the mp4v encoder has no such knob."""
    import cv2
    width, height = frames[0].size
    writer = cv2.VideoWriter(save_path, cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (width, height))
    if not writer.isOpened():
        raise RuntimeError(f"cannot open video writer for {save_path}")
    for frame in frames:
        writer.write(cv2.cvtColor(np.asarray(frame), cv2.COLOR_RGB2BGR))
    writer.release()


def split_checkpoint(state_dict: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """The reference's checkpoint convention (inference.py:107-125): one file, `buffer_embedder.*` and `dit.*` keys;
    returns the two sub-dicts with the prefixes stripped (anything else in the file is ignored, as there)."""
    parts = {"buffer_embedder.": {}, "dit.": {}}
    for key, value in state_dict.items():
        for prefix, bucket in parts.items():
            if key.startswith(prefix):
                bucket[key[len(prefix):]] = value
    return parts["buffer_embedder."], parts["dit."]


def check_buffer(buffer_array, name: str = "buffer_array") -> None:
    """Input contract of the reference's `_ndarray_to_pil_list` (:140-153), same order and exception types:
    ndarray (TypeError) -> (N, H, W, 3) (ValueError) -> uint8 (TypeError)."""
    if not isinstance(buffer_array, np.ndarray):
        raise TypeError(f"{name} must be numpy.ndarray, got {type(buffer_array)}")
    if buffer_array.ndim != 4 or buffer_array.shape[-1] != 3:
        raise ValueError(f"{name} shape must be (N, H, W, 3), got {buffer_array.shape}")
    if buffer_array.dtype != np.uint8:
        raise TypeError(f"{name} dtype must be uint8, got {buffer_array.dtype}")


def make_rank_generator(rank: int, world_size: int, **kwargs) -> "WanVideoGenerator":
    """Per-rank factory of the single-process multi-GPU entry (multiproc.RankPool): rank r drives cuda:r."""
    return WanVideoGenerator(device=f"cuda:{rank}", world_size=world_size, rank=rank, _in_pool=True, **kwargs)


class WanVideoGenerator:
    """Video generation from a semantic and a coordinate guidance buffer (Wan2.1 T2V + buffer embedder).

    Reference arguments: `checkpoint_path` (trained `buffer_embedder.*` + `dit.*` weights), `device` ("cuda:0"),
    `torch_dtype` (bfloat16), `buffer_channels` (16), `enable_vram_management` (accepted; weights always stay
    resident here), `use_wan_1pt3b`.
    Extra keyword arguments, all optional and off by default: `synthetic_weights` (random-init weights when no Wan
    checkpoint is on disk), `world_size` / `rank`, `cfg_parallel` (prompt / negative-prompt forwards on two rank
    groups; default on for an even world).  `world_size > 1` works in two ways: launched one process per GPU
    (`torchrun`; pass that process's `rank`), or from ONE plain process exactly as the stage-2 script constructs it
    (`device="cuda:0"`, no rank): the generator then spawns and owns the other ranks (`videogen/multiproc.py`).
    """

    def __init__(
        self,
        checkpoint_path: str,
        device: str = "cuda:0",
        torch_dtype: torch.dtype = torch.bfloat16,
        buffer_channels: int = 16,
        enable_vram_management: bool = True,
        use_wan_1pt3b: bool = False,
        synthetic_weights: Optional[bool] = None,
        world_size: int = 1,
        rank: int = 0,
        cfg_parallel: Optional[bool] = None,
        _in_pool: bool = False,
    ):
        self.checkpoint_path, self.device = checkpoint_path, device
        self.torch_dtype, self.buffer_channels = torch_dtype, buffer_channels
        self._pool = None
        if world_size > 1 and not _in_pool:
            from .multiproc import RankPool, launched_per_rank
            if not launched_per_rank():
                # ONE process, as the stage-2 caller is (guidance_buffer_generation.py:755-768): this object becomes
                # rank 0 and owns ranks 1..world_size-1 as worker processes, one per GPU
                print(f"[WanVideoGenerator] single-process entry: spawning {world_size - 1} worker rank(s)")
                self._pool = RankPool(world_size, make_rank_generator, dict(
                    checkpoint_path=checkpoint_path, torch_dtype=torch_dtype, buffer_channels=buffer_channels,
                    enable_vram_management=enable_vram_management, use_wan_1pt3b=use_wan_1pt3b,
                    synthetic_weights=synthetic_weights, cfg_parallel=cfg_parallel))
                self.pipe = self._pool.obj.pipe
                return
        model_id = f"Wan-AI/Wan2.1-T2V-{'1.3B' if use_wan_1pt3b else '14B'}"
        print(f"[WanVideoGenerator] base model {model_id}, device {device}, {world_size} rank(s)")
        self.pipe = WanVideoPipeline.from_pretrained(
            torch_dtype=torch_dtype, device=device,
            model_configs=[ModelConfig(model_id=model_id, origin_file_pattern=pat, skip_download=True) for pat in _BASE_FILES],
            synthetic_weights=synthetic_weights, world_size=world_size, rank=rank, cfg_parallel=cfg_parallel)
        self.pipe.initialize_buffer_embedder(buffer_channels=buffer_channels, zero_init=True)
        self._load_checkpoint()
        if enable_vram_management:
            self.pipe.enable_vram_management()
        print("[WanVideoGenerator] ready")

    def _load_checkpoint(self) -> None:
        if self.pipe.synthetic and not os.path.exists(self.checkpoint_path):
            print(f"[WanVideoGenerator] {self.checkpoint_path} not found: keeping the synthetic weights")
            return
        embedder, dit = split_checkpoint(load_state_dict(self.checkpoint_path))
        if self.pipe.buffer_embedder is not None:
            if embedder:
                self.pipe.buffer_embedder.load_state_dict(embedder)
            else:
                print("[WanVideoGenerator] warning: the checkpoint holds no buffer_embedder.* weights")
        if dit:
            self.pipe.dit.load_state_dict(dit, strict=False)
        print(f"[WanVideoGenerator] checkpoint {self.checkpoint_path}: {len(embedder)} buffer-embedder and {len(dit)} DiT tensors")

    # kept for callers / tests written against the reference class
    _validate_buffer = staticmethod(check_buffer)

    def _ndarray_to_pil_list(self, buffer_array: np.ndarray) -> List[Image.Image]:
        check_buffer(buffer_array)
        return [Image.fromarray(frame, mode="RGB") for frame in buffer_array]

    def generate(
        self,
        semantic_buffer: np.ndarray,
        coordinate_buffer: np.ndarray,
        prompt: str = DEFAULT_PROMPT,
        negative_prompt=DEFAULT_NEGATIVE_PROMPT,
        seed: int = 0,
        tiled: bool = True,
        output_path: Optional[str] = None,
        fps: int = 10,
        quality: int = 8,
    ) -> List[Image.Image]:
        """Two uint8 (N, H, W, 3) buffers -> N PIL frames (optionally also written to `output_path` as mp4)."""
        if semantic_buffer.shape != coordinate_buffer.shape:     # the reference compares shapes first (:194-198)
            raise ValueError(f"semantic_buffer and coordinate_buffer must have the same shape, "
                             f"got {semantic_buffer.shape} and {coordinate_buffer.shape}")
        check_buffer(semantic_buffer)
        check_buffer(coordinate_buffer)
        num_frames, height, width, _ = semantic_buffer.shape
        if self._pool is not None:   # every rank runs the same call on its shard; rank 0's frames are returned
            return self._pool.call("generate", semantic_buffer, coordinate_buffer, prompt=prompt,
                                   negative_prompt=negative_prompt, seed=seed, tiled=tiled, output_path=output_path,
                                   fps=fps, quality=quality)
        print(f"[WanVideoGenerator] generate: {num_frames} frames {width}x{height}, seed {seed}, tiled {tiled}, prompt {prompt!r}")
        video = self.pipe(prompt=prompt, negative_prompt=negative_prompt, semantic_buffer_video=semantic_buffer,
                          coordinate_buffer_video=coordinate_buffer, height=height, width=width, num_frames=num_frames,
                          seed=seed, tiled=tiled)
        if output_path is not None and getattr(self.pipe, "rank", 0) == 0:
            save_video(video, output_path, fps=fps, quality=quality)
            print(f"[WanVideoGenerator] wrote {output_path}")
        return video

    def close(self) -> None:
        """Stops the worker ranks of the single-process multi-GPU entry (no-op otherwise)."""
        if self._pool is not None:
            self._pool.close()
            self._pool = None

    def generate_device(self, semantic_buffer: torch.Tensor, coordinate_buffer: torch.Tensor, prompt: str = DEFAULT_PROMPT,
                        negative_prompt: str = "", seed: int = 0, tiled: bool = True, output_type: str = "tensor"):
        """GPU-resident hand-off: uint8 CUDA tensors (N, H, W, 3) straight from the rasteriser go
        to the VAE encoder without the GPU -> numpy -> PIL -> GPU round trip of the reference
        (guidance_buffer_generation.py:667-745, videogen/inference.py:130-162); returns uint8 frames on the device."""
        for name, b in (("semantic_buffer", semantic_buffer), ("coordinate_buffer", coordinate_buffer)):
            if not isinstance(b, torch.Tensor):
                raise TypeError(f"{name} must be a torch.Tensor, got {type(b)}")
            if b.dtype != torch.uint8:
                raise TypeError(f"{name} dtype must be uint8, got {b.dtype}")
            if b.ndim != 4 or b.shape[-1] != 3:
                raise ValueError(f"{name} shape must be (N, H, W, 3), got {tuple(b.shape)}")
        if semantic_buffer.shape != coordinate_buffer.shape:
            raise ValueError("semantic_buffer and coordinate_buffer must have the same shape")
        if self._pool is not None:
            return self._pool.call("generate_device", semantic_buffer, coordinate_buffer, prompt=prompt,
                                   negative_prompt=negative_prompt, seed=seed, tiled=tiled, output_type=output_type)
        n, h, w, _ = semantic_buffer.shape
        return self.pipe(prompt=prompt, negative_prompt=negative_prompt, semantic_buffer_video=semantic_buffer,
                         coordinate_buffer_video=coordinate_buffer, height=h, width=w, num_frames=n, seed=seed,
                         tiled=tiled, output_type=output_type)

    def __call__(self, *args, **kwargs):
        return self.generate(*args, **kwargs)

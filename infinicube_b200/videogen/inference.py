"""Video Generation Inference API — drop-in for `infinicube/videogen/inference.py` of the reference
(same class, constructor and `generate` signature, same validation and error types, lines 42-50,
130-162, 164-240), running on the B200-native engine instead of diffsynth."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
from PIL import Image

from .pipeline import ModelConfig, WanVideoPipeline


def load_state_dict(path: str, device="cpu"):
    """diffsynth.load_state_dict: safetensors or torch checkpoint -> dict of tensors."""
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device=str(device))
    sd = torch.load(path, map_location=device, weights_only=True)
    return sd.get("state_dict", sd) if isinstance(sd, dict) else sd


def save_video(frames, save_path: str, fps: int = 10, quality: int = 8):
    """diffsynth.save_video: H.264 mp4 through OpenCV's writer (imageio/ffmpeg are not in this image)."""
    import cv2
    w, h = frames[0].size
    wr = cv2.VideoWriter(save_path, cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (w, h))
    if not wr.isOpened():
        raise RuntimeError(f"cannot open video writer for {save_path}")
    for f in frames:
        wr.write(cv2.cvtColor(np.asarray(f), cv2.COLOR_RGB2BGR))
    wr.release()


class WanVideoGenerator:
    """
    Wan Video Generator - Video generation based on semantic and coordinate buffers

    Args:
        checkpoint_path: Path to trained model checkpoint (contains buffer_embedder and dit weights)
        device: Device to run on, default "cuda:0"
        torch_dtype: Torch data type, default torch.bfloat16
        buffer_channels: Number of channels for buffer embedder, default 16
        enable_vram_management: Whether to enable VRAM management, default True
    Extra optional keyword (same default behaviour as the reference when omitted):
        synthetic_weights: run with random-init weights when no Wan checkpoint exists on disk
        world_size, rank: one process per GPU under torch.distributed (the loop shards over the ranks)
        cfg_parallel: run the prompt / negative-prompt forwards on two rank groups (default: on for even worlds)
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
    ):
        self.checkpoint_path = checkpoint_path
        self.device = device
        self.torch_dtype = torch_dtype
        self.buffer_channels = buffer_channels

        model_id = "Wan-AI/Wan2.1-T2V-1.3B" if use_wan_1pt3b else "Wan-AI/Wan2.1-T2V-14B"
        print(f"Loading {model_id.split('/')[-1].replace('Wan2.1-', 'Wan2.1-')} base model...")
        self.pipe = WanVideoPipeline.from_pretrained(
            torch_dtype=torch_dtype,
            device=device,
            model_configs=[
                ModelConfig(model_id=model_id, origin_file_pattern="diffusion_pytorch_model*.safetensors", skip_download=True),
                ModelConfig(model_id=model_id, origin_file_pattern="models_t5_umt5-xxl-enc-bf16.pth", skip_download=True),
                ModelConfig(model_id=model_id, origin_file_pattern="Wan2.1_VAE.pth", skip_download=True),
            ],
            synthetic_weights=synthetic_weights,
            world_size=world_size,
            rank=rank,
            cfg_parallel=cfg_parallel,
        )

        print(f"Initializing buffer embedder (channels={buffer_channels})...")
        self.pipe.initialize_buffer_embedder(buffer_channels=buffer_channels, zero_init=True)

        print(f"Loading checkpoint: {checkpoint_path}")
        self._load_checkpoint()

        if enable_vram_management:
            print("Enabling VRAM management...")
            self.pipe.enable_vram_management()

        print("✓ WanVideoGenerator initialization complete")

    def _load_checkpoint(self):
        """Load trained checkpoint (prefix convention of the reference, inference.py:101-128)"""
        import os
        if self.pipe.synthetic and not os.path.exists(self.checkpoint_path):
            print("  ⚠ Warning: checkpoint not found; keeping synthetic weights")
            return
        state_dict = load_state_dict(self.checkpoint_path)

        if self.pipe.buffer_embedder is not None:
            buffer_embedder_state = {
                k.replace("buffer_embedder.", ""): v for k, v in state_dict.items() if k.startswith("buffer_embedder.")
            }
            if buffer_embedder_state:
                self.pipe.buffer_embedder.load_state_dict(buffer_embedder_state)
                print(f"  ✓ Buffer embedder weights loaded, {len(buffer_embedder_state)} parameters")
            else:
                print("  ⚠ Warning: buffer_embedder weights not found in checkpoint")

        dit_state = {k.replace("dit.", ""): v for k, v in state_dict.items() if k.startswith("dit.")}
        if dit_state:
            self.pipe.dit.load_state_dict(dit_state, strict=False)
            print(f"  ✓ DiT weights loaded, {len(dit_state)} parameters")

    def _ndarray_to_pil_list(self, buffer_array: np.ndarray) -> List[Image.Image]:
        """Validate a (N, H, W, 3) uint8 buffer and split it into PIL frames (inference.py:130-162)."""
        if not isinstance(buffer_array, np.ndarray):
            raise TypeError(f"buffer_array must be numpy.ndarray, got {type(buffer_array)}")
        if buffer_array.ndim != 4 or buffer_array.shape[-1] != 3:
            raise ValueError(f"buffer_array shape must be (N, H, W, 3), got {buffer_array.shape}")
        if buffer_array.dtype != np.uint8:
            raise TypeError(f"buffer_array dtype must be uint8, got {buffer_array.dtype}")
        return [Image.fromarray(buffer_array[i], mode="RGB") for i in range(buffer_array.shape[0])]

    @staticmethod
    def _validate_buffer(buffer_array) -> None:
        if not isinstance(buffer_array, np.ndarray):
            raise TypeError(f"buffer_array must be numpy.ndarray, got {type(buffer_array)}")
        if buffer_array.ndim != 4 or buffer_array.shape[-1] != 3:
            raise ValueError(f"buffer_array shape must be (N, H, W, 3), got {buffer_array.shape}")
        if buffer_array.dtype != np.uint8:
            raise TypeError(f"buffer_array dtype must be uint8, got {buffer_array.dtype}")

    def generate(
        self,
        semantic_buffer: np.ndarray,
        coordinate_buffer: np.ndarray,
        prompt: str = "The video is about a driving scene captured at daytime. The weather is clear.",
        negative_prompt="色调艳丽，过曝，静态，细节模糊不清，字幕，风格，作品，画作，画面，静止，整体发灰，最差质量，低质量，JPEG压缩残留，丑陋的，残缺的，多余的手指，画得不好的手部，画得不好的脸部，畸形的，毁容的，形态畸形的肢体，手指融合，静止不动的画面，杂乱的背景，三条腿，背景人很多，倒着走",
        seed: int = 0,
        tiled: bool = True,
        output_path: Optional[str] = None,
        fps: int = 10,
        quality: int = 8,
    ) -> List[Image.Image]:
        if semantic_buffer.shape != coordinate_buffer.shape:
            raise ValueError(
                f"semantic_buffer and coordinate_buffer must have the same shape, "
                f"got {semantic_buffer.shape} and {coordinate_buffer.shape}"
            )
        # same checks, same order and exception types as _ndarray_to_pil_list in the reference; the frames
        # themselves go to the GPU as one uint8 tensor instead of 2 x N PIL images
        self._validate_buffer(semantic_buffer)
        self._validate_buffer(coordinate_buffer)
        num_frames, height, width, channels = semantic_buffer.shape

        print("\nStarting video generation...")
        print(f"  - Prompt: {prompt}")
        print(f"  - Frames: {num_frames}")
        print(f"  - Resolution: {height}x{width}")
        print(f"  - Seed: {seed}")
        print(f"  - Tiled: {tiled}")

        print("Executing video generation...")
        video = self.pipe(
            prompt=prompt,
            negative_prompt=negative_prompt,
            semantic_buffer_video=semantic_buffer,
            coordinate_buffer_video=coordinate_buffer,
            height=height,
            width=width,
            num_frames=num_frames,
            seed=seed,
            tiled=tiled,
        )

        if output_path is not None:
            print(f"Saving video to: {output_path}")
            save_video(video, output_path, fps=fps, quality=quality)
            print("✓ Video saved")

        print(f"✓ Video generation complete ({len(video)} frames)")
        return video

    def generate_device(self, semantic_buffer: torch.Tensor, coordinate_buffer: torch.Tensor, prompt: str = "The video is about a driving scene captured at daytime. The weather is clear.",
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
        n, h, w, _ = semantic_buffer.shape
        return self.pipe(prompt=prompt, negative_prompt=negative_prompt, semantic_buffer_video=semantic_buffer,
                         coordinate_buffer_video=coordinate_buffer, height=h, width=w, num_frames=n, seed=seed,
                         tiled=tiled, output_type=output_type)

    def __call__(self, *args, **kwargs):
        """Make instance callable like a function"""
        return self.generate(*args, **kwargs)

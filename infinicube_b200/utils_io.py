"""On-disk formats of stage 2: webdataset-style tars (`{clip}.{name}` members), 16-bit PNG depth /
instance images, .npy poses and intrinsics, mp4 previews.  Mirrors infinicube/utils/wds_utils.py:300-313,
infinicube/data_process/waymo_utils.py:32-45 and infinicube/utils/fileio_utils.py:58-138 with the tools this image
has (stdlib tarfile, OpenCV) instead of webdataset / imageio."""
from __future__ import annotations

import io
import json
import tarfile
from pathlib import Path
from typing import Dict, List, Union

import numpy as np


def encode_png(image: np.ndarray) -> bytes:
    """PNG bytes; handles uint16 (depth x 100, instance ids) like imageencoder_imageio_png."""
    import cv2
    img = np.ascontiguousarray(image)
    if img.ndim == 3 and img.shape[2] == 3:
        img = cv2.cvtColor(img, cv2.COLOR_RGB2BGR)
    ok, buf = cv2.imencode(".png", img)
    if not ok:
        raise RuntimeError("PNG encoding failed")
    return buf.tobytes()


def decode_png(data: bytes) -> np.ndarray:
    import cv2
    img = cv2.imdecode(np.frombuffer(data, dtype=np.uint8), cv2.IMREAD_UNCHANGED)
    if img is not None and img.ndim == 3 and img.shape[2] == 3:
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    return img


def _encode_member(name: str, value) -> bytes:
    if isinstance(value, (bytes, bytearray)):
        return bytes(value)
    if name.endswith(".npy"):
        with io.BytesIO() as b:
            np.save(b, np.asarray(value))
            return b.getvalue()
    if name.endswith(".json") or isinstance(value, (dict, list)):
        return json.dumps(value).encode("utf-8")
    if isinstance(value, str):
        return value.encode("utf-8")
    raise TypeError(f"cannot encode tar member {name} of type {type(value)}")


def write_to_tar(sample: Dict, output_file: Union[str, Path], __key__: str = None) -> None:
    """One webdataset sample per tar: every entry `name -> value` becomes the member `{__key__}.{name}`."""
    sample = dict(sample)
    key = __key__ if __key__ is not None else sample.pop("__key__", "sample")
    sample.pop("__key__", None)
    output_file = Path(output_file)
    output_file.parent.mkdir(parents=True, exist_ok=True)
    with tarfile.open(output_file, "w") as tar:
        for name, value in sample.items():
            data = _encode_member(name, value)
            info = tarfile.TarInfo(f"{key}.{name}")
            info.size = len(data)
            tar.addfile(info, io.BytesIO(data))
    print(f"Saved {output_file}")


def get_sample(tar_file: Union[str, Path]) -> Dict:
    """Inverse of write_to_tar: {name: decoded value} plus `__key__` (wds_utils.get_sample)."""
    out: Dict = {}
    with tarfile.open(tar_file, "r") as tar:
        for m in tar.getmembers():
            data = tar.extractfile(m).read()
            key, name = m.name.split(".", 1)
            out["__key__"] = key
            if name.endswith(".png"):
                out[name] = decode_png(data)
            elif name.endswith(".npy"):
                out[name] = np.load(io.BytesIO(data), allow_pickle=False)
            elif name.endswith(".json"):
                out[name] = json.loads(data.decode("utf-8"))
            else:
                out[name] = data
    return out


def write_video_file(frames: Union[np.ndarray, List, Dict], output_file: Union[str, Path], fps: int = 30) -> None:
    """mp4 preview (OpenCV mp4v; the reference uses x264 `-preset veryslow` through imageio/ffmpeg)."""
    import cv2
    output_file = Path(output_file).with_suffix(".mp4")
    output_file.parent.mkdir(parents=True, exist_ok=True)
    if isinstance(frames, dict):
        frames = [frames[k] for k in sorted(k for k in frames if k != "__key__")]
    frames = [np.asarray(f) for f in frames]
    assert len(frames) > 0
    h, w = frames[0].shape[:2]
    wr = cv2.VideoWriter(str(output_file), cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (w, h))
    if not wr.isOpened():
        raise RuntimeError(f"cannot open video writer for {output_file}")
    for f in frames:
        wr.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    wr.release()


def vis_depth(depth: np.ndarray, valid_farthest: float = 300.0) -> np.ndarray:
    """Depth preview (H, W, 3) uint8: adaptive 0.5 / 99.5 percentile normalisation like depth_utils.vis_depth:20-69,
    rendered with OpenCV's magma map reversed (matplotlib's 'magma_r' is not available here)."""
    import cv2
    d = np.nan_to_num(np.asarray(depth, dtype=np.float32))
    valid = d[d < valid_farthest]
    hi = np.percentile(valid, 99.5) if valid.size else 1.0
    lo = np.percentile(d, 0.5)
    lo = lo if lo < hi else 0.0
    n = np.clip((d - lo) / max(hi - lo, 1e-12), 0.0, 1.0)
    img = cv2.applyColorMap(((1.0 - n) * 255).astype(np.uint8), cv2.COLORMAP_MAGMA)
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


def read_video_file(video_file: Union[str, Path]) -> np.ndarray:
    """All frames of an mp4 as uint8 [N, H, W, 3] RGB (fileio_utils.read_video_file + VideoReader.get_batch of the
    reference, which use decord; OpenCV here)."""
    import cv2
    cap = cv2.VideoCapture(str(video_file))
    if not cap.isOpened():
        raise FileNotFoundError(f"cannot open video {video_file}")
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(cv2.cvtColor(f, cv2.COLOR_BGR2RGB))
    cap.release()
    if not frames:
        raise ValueError(f"{video_file} holds no frames")
    return np.stack(frames)

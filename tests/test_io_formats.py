"""Stage-2 file formats: tar layout, 16-bit PNG depth x 100 / instance ids, npy poses."""
import numpy as np
import pytest
import torch

from infinicube_b200.utils_io import decode_png, encode_png, get_sample, vis_depth, write_to_tar, write_video_file


def test_png16_and_tar_roundtrip(tmp_path):
    depth = (np.random.RandomState(0).rand(30, 52).astype(np.float32) * 80)
    d16 = (depth * 100).astype(np.uint16)  # guidance_buffer_generation.py:670
    assert np.array_equal(decode_png(encode_png(d16)), d16)
    rgb = np.random.RandomState(1).randint(0, 255, (8, 9, 3), dtype=np.uint8)
    assert np.array_equal(decode_png(encode_png(rgb)), rgb)
    sample = {"000000.voxel_depth_100.front.png": encode_png(d16), "000000.pose.front.npy": np.eye(4, dtype=np.float32),
              "000000.dynamic_object_info.json": {"car": {"object_lwh": [4.0, 2.0, 1.5]}}}
    write_to_tar(sample, tmp_path / "x" / "a.tar", __key__="clip123")
    import tarfile
    names = tarfile.open(tmp_path / "x" / "a.tar").getnames()
    assert names[0] == "clip123.000000.voxel_depth_100.front.png"
    back = get_sample(tmp_path / "x" / "a.tar")
    assert back["__key__"] == "clip123"
    assert np.array_equal(back["000000.voxel_depth_100.front.png"], d16)
    assert np.array_equal(back["000000.pose.front.npy"], np.eye(4, dtype=np.float32))
    assert back["000000.dynamic_object_info.json"]["car"]["object_lwh"] == [4.0, 2.0, 1.5]


def test_video_and_depth_preview(tmp_path):
    frames = [np.full((32, 48, 3), i * 20, np.uint8) for i in range(5)]
    write_video_file(frames, tmp_path / "v.mp4", fps=10)
    assert (tmp_path / "v.mp4").stat().st_size > 0
    vis = vis_depth(np.linspace(0, 50, 32 * 48, dtype=np.float32).reshape(32, 48))
    assert vis.shape == (32, 48, 3) and vis.dtype == np.uint8


@pytest.mark.gpu
def test_generate_guidance_buffer_and_save(tmp_path):
    from infinicube_b200.inference.guidance_buffer_generation import generate_guidance_buffer_and_save
    from infinicube_b200.raster import PinholeCamera, synthetic as syn
    dev = torch.device("cuda:0")
    pts, sem, inst, _ = syn.synthetic_scene(32)
    cam = PinholeCamera.from_numpy(np.array([100.0, 90.0, 48.0, 32.0, 96, 64]), device=dev)
    poses = torch.from_numpy(syn.synthetic_poses(32, n=5)).to(dev)
    torch.manual_seed(0)
    d, s, i = generate_guidance_buffer_and_save("clipA", tmp_path, "64p", cam, poses, torch.from_numpy(pts).to(dev),
                                                torch.from_numpy(sem).to(dev).long(), {}, {}, "a street", True, "none.safetensors",
                                                True, rng=np.random.RandomState(0))
    dep = get_sample(tmp_path / "voxel_depth_100_64p_front.tar")
    assert np.array_equal(dep["000003.voxel_depth_100.front.png"], (d[3].cpu().numpy() * 100).astype(np.uint16))
    ins = get_sample(tmp_path / "instance_buffer_64p_front.tar")
    assert np.array_equal(ins["000000.instance_buffer.front.png"], i[0].cpu().numpy().astype(np.uint16))
    assert np.array_equal(get_sample(tmp_path / "pose.tar")["000004.pose.front.npy"], poses[4].cpu().numpy())
    assert np.array_equal(get_sample(tmp_path / "intrinsic.tar")["intrinsic.front.npy"], cam.intrinsics)
    for f in ("semantic_buffer_video_64p_front.mp4", "depth_vis_video_64p_front.mp4", "coordinate_buffer_video_64p_front.mp4"):
        assert (tmp_path / f).stat().st_size > 0
    assert not (s == 1).any()  # cad_model_for_static_object=True drops the scene's car voxels (no boxes given)

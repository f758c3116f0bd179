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


def test_voxel_world_file_selection_and_roundtrip(tmp_path):
    """load_voxel's file choice (guidance_buffer_generation.py:446-456) and the neutral .npz wire format."""
    from infinicube_b200.inference.guidance_buffer_generation import read_voxel_file, save_voxel_npz, select_voxel_file
    rng = np.random.default_rng(0)
    ijk = rng.integers(-50, 50, (1000, 3)).astype(np.int32)
    sem = rng.integers(0, 23, 1000)
    clip = "clip0"
    for step in (3, 20, 100):
        save_voxel_npz(tmp_path / clip / f"{step}.npz", ijk + step, sem, 0.2, 0.1)
    torch.save({"points": torch.zeros(4, 3), "semantics": torch.zeros(4)}, tmp_path / clip / "100.pt")
    torch.save({"points": {"ijk": torch.from_numpy(ijk), "voxel_size": 0.2, "origin": 0.1},
                "semantics": torch.from_numpy(sem)}, tmp_path / clip / "7.pt")
    assert select_voxel_file(tmp_path, clip).name == "100.npz"          # numeric, not lexicographic, maximum
    assert select_voxel_file(tmp_path, clip, 20).name == "20.npz"
    assert select_voxel_file(tmp_path, clip, 7).name == "7.pt"
    with pytest.raises(FileNotFoundError):
        select_voxel_file(tmp_path, clip, 5)
    with pytest.raises(FileNotFoundError):
        select_voxel_file(tmp_path, "missing")
    pts, s = read_voxel_file(tmp_path / clip / "20.npz")
    assert pts.dtype == torch.float32 and s.dtype == torch.int64
    # round((p - origin) / vs) recovers ijk (points_to_fvdb's rule, utils/fvdb_utils.py:156)
    assert np.array_equal(torch.round((pts - 0.1) / 0.2).int().numpy(), ijk + 20)
    assert np.array_equal(s.numpy(), sem)
    pts7, s7 = read_voxel_file(tmp_path / clip / "7.pt")
    assert np.array_equal(torch.round((pts7 - 0.1) / 0.2).int().numpy(), ijk) and np.array_equal(s7.numpy(), sem)
    with pytest.raises(ValueError):
        save_voxel_npz(tmp_path / "bad.npz", ijk, sem[:10])


def test_stage3_reader_consumes_stage2_folder(tmp_path):
    """The folder layout our stage 2 writes (guidance_buffer_generation.py:645-728) read back the way stage 3 does
    (scene_gaussian_generation.py:258-372): depth / 100, instance >= 10000 -> dynamic, key-frame priority, GSM masks."""
    import json
    from types import SimpleNamespace
    from infinicube_b200.inference.scene_gaussian_generation import get_data_dict_from_folder
    rs = np.random.RandomState(0)
    n, h, w = 9, 32, 48
    depth = rs.rand(n, h, w).astype(np.float32) * 60
    depth[:, :6] = 0                                   # sky rows
    inst = np.zeros((n, h, w), np.uint16)
    inst[:, 20:26, 10:20] = 10003                      # a dynamic object
    inst[:, 10:14, 30:40] = 17                         # a static one
    poses = np.stack([np.eye(4, dtype=np.float32) for _ in range(n)])
    poses[:, 0, 3] = np.arange(n)
    intr = np.array([50.0, 45.0, 24.0, 16.0, w, h], dtype=np.float32)
    clip = "clipZ"
    write_to_tar({f"{i:06d}.voxel_depth_100.front.png": encode_png((depth[i] * 100).astype(np.uint16)) for i in range(n)},
                 tmp_path / "voxel_depth_100_480p_front.tar", __key__=clip)
    write_to_tar({f"{i:06d}.instance_buffer.front.png": encode_png(inst[i]) for i in range(n)},
                 tmp_path / "instance_buffer_480p_front.tar", __key__=clip)
    write_to_tar({f"{i:06d}.pose.front.npy": poses[i] for i in range(n)}, tmp_path / "pose.tar", __key__=clip)
    write_to_tar({"intrinsic.front.npy": intr}, tmp_path / "intrinsic.tar", __key__=clip)
    write_to_tar({f"{i:06d}.dynamic_object_info.json": {"10003": {"object_lwh": [4.0, 2.0, 1.5], "frame": i}} for i in range(n)},
                 tmp_path / "dynamic_object_info.tar", __key__=clip)
    write_video_file([np.full((h, w, 3), 25 * i, np.uint8) for i in range(n)], tmp_path / "video_480p_front.mp4", fps=10)
    args = SimpleNamespace(data_folder=str(tmp_path), start_frame_index=1, active_frame_proportion=0.7, use_frame_interval=2,
                           enable_pixel_branch_last_n_frame=1)
    d = get_data_dict_from_folder(args)
    assert d["key_frame_indices"] == [1, 3, 5]                      # range(1, min(1 + int(0.7 * 9), 9), 2)
    k = d["key_frame_indices"]
    assert torch.equal(d["poses"], torch.from_numpy(poses[k])) and tuple(d["intrinsics"].shape) == (3, 6)
    want_depth = torch.from_numpy((depth[k] * 100).astype(np.uint16) / 100.0).float()
    assert torch.equal(d["depth_buffers"], want_depth)
    assert torch.equal(d["dynamic_masks"], torch.from_numpy(inst[k].astype(np.int32) >= 10000))
    assert torch.equal(d["non_dynamic_masks"], ~d["dynamic_masks"])
    assert [x["10003"]["frame"] for x in d["dynamic_object_infos"]] == k
    assert tuple(d["video_array"].shape) == (3, h, w, 3) and 0.0 <= float(d["video_array"].min()) and float(d["video_array"].max()) <= 1.0
    assert abs(float(d["video_array"][1].mean()) * 255 - 75) < 6     # frame 3 was grey level 75 (lossy codec)
    m = d["gsm_images_input_mask"]
    assert tuple(m.shape) == (3, h, w, 4)
    assert torch.equal(m[..., 3].bool(), want_depth != 0)           # foreground from the depth grid
    assert torch.equal(m[:-1, ..., 0], m[:-1, ..., 3])              # pixel branch disabled before the last frame
    assert bool(m[-1, ..., 0].all()) and bool(m[..., 1].all()) and bool(m[..., 2].all())
    # priority: meta.json beats the arguments, key_frame_indices.json beats meta.json
    json.dump({"active_frame_proportion": 1.0, "use_frame_interval": 4, "start_frame_index": 0}, open(tmp_path / "meta.json", "w"))
    assert get_data_dict_from_folder(args)["key_frame_indices"] == [0, 4, 8]
    json.dump([2, 7], open(tmp_path / "key_frame_indices.json", "w"))
    assert get_data_dict_from_folder(args)["key_frame_indices"] == [2, 7]
    # a user-supplied sky segmenter feeds channel 0; one that raises falls back to the depth buffer
    d2 = get_data_dict_from_folder(args, sky_segmenter=lambda v: np.ones(v.shape[:3], bool))
    assert not bool(d2["foreground_mask_from_seg"].any())
    d3 = get_data_dict_from_folder(args, sky_segmenter=lambda v: (_ for _ in ()).throw(NotImplementedError("no model")))
    assert bool(d3["foreground_mask_from_seg"].all())

"""CAD-mesh voxelisation: PLY loading and the oracle on the CPU; CUDA kernel vs oracle on the GPU."""
import struct

import numpy as np
import pytest
import torch


def _write_ply(path, verts, faces):
    """binary little-endian PLY with the same vertex properties / polygon faces as the reference's car.ply."""
    with open(path, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\ncomment test\n")
        f.write(f"element vertex {len(verts)}\n".encode())
        for p in ("x", "y", "z", "nx", "ny", "nz", "s", "t"):
            f.write(f"property float {p}\n".encode())
        f.write(f"element face {len(faces)}\nproperty list uchar uint vertex_indices\nend_header\n".encode())
        for v in verts:
            f.write(struct.pack("<8f", v[0], v[1], v[2], 0, 0, 1, 0, 0))
        for fc in faces:
            f.write(struct.pack("<B" + "I" * len(fc), len(fc), *fc))


def _car_like_mesh(seed=0):
    """closed box hull (quads) with a cabin on top: ~car proportions, arbitrary (non-lattice) coordinates."""
    rng = np.random.RandomState(seed)
    def box(lo, hi):
        x0, y0, z0 = lo
        x1, y1, z1 = hi
        v = [(x0, y0, z0), (x1, y0, z0), (x1, y1, z0), (x0, y1, z0), (x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)]
        q = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (0, 3, 7, 4)]
        return v, q
    v1, q1 = box((-2.31, -0.93, -0.77), (2.29, 0.91, 0.12))
    v2, q2 = box((-1.13, -0.81, 0.12), (1.02, 0.79, 0.74))
    verts = np.array(v1 + v2, dtype=np.float64) + rng.rand(16, 3) * 0.013
    faces = [tuple(q) for q in q1] + [tuple(i + 8 for i in q) for q in q2]
    return verts, faces


def test_ply_loader_triangulates_polygons(tmp_path):
    from infinicube_b200.raster.mesh import load_ply
    verts, faces = _car_like_mesh()
    p = tmp_path / "car.ply"
    _write_ply(p, verts, faces)
    v, f = load_ply(p)
    assert v.shape == (16, 3) and f.shape == (24, 3)            # 12 quads -> 24 triangles
    assert np.allclose(v, verts.astype(np.float32), atol=1e-6)
    assert tuple(f[0]) == (0, 1, 2) and tuple(f[1]) == (0, 2, 3)
    # ascii variant
    a = tmp_path / "a.ply"
    a.write_text("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                 "element face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n")
    v, f = load_ply(a)
    assert v.shape == (3, 3) and f.tolist() == [[0, 1, 2]]


def test_oracle_voxelises_a_single_triangle():
    from oracle import mesh_oracle as mo
    # right triangle in the plane z = 0.02 with legs of 0.35 starting at (0.01, 0.01): hypotenuse x + y = 0.37
    v = np.array([[0.01, 0.01, 0.02], [0.36, 0.01, 0.02], [0.01, 0.36, 0.02]])
    ijk = mo.voxelize_mesh(v, np.array([[0, 1, 2]]), 0.1, 0.05)
    got = set(map(tuple, ijk))
    assert set(ijk[:, 2]) == {0}
    # voxel (i,j) spans [0.1i, 0.1(i+1)] x [0.1j, 0.1(j+1)]: it is hit iff its lower-left corner is under the hypotenuse
    expect = {(i, j, 0) for i in range(4) for j in range(4) if 0.1 * i + 0.1 * j < 0.37}
    assert got == expect and len(got) == 10
    # closed sets: a triangle edge lying exactly on a voxel face activates both neighbours
    v2 = np.array([[0.0, 0.0, 0.02], [0.3, 0.0, 0.02], [0.0, 0.3, 0.02]])
    assert (-1, 0, 0) in set(map(tuple, mo.voxelize_mesh(v2, np.array([[0, 1, 2]]), 0.1, 0.05)))


@pytest.mark.gpu
def test_cuda_mesh_voxelisation_matches_oracle():
    from oracle import mesh_oracle as mo
    from infinicube_b200.raster.mesh import voxelize_mesh
    verts, faces = _car_like_mesh(1)
    tris = []
    for fc in faces:
        tris += [(fc[0], fc[1], fc[2]), (fc[0], fc[2], fc[3])]
    tris = np.array(tris)
    scale = np.array([4.7, 1.9, 1.6]) / (verts.max(0) - verts.min(0))
    v = verts * scale
    got = voxelize_mesh(v, tris, 0.1, 0.05)
    ref = mo.voxelize_mesh(v, tris, 0.1, 0.05)
    assert len(ref) > 3000
    assert np.array_equal(got, ref)


@pytest.mark.gpu
def test_cad_objects_in_buffer_generation(tmp_path):
    """cad_model_for_static_object / cad_model_for_dynamic_objects as guidance_buffer_generation.py:629-642 passes
    them: car voxels of the scene are dropped, CAD cars are inserted per frame with their instance ids."""
    from oracle import mesh_oracle as mo, raster_oracle as ro
    from infinicube_b200.raster import PinholeCamera, generate_infinicube_buffer_from_fvdb_grid, synthetic as syn
    from infinicube_b200.raster.mesh import load_ply
    dev = torch.device("cuda:0")
    verts, faces = _car_like_mesh(2)
    ply = tmp_path / "car.ply"
    _write_ply(ply, verts, faces)
    vs, S = 0.2, 32
    pts, sem, inst, _ = syn.synthetic_scene(S, voxel_size=vs)
    intr = np.array([100.0, 90.0, 48.0, 27.0, 96, 54])
    cam = PinholeCamera.from_numpy(intr, device=dev)
    poses = torch.from_numpy(syn.synthetic_poses(S, n=2, voxel_size=vs)).to(dev)
    o2w = np.eye(4)
    o2w[:3, 3] = [3.4, 3.1, 1.8]
    info = lambda kind, oid: {f"{f:06d}.{kind}_object_info.json": {  # noqa: E731
        "c%d" % oid: {"object_to_world": (o2w + np.eye(4, k=0) * 0 + np.array([[0, 0, 0, 0.2 * f * (kind == "dynamic")]] * 4) * np.eye(4)[:, 3:4].T * 0).tolist(),
                      "object_lwh": [2.4, 1.1, 0.9], "object_type": "car", "object_id_int": oid}} for f in range(2)}
    static_info, dyn_info = info("static", 7), info("dynamic", 9)
    for f in range(2):  # move the dynamic car away from the static one
        m = np.array(dyn_info[f"{f:06d}.dynamic_object_info.json"]["c9"]["object_to_world"])
        m[:3, 3] = [4.6 + 0.2 * f, 2.0, 1.8]
        dyn_info[f"{f:06d}.dynamic_object_info.json"]["c9"]["object_to_world"] = m.tolist()
    d, s, i = generate_infinicube_buffer_from_fvdb_grid(
        cam, poses, torch.from_numpy(pts).to(dev), torch.from_numpy(sem).to(dev).long(), torch.eye(4), static_info, dyn_info,
        cad_model_for_static_object=True, cad_model_for_dynamic_objects=True, cad_model_location=ply)
    assert set(torch.unique(i).tolist()) <= {0, 7, 9} and (i == 7).any() and (i == 9).any()
    # oracle composition for frame 1
    v, tr = load_ply(ply)
    mesh_lwh = v.max(0) - v.min(0)
    cad = mo.voxelize_mesh(v * (np.array([2.4, 1.1, 0.9]) / mesh_lwh), tr, 0.1, 0.05).astype(np.float32) * np.float32(0.1) + np.float32(0.05)
    keep = ~np.isin(sem, [1, 2, 3, 4])
    all_pts, all_sem, all_inst = [pts[keep]], [sem[keep]], [np.zeros(keep.sum(), np.int32)]
    for dic, key, oid in ((dyn_info, "000001.dynamic_object_info.json", 9), (static_info, "000001.static_object_info.json", 7)):
        m = np.array(list(dic[key].values())[0]["object_to_world"])
        p = (m[:3, :3] @ cad.astype(np.float64).T + m[:3, 3:4]).T.astype(np.float32)
        all_pts.append(p)
        all_sem.append(np.full(len(p), 1, np.int32))
        all_inst.append(np.full(len(p), oid, np.int32))
    og = ro.OracleGrid(np.concatenate(all_pts), [vs] * 3, [vs / 2] * 3, np.concatenate(all_sem), np.concatenate(all_inst))
    od, os_, oi = og.render(ro.inv_intrinsics_matrix(intr), poses[1:2].cpu().numpy(), 96, 54)
    assert np.array_equal(s[1].cpu().numpy(), os_[0]) and np.array_equal(i[1].cpu().numpy(), oi[0])
    assert np.array_equal(d[1].cpu().numpy(), od[0])

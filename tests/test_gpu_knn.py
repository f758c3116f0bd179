"""GPU parity of the nearest-neighbour label transfer against oracle/knn_oracle.c (pinned to
scipy's cKDTree in tests/test_oracle_knn.py).  Integer / index work: the bar is bit-exact — indices, labels and the
fp32 squared distances — including ties (smallest reference index)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import knn_oracle as ko  # noqa: E402  (checker only)


@pytest.fixture(scope="module")
def dev():
    from infinicube_b200 import _lib
    _lib.require_device()
    return torch.device("cuda:0")


def _lattice(n, seed, span=60, vs=0.2, angle=0.05):
    rs = np.random.RandomState(seed)
    p = rs.randint(0, span, size=(n, 3)).astype(np.float32) * np.float32(vs) + np.float32(vs / 2)
    R = np.array([[np.cos(angle), -np.sin(angle), 0], [np.sin(angle), np.cos(angle), 0], [0, 0, 1]], dtype=np.float32)
    return (p @ R.T).astype(np.float32)


def _uniform(n, seed, lo=(-20, -15, -3), hi=(20, 15, 3)):
    rs = np.random.RandomState(seed)
    return (rs.rand(n, 3) * (np.array(hi) - np.array(lo)) + np.array(lo)).astype(np.float32)


CASES = {
    "uniform_auto_cell": lambda: (_uniform(5000, 1), _uniform(4000, 2), 0.0),
    "lattice_voxel_cell": lambda: (_lattice(6000, 3), _lattice(5000, 4, angle=0.0), 0.2),   # many exact ties
    "queries_outside_bbox": lambda: (_uniform(3000, 5), _uniform(2000, 6, lo=(-40, -30, -10), hi=(40, 30, 10)), 0.5),
    "tiny_reference": lambda: (_uniform(3, 7), _uniform(500, 8), 0.0),                      # below the reference's 64-point switch
    "single_point": lambda: (_uniform(1, 9), _uniform(50, 10), 0.0),
    "duplicates": lambda: (np.repeat(_uniform(40, 11), 5, axis=0), _uniform(700, 12), 1.0),
    "flat_sheet": lambda: (_uniform(4000, 13, lo=(0, 0, 1.5), hi=(50, 40, 1.5)), _uniform(3000, 14, lo=(0, 0, 0), hi=(50, 40, 3)), 0.0),
    "coarse_cells": lambda: (_uniform(3000, 15), _uniform(1000, 16), 25.0),                 # 2 x 2 x 1 cells: long row scans
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_nn1_bit_exact_vs_oracle(dev, name):
    from infinicube_b200.voxelgen.utils.color_util import KnnIndex
    ref, q, cell = CASES[name]()
    sem = (np.arange(ref.shape[0], dtype=np.int64) * 7919) % 23
    index = KnnIndex(torch.from_numpy(ref).to(dev), cell)
    d2, idx, lab = index.query(torch.from_numpy(q).to(dev), torch.from_numpy(sem))
    oi, od = ko.nn1(q, ref)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(d2.cpu().numpy(), od)
    assert np.array_equal(lab.cpu().numpy(), sem[oi])
    info = index.info()
    assert info["points"] == ref.shape[0] and info["cells"] == int(np.prod(info["dims"]))


def test_reference_api_surface(dev):
    """knn_query_fast / semantic_from_points keep the reference's conventions (knn.cu:15-50, color_util.py:52-60)."""
    from infinicube_b200._lib import ICError
    from infinicube_b200.voxelgen import knn_query_fast, semantic_from_points
    ref, q = _uniform(2000, 21), _uniform(300, 22)
    sem = torch.from_numpy(np.random.RandomState(0).randint(0, 30, size=2000)).to(dev)
    dist, idx = knn_query_fast(torch.from_numpy(q).to(dev), torch.from_numpy(ref).to(dev), 1)
    assert tuple(dist.shape) == (300, 1) and dist.dtype == torch.float32
    assert tuple(idx.shape) == (300, 1) and idx.dtype == torch.int32
    out = semantic_from_points(torch.from_numpy(q).to(dev), torch.from_numpy(ref).to(dev), sem)
    assert out.dtype == torch.int64 and torch.equal(out, sem[idx[:, 0].long()])
    assert np.array_equal(out.cpu().numpy(), ko.semantic_from_points(q, ref, sem.cpu().numpy()))
    empty = semantic_from_points(torch.zeros(0, 3, device=dev), torch.from_numpy(ref).to(dev), sem)
    assert tuple(empty.shape) == (0,) and empty.dtype == torch.int64
    # strided (N, 4) inputs, as the reference's index wants them
    ref4 = torch.zeros(2000, 4, device=dev)
    ref4[:, :3] = torch.from_numpy(ref).to(dev)
    _, idx4 = knn_query_fast(torch.from_numpy(q).to(dev), ref4, 1)
    assert torch.equal(idx4, idx)
    d8, i8 = knn_query_fast(torch.from_numpy(q).to(dev), torch.from_numpy(ref).to(dev), 8)
    assert tuple(d8.shape) == (300, 8) and torch.equal(i8[:, 0], idx[:, 0]) and torch.equal(d8[:, 0], dist[:, 0])
    with pytest.raises(ICError):
        knn_query_fast(torch.from_numpy(q).to(dev), torch.from_numpy(ref).to(dev), 33)
    with pytest.raises(ICError):
        knn_query_fast(torch.from_numpy(q), torch.from_numpy(ref), 1)
    with pytest.raises(TypeError):
        knn_query_fast(torch.from_numpy(q).to(dev).double(), torch.from_numpy(ref).to(dev), 1)


def test_full_size_properties(dev):
    """2 M voxel centres against 2 M (a stage-1 chunk merge): identity, consistency, and a sampled oracle check."""
    from infinicube_b200.voxelgen.utils.color_util import KnnIndex
    g = torch.Generator(device="cpu").manual_seed(0)
    ijk = torch.unique(torch.randint(0, 400, (2_400_000, 3), generator=g) * torch.tensor([1, 1, 0]) +
                       torch.randint(0, 40, (2_400_000, 3), generator=g) * torch.tensor([0, 0, 1]), dim=0)
    ref = (ijk.float() * 0.2 + 0.1).to(dev)
    m = ref.shape[0]
    sem = (torch.arange(m, device=dev) % 19).long()
    index = KnnIndex(ref, 0.2)
    # (1) identity: every reference point is its own nearest neighbour at distance 0
    d2, idx, lab = index.query(ref, sem)
    assert torch.equal(idx.long(), torch.arange(m, device=dev)) and float(d2.max()) == 0.0 and torch.equal(lab, sem)
    # (2) rigid motion of the cloud (the extrapolation use): returned d2 is the distance to the returned point
    c, s = np.cos(0.03), np.sin(0.03)
    R = torch.tensor([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=torch.float32, device=dev)
    q = ref @ R.T + torch.tensor([0.07, -0.04, 0.02], device=dev)
    d2, idx, _ = index.query(q)
    dd = q - ref[idx.long()]
    assert float(((dd * dd).sum(1) - d2).abs().max()) <= 1e-5
    assert float(d2.max()) < 40.0 ** 2
    # (3) a 300-query sample against the brute-force oracle, bit-exact
    pick = torch.randperm(m, generator=g)[:300]
    oi, od = ko.nn1(q[pick.to(dev)].cpu().numpy(), ref.cpu().numpy())
    assert np.array_equal(idx[pick.to(dev)].cpu().numpy(), oi) and np.array_equal(d2[pick.to(dev)].cpu().numpy(), od)


@pytest.mark.parametrize("k", [2, 3, 8, 16])
def test_knn_k_bit_exact_vs_oracle(k):
    """knn_query_fast(q, ref, k) (knn.cu:15-51) for k > 1: indices and fp32 squared distances bit-exact against the
    brute-force definition, incl. ties (lattice cloud), rows sorted, cell size irrelevant to the result."""
    from oracle import knn_oracle as ko
    from infinicube_b200 import _lib
    from infinicube_b200.voxelgen.utils.color_util import KnnIndex, color_from_points, knn_query_fast
    _lib.require_device()
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(k)
    ref = (rs.randint(0, 24, size=(3000, 3)) * 0.2 + 0.1).astype(np.float32)          # voxel centres: many exact ties
    ref = np.unique(ref, axis=0)
    q = (ref[rs.randint(0, len(ref), size=700)] + rs.randn(700, 3).astype(np.float32) * 0.07).astype(np.float32)
    q[:50] = ref[rs.randint(0, len(ref), size=50)]                                    # queries on lattice points
    want_i, want_d = ko.nnk(q, ref, k)
    for cell in (0.0, 0.2, 0.7):
        d2, idx = knn_query_fast(torch.from_numpy(q).to(dev), torch.from_numpy(ref).to(dev), k, cell_size=cell)
        assert idx.dtype == torch.int32 and idx.shape == (700, k) and d2.shape == (700, k)
        assert np.array_equal(d2.cpu().numpy(), want_d)
        assert np.array_equal(idx.cpu().numpy(), want_i)
    # fewer reference points than k
    d2, idx = KnnIndex(torch.from_numpy(ref[:3]).to(dev)).query_k(torch.from_numpy(q[:9]).to(dev), k)
    wi, wd = ko.nnk(q[:9], ref[:3], k)
    assert np.array_equal(idx.cpu().numpy(), wi) and np.array_equal(d2.cpu().numpy(), wd)
    if k == 8:   # color_from_points (color_util.py:21-49)
        colors = rs.rand(len(ref), 3).astype(np.float32)
        got = color_from_points(torch.from_numpy(q).to(dev), torch.from_numpy(ref).to(dev), torch.from_numpy(colors).to(dev), k=8)
        assert np.allclose(got.cpu().numpy(), ko.color_from_points(q, ref, colors, 8), atol=2e-6)

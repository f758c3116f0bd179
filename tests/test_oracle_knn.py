"""Pins oracle/knn_oracle.c (semantic_from_points / knn_query_fast with k = 1) against
scipy.spatial.cKDTree, an independent exact nearest-neighbour search."""
import numpy as np
import pytest

from oracle import knn_oracle as ko


def _cloud(n, seed, lattice=False):
    rs = np.random.RandomState(seed)
    if lattice:  # voxel centres on a 0.2 m lattice, rotated a little (the stage-1 chunk-merge shape)
        ijk = rs.randint(0, 60, size=(n, 3)).astype(np.float32)
        p = ijk * 0.2 + 0.1
        a = 0.05
        R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], dtype=np.float32)
        return (p @ R.T).astype(np.float32)
    return (rs.rand(n, 3) * np.array([40, 30, 6]) - np.array([20, 15, 3])).astype(np.float32)


@pytest.mark.parametrize("n,m,lattice", [(2000, 3000, False), (1500, 4000, True), (10, 1, False), (300, 63, False)])
def test_nn1_matches_ckdtree(n, m, lattice):
    from scipy.spatial import cKDTree
    ref, q = _cloud(m, 1, lattice), _cloud(n, 2, lattice)
    idx, d2 = ko.nn1(q, ref)
    dist, j = cKDTree(ref.astype(np.float64)).query(q.astype(np.float64), k=1)
    # distances agree to fp32 rounding; the oracle's pick is optimal in fp32 and (to rounding) in fp64
    assert np.allclose(np.sqrt(d2.astype(np.float64)), dist, rtol=2e-6, atol=1e-6)
    d2_at_scipy = ((q - ref[j]) ** 2).sum(1)
    assert np.all(d2 <= d2_at_scipy * (1 + 1e-6) + 1e-12)
    d64 = np.sqrt(((q.astype(np.float64) - ref[idx].astype(np.float64)) ** 2).sum(1))
    assert np.all(d64 <= dist * (1 + 1e-5) + 1e-7)
    if not lattice:   # generic positions: no ties, so the index itself must agree (lattice clouds tie legitimately)
        assert np.mean(idx == j) > 0.999
    else:             # where the picks differ it is a tie at fp32 resolution, and the oracle holds the smaller index
        diff = idx != j
        assert np.all(np.abs(d2_at_scipy[diff] - d2[diff]) <= 1e-5 * np.maximum(d2[diff], 1e-6))
        exact_tie = diff & (d2_at_scipy == d2)
        assert np.all(idx[exact_tie] < j[exact_tie])


def test_ties_pick_smallest_index_and_labels_transfer():
    ref = np.array([[0, 0, 0], [2, 0, 0], [2, 0, 0], [0, 2, 0]], dtype=np.float32)
    q = np.array([[1, 0, 0], [2, 0.1, 0], [0, 1, 0], [5, 5, 5]], dtype=np.float32)
    idx, d2 = ko.nn1(q, ref)
    assert idx.tolist() == [0, 1, 0, 1] and np.allclose(d2[:3], [1, 0.01, 1])
    sem = np.array([7, 8, 9, 10])
    assert ko.semantic_from_points(q, ref, sem).tolist() == [7, 8, 7, 8]
    out = ko.semantic_from_points(np.zeros((0, 3), np.float32), ref, sem)
    assert out.shape == (0,) and out.dtype == np.int64


def test_product_knn_has_no_cpu_path():
    """The product mirror of semantic_from_points / knn_query_fast raises on CPU tensors instead of falling back."""
    import torch
    from infinicube_b200._lib import ICError
    from infinicube_b200.voxelgen import knn_query_fast, semantic_from_points
    q, r = torch.rand(5, 3), torch.rand(9, 3)
    with pytest.raises(ICError):
        knn_query_fast(q, r, 1)
    with pytest.raises(ICError):
        semantic_from_points(q, r, torch.zeros(9, dtype=torch.long))
    with pytest.raises(ICError):
        knn_query_fast(q, r, 8)       # k > 1 is not built on this path
    # the reference's empty-target convention needs no device work (color_util.py:53-54)
    out = semantic_from_points(torch.zeros(0, 3), r, torch.zeros(9, dtype=torch.long))
    assert out.shape == (0,) and out.dtype == torch.int64


def test_oracle_matches_the_references_small_cloud_branch():
    """knn_query_fast switches to `torch.cdist` + `topk` below 64 reference points (knn.cu:23-28) - the one branch of
    the reference's search that runs without its CUDA KD-tree.  Restated here with the same torch calls on CPU."""
    import torch
    rs = np.random.RandomState(5)
    for m in (1, 2, 17, 63):
        ref = (rs.rand(m, 3) * 10).astype(np.float32)
        q = (rs.rand(400, 3) * 12 - 1).astype(np.float32)
        cd = torch.cdist(torch.from_numpy(q), torch.from_numpy(ref))
        dist, idx = torch.topk(cd, 1, dim=1, largest=False)
        d2_ref, idx_ref = torch.square(dist)[:, 0].numpy(), idx[:, 0].to(torch.int32).numpy()
        oi, od = ko.nn1(q, ref)
        assert np.allclose(od, d2_ref, rtol=1e-4, atol=1e-5)          # sqrt-then-square vs direct squared distance
        same = oi == idx_ref
        # a different pick is only acceptable as a tie at the branch's own (cdist) resolution
        assert np.all(np.abs(cd.numpy()[~same, oi[~same]] - dist.numpy()[~same, 0]) <= 1e-5)
        assert same.mean() > 0.99


def test_knn_k_oracle_matches_ckdtree_and_the_reference_small_cloud_branch():
    """nnk against scipy's exact k-NN and against the reference's own `ref < 64` branch (torch.cdist + topk,
    knn.cu:23-28) restated with the same torch calls."""
    import torch
    from scipy.spatial import cKDTree
    from oracle import knn_oracle as ko
    rs = np.random.RandomState(5)
    ref = rs.rand(500, 3).astype(np.float32) * 10
    q = rs.rand(200, 3).astype(np.float32) * 10
    for k in (2, 8):
        idx, d2 = ko.nnk(q, ref, k)
        dd, ii = cKDTree(ref.astype(np.float64)).query(q.astype(np.float64), k=k)
        assert np.array_equal(idx, ii.astype(np.int32))                  # random cloud: no ties
        assert np.allclose(d2, dd ** 2, rtol=1e-5)
        assert np.all(np.diff(d2, axis=1) >= 0)
    small = ref[:40]
    cd = torch.cdist(torch.from_numpy(q), torch.from_numpy(small))
    top = torch.topk(cd, 4, 1, False)
    idx, d2 = ko.nnk(q, small, 4)
    assert np.array_equal(idx, top.indices.numpy().astype(np.int32))
    assert np.allclose(d2, top.values.numpy() ** 2, rtol=2e-3, atol=1e-5)   # cdist's matmul path is the imprecise side
    # ties resolve to the smaller reference index; fewer reference points than k pad with (-1, inf)
    lat = np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0]], np.float32)
    idx, d2 = ko.nnk(np.zeros((1, 3), np.float32), lat, 6)
    assert idx.tolist() == [[0, 1, 2, 3, -1, -1]] and np.isinf(d2[0, 4:]).all()
    col = ko.color_from_points(q[:5], ref, rs.rand(500, 3).astype(np.float32), k=8)
    assert col.shape == (5, 3) and np.all((col >= 0) & (col <= 1))

"""Oracle cloud primitives against independent restatements available in the container."""
import numpy as np
import pytest

import oracle
import synth


def _numpy_voxel(xyzi, leaf):
    """Independent numpy statement of pcl::VoxelGrid keys + membership."""
    p = xyzi[:, :3].astype(np.float32)
    inv = np.float32(1.0) / np.float32(leaf)
    mn, mx = p.min(0), p.max(0)
    min_b = np.floor(mn * inv).astype(np.int32)
    max_b = np.floor(mx * inv).astype(np.int32)
    div = max_b - min_b + 1
    ijk = (np.floor(p * inv) - min_b.astype(np.float32)).astype(np.int32)
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    return key


def test_voxel_keys_membership_and_centroids():
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.uniform(-20, 20, (5000, 3)), rng.uniform(0, 255, (5000, 1))], 1).astype(np.float32)
    pts[:2500, 2] = rng.normal(0, 0.01, 2500)            # a dense floor so voxels hold many points
    r = oracle.voxel_grid(pts, 0.4)
    key = _numpy_voxel(pts, 0.4)
    assert np.array_equal(r["point_keys"], key)
    uk = np.unique(key)
    assert np.array_equal(r["out_keys"], uk)
    # centroid = sequential f32 sum in point-index order / count
    for j in rng.choice(len(uk), 50, replace=False):
        members = np.nonzero(key == uk[j])[0]
        acc = np.zeros(4, np.float32)
        for m in members:
            acc = (acc + pts[m]).astype(np.float32)
        want = acc / np.float32(len(members))
        assert np.array_equal(r["points"][j], want)


def test_voxel_empty_single_and_overflow():
    assert oracle.voxel_grid(np.zeros((0, 4), np.float32), 0.2)["points"].shape[0] == 0
    one = np.array([[1.5, -2.5, 3.5, 7.0]], np.float32)
    r = oracle.voxel_grid(one, 0.2)
    assert np.array_equal(r["points"], one)
    far = np.array([[0, 0, 0, 1], [1e6, 1e6, 1e6, 2]], np.float32)
    r = oracle.voxel_grid(far, 0.01)
    assert r["overflow"] and np.array_equal(r["points"], far)


def test_kdtree_equals_bruteforce_with_index_tiebreak():
    rng = np.random.default_rng(1)
    m = np.concatenate([rng.uniform(-10, 10, (20000, 3)), np.zeros((20000, 1))], 1).astype(np.float32)
    m[5000:5200] = m[4800:5000]                              # exact duplicates: distance ties
    q = rng.uniform(-11, 11, (3000, 3)).astype(np.float32)
    q[:200] = m[4800:5000, :3]
    i_t, d_t = oracle.knn5(m, q)
    i_b, d_b = oracle.knn5(m, q, brute=True)
    assert np.array_equal(i_t, i_b)
    assert np.array_equal(d_t, d_b)
    assert np.all(np.diff(d_t, axis=1) >= 0)


def test_kdtree_agrees_with_scipy_sets():
    sp = pytest.importorskip("scipy.spatial")
    fr = synth.make_frame(1, 3, small=(16, 600, 4000, 20000))
    m = fr["map_surf"]
    rng = np.random.default_rng(2)
    q = (m[rng.choice(len(m), 2000), :3] + rng.normal(0, 0.05, (2000, 3))).astype(np.float32)
    idx, d2 = oracle.knn5(m, q)
    _, ref = sp.cKDTree(m[:, :3].astype(np.float64)).query(q.astype(np.float64), k=5)
    same = [set(a) == set(b) for a, b in zip(idx, ref)]
    assert np.mean(same) > 0.999                             # f32 vs f64 distance rounding can flip a near-tie


def test_kdtree_small_maps():
    m = np.array([[0, 0, 0, 0], [1, 0, 0, 0], [0, 2, 0, 0]], np.float32)
    idx, d2 = oracle.knn5(m, np.array([[0.1, 0, 0]], np.float32))
    assert list(idx[0][:3]) == [0, 1, 2] and idx[0][3] == np.iinfo(np.int32).max


def test_crop_box_inclusive_and_order_preserving():
    pts = np.array([[0, 0, 0, 1], [30, 0, 0, 2], [30.0001, 0, 0, 3], [-30, -30, -10, 4], [1, 1, 10.5, 5]], np.float32)
    out = oracle.crop_box(pts, [-30, -30, -10], [30, 30, 10])
    assert list(out[:, 3]) == [1, 2, 4]


def test_kdtree_split_rules_give_identical_results():
    """The oracle's kd-tree is built with FLANN's middleSplit_ rule by default (CPU-timing fidelity); the
    median-split build and brute force must return the same index sets and squared distances (the search is exact)."""
    rng = np.random.default_rng(3)
    # clustered + duplicated points: planeSplit's == cutval band and degenerate boxes are exercised
    pts = np.concatenate([rng.uniform(-20, 20, (6000, 3)), np.repeat(rng.uniform(-1, 1, (40, 3)), 25, axis=0),
                          np.full((40, 3), 2.5), rng.normal(0, 0.05, (2000, 3)) + [5, 5, 0]]).astype(np.float32)
    m = np.concatenate([pts, np.arange(len(pts), dtype=np.float32)[:, None]], 1)
    q = np.concatenate([rng.uniform(-21, 21, (400, 3)), pts[::40] + np.float32(0.01)]).astype(np.float32)
    assert oracle.set_kdtree_flann_split(1) == 1                  # the default
    try:
        i_f, d_f = oracle.knn5(m, q)
        oracle.set_kdtree_flann_split(0)
        i_m, d_m = oracle.knn5(m, q)
    finally:
        oracle.set_kdtree_flann_split(1)
    i_b, d_b = oracle.knn5(m, q, brute=True)
    assert np.array_equal(i_f, i_b) and np.array_equal(d_f, d_b)
    assert np.array_equal(i_m, i_b) and np.array_equal(d_m, d_b)

"""GPU parity on the configurations that bench.py actually runs ("parity holes"):

  - 128 independent BASELINE-config-4 frames in ONE batched launch (the benched shape), for every LM cluster size, every
    frame's iteration count / flags / pose against the CPU oracle, with per-point debug capture (kNN sets, d^2, flags,
    coefficients, AtA, AtB, X, pose trace) for slots of that batched launch;
  - the two single-frame shapes (whole-GPU cooperative grid, one cluster);
  - the committed golden fixture tests/golden/pipeline_small.npz through the whole CUDA path;
  - the DEVICE small-matrix routines directly against the committed cv2 vectors;
  - the feature kernel's full-sort fallback (Horizon_SCAN > 3060 and < 180) and a seeded fuzz of the parallel corner /
    flat-loop rounds against the sequential oracle.
"""
import os

import numpy as np
import pytest

import oracle
import synth

pytestmark = pytest.mark.gpu

POSE_TOL = 1e-4
RES_REL_TOL = 1e-5
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_BATCH = int(os.environ.get("FBPR_TEST_BATCH", "128"))


@pytest.fixture(scope="module")
def fb():
    import feature_base_pointcloud_registration_b200 as m
    return m


def _oracle_whole_path(fr, debug_iter=-1, threads=None):
    P = dict(fr["params"])
    if threads:
        P["numberOfCores"] = threads
    ci = oracle.project(P, fr["scan"], fr["imu"], fr["imu_available"])
    fe = oracle.extract_features(P, ci)
    mo = oracle.MapOptimization(P)
    mo.set_imu(fr["imu_available"], 0.0, 0.0)
    mo.set_scan(fe["corner"], fe["surface"]); mo.set_map(fr["map_corner"], fr["map_surf"]); mo.downsample()
    pose, iters, flags, _ = mo.scan2map(fr["guess"], debug_iter=debug_iter)
    return mo, pose, iters, flags, fe


@pytest.fixture(scope="module")
def batch(fb):
    """N_BATCH config-4 frames, their host inputs and the oracle's answer for every one of them."""
    frames = [synth.make_frame(4, 300 + i) for i in range(N_BATCH)]
    want = np.zeros(N_BATCH, fb.api.RESULT_DTYPE)
    dbg = {}
    feats = []
    DEBUG_ITER = 1
    for i, fr in enumerate(frames):
        mo, pose, iters, flags, fe = _oracle_whole_path(fr, debug_iter=DEBUG_ITER if i < 2 else -1, threads=os.cpu_count())
        want[i] = (pose, iters, flags)
        feats.append((fe["corner"], fe["surface"]))
        if i < 2:
            d = mo.debug(); d["trace"] = mo.pose_trace()
            dbg[i] = d
    raws = [fb.api.pack_raw(fr["scan"]) for fr in frames]
    return dict(frames=frames, raws=raws, want=want, dbg=dbg, debug_iter=DEBUG_ITER, feats=feats)


def _inputs(reg, batch, lo=0, hi=None):
    frames, raws = batch["frames"][lo:hi], batch["raws"][lo:hi]
    return reg.make_frame_inputs([dict(raw_ptr=raw.ctypes.data, n_raw=len(raw), imu=fr["imu"], imu_available=fr["imu_available"],
                                       map_corner_ptr=fr["map_corner"].ctypes.data, n_map_corner=len(fr["map_corner"]),
                                       map_surf_ptr=fr["map_surf"].ctypes.data, n_map_surf=len(fr["map_surf"]), pose=fr["guess"])
                                  for fr, raw in zip(frames, raws)])


def _check_results(got, want):
    assert np.array_equal(got["iters"], want["iters"]), np.flatnonzero(got["iters"] != want["iters"])
    assert np.array_equal(got["flags"], want["flags"]), np.flatnonzero(got["flags"] != want["flags"])
    err = np.abs(got["pose"] - want["pose"])
    assert err.max() <= POSE_TOL, (err.max(), np.unravel_index(err.argmax(), err.shape))


def _check_debug(r, slot, d, debug_iter):
    """per-point capture of one LM iteration of `slot` against the oracle's IterDebug"""
    assert d["iter"] == debug_iter
    for kind, K in (("CORNER", "corner"), ("SURF", "surf")):
        knn = r.get_buffer(slot, "KNN_" + kind); d2 = r.get_buffer(slot, "KNN_D2_" + kind)
        accept = d[K + "D2"][:, 4] < 1.0
        assert np.array_equal(knn[accept], d[K + "Knn"][accept]), kind          # kNN index sets bit-exact, ties by index
        assert np.array_equal(d2[accept], d[K + "D2"][accept]), kind
        assert np.all(knn[~accept] == -1)
        flag = r.get_buffer(slot, "FLAG_" + kind)
        assert np.array_equal(flag, d[K + "Flag"]), kind
        sel = flag.astype(bool)
        co = r.get_buffer(slot, "COEFF_" + kind)[sel]; cw = d[K + "Coeff"][sel]
        rel = np.abs(co - cw) / np.maximum(np.abs(cw), 1e-3)
        assert rel.max() <= RES_REL_TOL, (kind, rel.max())
    # normal equations: f64 sums in a different association order, rounded once to f32 -> equal up to that one rounding
    assert np.allclose(r.get_buffer(slot, "ATA"), d["AtA"], rtol=1e-6, atol=0)
    assert np.allclose(r.get_buffer(slot, "ATB"), d["AtB"], rtol=1e-5, atol=1e-6)
    assert np.allclose(r.get_buffer(slot, "X"), d["X"], rtol=1e-3, atol=1e-7)     # the 6x6 QR solve amplifies the last-bit difference of AtA
    trace = r.get_buffer(slot, "POSE_TRACE")[: len(d["trace"])]
    assert np.abs(trace - d["trace"]).max() <= POSE_TOL                           # pose after EVERY iteration, not only the last


@pytest.mark.parametrize("cluster", [0, 1, 2, 4, 8, 16])
def test_benched_batch_every_cluster_size(fb, batch, cluster):
    """The shape bench.py times: N_BATCH frames in one run_frames call, lm_kernel<false>, cluster size forced (0 = auto)."""
    F = N_BATCH
    r = fb.Registration(batch["frames"][0]["params"], max_frames=F, max_map_corner=40064, max_map_surf=160064, lm_cluster_size=cluster)
    guesses = np.stack([fr["guess"] for fr in batch["frames"]])
    r.set_frames(0, _inputs(r, batch))
    r.run_frames(0, F)                                    # the whole path from the raw sweeps, as benched
    got = r.get_results(0, F)
    _check_results(got, batch["want"])
    # a second pass over the same slots (poses restored) gives identical bytes: no state leaks between launches
    r.set_poses(0, guesses)
    r.run_frames(0, F)
    again = r.get_results(0, F)
    assert np.array_equal(again["pose"], got["pose"]) and np.array_equal(again["iters"], got["iters"])
    # the registration half on the oracle's own feature clouds (bit-identical LM inputs on both sides; the deskewed projection
    # may differ from glibc's trig in the last ulp of a few points), with per-point capture of slots INSIDE the batched launch
    for s, (c, sf) in enumerate(batch["feats"]):
        r.set_feature_clouds(s, c, sf)
    r.set_poses(0, guesses)
    r.set_debug_iteration(batch["debug_iter"])
    r.run_frames(0, F, with_projection=False, with_features=False)
    _check_results(r.get_results(0, F), batch["want"])
    for slot in (0, 1):
        _check_debug(r, slot, batch["dbg"][slot], batch["debug_iter"])
    r.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_single_frame_shapes(fb, batch, mode):
    """count == 1 calls: lm_single_frame_mode 0 = cooperative grid over the whole GPU, 1 = one cluster."""
    n = min(8, N_BATCH)
    r = fb.Registration(batch["frames"][0]["params"], max_frames=n, max_map_corner=40064, max_map_surf=160064, lm_single_frame_mode=mode)
    r.set_frames(0, _inputs(r, batch, 0, n))
    for s in range(n):
        r.run_frames(s, 1)
    _check_results(r.get_results(0, n), batch["want"][:n])
    r.set_poses(0, np.stack([fr["guess"] for fr in batch["frames"][:n]]))
    r.set_debug_iteration(batch["debug_iter"])
    for s in range(n):
        r.set_feature_clouds(s, *batch["feats"][s])
        r.run_frames(s, 1, with_projection=False, with_features=False)
    _check_results(r.get_results(0, n), batch["want"][:n])
    for slot in (0, 1):
        _check_debug(r, slot, batch["dbg"][slot], batch["debug_iter"])
    r.close()


def test_pipelined_batch_equals_oracle(fb, batch):
    """fbpr_register_frames_begin/_end (the e2e path of bench.py) on the same frames, two batches in flight."""
    F = min(N_BATCH, 64)
    r = fb.Registration(batch["frames"][0]["params"], max_frames=2 * F, max_map_corner=40064, max_map_surf=160064)
    fin = _inputs(r, batch, 0, F)
    t0 = r.register_frames_begin(0, fin, 16)
    t1 = r.register_frames_begin(F, fin, 0)
    _check_results(r.register_frames_end(t0), batch["want"][:F])
    _check_results(r.register_frames_end(t1), batch["want"][:F])
    r.close()


# ------------------------------------------------------------------ committed golden fixture through the CUDA path
def test_golden_pipeline_small_fixture(fb):
    g = np.load(os.path.join(GOLD, "pipeline_small.npz"))
    n_scan, horizon, mc, ms = (int(v) for v in g["small"])
    P = synth.params_for(int(g["config"])); P["N_SCAN"] = n_scan; P["Horizon_SCAN"] = horizon
    n = int(g["n_raw"])
    raw = np.zeros(n, fb.api.RAW_POINT_DTYPE)
    for k, src in (("x", "raw_x"), ("y", "raw_y"), ("z", "raw_z"), ("intensity", "raw_i"), ("ring", "raw_ring"), ("time", "raw_time")):
        raw[k] = g[src][:n]
    imu = synth.make_imu_ramp(synth.IMU_RATES)
    r = fb.Registration(P, max_frames=1, max_map_corner=mc + 64, max_map_surf=ms + 64)
    r.set_raw_scan(0, raw, imu=imu, imu_available=1)
    r.set_local_map(0, g["map_corner"], g["map_surf"])
    r.set_pose(0, g["guess"])
    r.set_debug_iteration(0)
    r.run_frames(0, 1)
    r.sync()
    assert np.array_equal(r.get_buffer(0, "START_RING"), g["startRing"]) and np.array_equal(r.get_buffer(0, "END_RING"), g["endRing"])
    assert np.array_equal(r.get_buffer(0, "COL_IND"), g["colInd"].astype(np.int32))
    assert np.array_equal(r.get_buffer(0, "RANGE"), g["rng"])
    assert np.array_equal(r.get_buffer(0, "WINNER_RAW"), g["winner"])
    cloud = r.get_buffer(0, "CLOUD")
    assert np.allclose(cloud, g["cloud"], rtol=0, atol=2e-5)        # deskew trig: device f64 sincos vs glibc, <= 1 ulp of the f32 result
    exact = np.array_equal(cloud, g["cloud"])
    if exact:                                                         # everything downstream is then bit-comparable
        assert np.array_equal(r.get_buffer(0, "LABEL"), g["label"].astype(np.int32))
        assert np.array_equal(r.get_buffer(0, "PICKED"), g["picked"].astype(np.int32))
        assert np.array_equal(r.get_buffer(0, "CORNER_INDEX"), g["corner_index"])
        assert np.array_equal(r.get_buffer(0, "SURF"), g["surface"])
        assert np.array_equal(r.get_buffer(0, "CORNER_DS"), g["cornerDS"]) and np.array_equal(r.get_buffer(0, "SURF_DS"), g["surfDS"])
        assert np.array_equal(r.get_buffer(0, "FLAG_CORNER"), g["cornerFlag"]) and np.array_equal(r.get_buffer(0, "FLAG_SURF"), g["surfFlag"])
        kc = r.get_buffer(0, "KNN_CORNER"); ks = r.get_buffer(0, "KNN_SURF")
        ac = g["cornerKnn"][:, 0] >= 0; as_ = g["surfKnn"][:, 0] >= 0
        assert np.array_equal(kc[ac & (kc[:, 0] >= 0)], g["cornerKnn"][ac & (kc[:, 0] >= 0)])
        assert np.array_equal(ks[as_ & (ks[:, 0] >= 0)], g["surfKnn"][as_ & (ks[:, 0] >= 0)])
        assert np.allclose(r.get_buffer(0, "ATA"), g["AtA"], rtol=1e-6, atol=0)
        assert np.allclose(r.get_buffer(0, "ATB"), g["AtB"], rtol=1e-5, atol=1e-6)
        assert np.allclose(r.get_buffer(0, "X"), g["X"], rtol=1e-3, atol=1e-7)
    else:                                                             # a 1-ulp coordinate may flip nothing or a voxel; labels come from ranges only
        assert np.array_equal(r.get_buffer(0, "LABEL"), g["label"].astype(np.int32))
        assert np.array_equal(r.get_buffer(0, "CORNER_INDEX"), g["corner_index"])
    pose, iters, flags = r.get_pose(0)
    assert (iters, flags) == (int(g["iters"]), int(g["flags"]))
    assert np.abs(pose - g["pose"]).max() <= POSE_TOL
    tr = r.get_buffer(0, "POSE_TRACE")[: len(g["pose_trace"])]
    assert np.abs(tr - g["pose_trace"]).max() <= POSE_TOL
    r.close()


# ------------------------------------------------------------------ device small-matrix routines vs cv2's own results
def test_device_smallmat_against_committed_cv2_vectors(fb):
    g = np.load(os.path.join(GOLD, "smallmat_cv2.npz"))
    r = fb.Registration(synth.params_for(1), max_frames=1, max_map_corner=1024, max_map_surf=1024)
    out = r.selftest_smallmat("JACOBI3", g["A3"].reshape(-1, 9))
    assert np.array_equal(out[:, :3], g["W3"]) and np.array_equal(out[:, 3:].reshape(-1, 3, 3), g["V3"])       # cv::eigen 3x3, bit for bit
    out = r.selftest_smallmat("JACOBI6", g["A6"].reshape(-1, 36))
    assert np.array_equal(out[:, :6], g["W6"]) and np.array_equal(out[:, 6:].reshape(-1, 6, 6), g["V6"])       # cv::eigen 6x6
    out = r.selftest_smallmat("QR6", np.concatenate([g["A6"].reshape(-1, 36), g["B6"]], 1))
    assert np.array_equal(out, g["X6"])                                                                         # cv::solve(DECOMP_QR)
    out = r.selftest_smallmat("QR6_WARP", np.concatenate([g["A6"].reshape(-1, 36), g["B6"]], 1))
    assert np.array_equal(out, g["X6"])                               # the warp-parallel form the LM kernel runs: same bits
    sing = np.concatenate([np.zeros((3, 36), np.float32), np.ones((3, 6), np.float32)], 1)
    sing[1, :36:7] = 1.0; sing[1, 35] = 1e-9                          # |R[5][5]| below OpenCV's eps: reported singular, x = 0
    sing[2, :36:7] = 1.0; sing[2, 35] = 0.0                           # an all-zero column: NaN out of hal::QR32f, the same NaN here
    xs, xw = r.selftest_smallmat("QR6", sing), r.selftest_smallmat("QR6_WARP", sing)
    assert np.array_equal(xs, xw, equal_nan=True) and not xs[1].any()
    out = r.selftest_smallmat("LU6", g["V6"].reshape(-1, 36))
    assert np.array_equal(out.reshape(-1, 6, 6), g["I6"])                                                       # cv::Mat::inv (LU)
    # Eigen's 5x3 column-pivoted Householder has no library here: device against the oracle's restatement, and against f64 lstsq
    rng = np.random.default_rng(7)
    A = []
    for _ in range(500):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        c = rng.uniform(-30, 30, 3)
        basis = np.linalg.svd(n.reshape(1, 3))[2][1:]
        pts = c + rng.uniform(-0.4, 0.4, (5, 2)) @ basis + rng.normal(0, 0.01, (5, 1)) * n
        A.append(pts.astype(np.float32))
    A = np.array(A)
    out = r.selftest_smallmat("PLANE5X3", A.reshape(-1, 15))
    want = np.array([oracle.colpiv_solve_5x3(a, -np.ones(5, np.float32)) for a in A])
    assert np.array_equal(out, want)
    ref = np.array([np.linalg.lstsq(a.astype(np.float64), -np.ones(5), rcond=None)[0] for a in A])
    assert np.median(np.abs(out - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-3
    # the LDL^T shortcut may only say "not degenerate" when cv::eigen agrees that every eigenvalue is >= 100
    sure = r.selftest_smallmat("NOT_DEGENERATE", g["A6"].reshape(-1, 36))[:, 0] > 0
    assert np.all(g["W6"][sure].min(axis=1) >= 100.0)
    weak = g["A6"].copy(); weak[:, 5, :] *= 1e-3; weak[:, :, 5] *= 1e-3          # one weak direction: eigenvalue far below 100
    sure_w = r.selftest_smallmat("NOT_DEGENERATE", weak.reshape(-1, 36))[:, 0] > 0
    assert not np.any(sure_w & (np.linalg.eigvalsh(weak.astype(np.float64)).min(axis=1) < 100.0))
    r.close()


def test_device_trig_contract_against_libm(fb):
    """sincosf_c on the device (the trig of pcl::getTransformation, imageProjection.cpp:564-570, mapOptmization.h:309; polynomial
    kernels inside |x| <= pi/4, sincos(double) outside) against (float)sin((double)x) / (float)cos((double)x) of this host's libm,
    bit for bit: deskew-sized angles, the whole reduction-free range, both sides of its boundary, pose-sized angles."""
    import math
    r = fb.Registration(synth.params_for(1), max_frames=1, max_map_corner=1024, max_map_surf=1024)
    rng = np.random.default_rng(7)
    b = np.float32(0.78539816)
    edge = np.array([np.nextafter(b, np.float32(0)), b, np.nextafter(b, np.float32(1)), -b, np.nextafter(-b, np.float32(-1)), 0.0, -0.0, 0.3, -0.3,
                     np.nextafter(np.float32(0.3), np.float32(1)), 0.78125, 1e-30, -1e-20, 1e-6], np.float32)
    x = np.concatenate([edge, rng.uniform(-0.05, 0.05, 40000), rng.uniform(-0.7854, 0.7854, 40000), rng.uniform(-3.2, 3.2, 40000),
                        np.ldexp(rng.uniform(0.5, 1.0, 20000), rng.integers(-40, 0, 20000)) * rng.choice([-1.0, 1.0], 20000)]).astype(np.float32)
    got = r.selftest_smallmat("SINCOS", x.reshape(-1, 1))
    want = np.array([[np.float32(math.sin(float(v))), np.float32(math.cos(float(v)))] for v in x], np.float32)
    bad = np.nonzero(np.any(got.view(np.uint32) != want.view(np.uint32), axis=1))[0]
    assert len(bad) == 0, [(float(x[i]), got[i].tolist(), want[i].tolist()) for i in bad[:8]]
    r.close()


# ------------------------------------------------------------------ feature kernel: full-sort fallback and fuzz
@pytest.mark.parametrize("n_scan,horizon", [(4, 4096), (16, 120), (2, 6000)])
def test_feature_extraction_full_sort_fallback(fb, n_scan, horizon):
    """Horizon_SCAN > 3060 (segments longer than one CTA-wide register sort) and < 180 (segments shorter than a warp) take the
    kernel's full-sort path (features.cu, feat_sort_free == false)."""
    fr = synth.make_frame(1, 3, small=(n_scan, horizon, 2000, 8000))
    P = fr["params"]
    ci = oracle.project(P, fr["scan"], fr["imu"], 0)
    want = oracle.extract_features(P, ci)
    r = fb.Registration(P, max_frames=1, max_map_corner=4096, max_map_surf=16384)
    r.set_cloud_info(0, ci)
    r.featureExtra(0, 1)
    r.sync()
    assert np.array_equal(r.get_buffer(0, "CURVATURE"), want["curvature"])
    assert np.array_equal(r.get_buffer(0, "LABEL"), want["label"])
    assert np.array_equal(r.get_buffer(0, "PICKED"), want["picked"])
    assert np.array_equal(r.get_buffer(0, "CORNER_INDEX"), want["corner_index"])
    assert np.array_equal(r.get_buffer(0, "SURF"), want["surface"])
    r.close()


@pytest.mark.parametrize("n_scan,horizon,leaf", [(16, 2048, 0.4), (16, 2048, 0.1), (16, 2048, 0.02), (8, 1024, 0.05), (16, 900, 0.004)])
def test_per_ring_voxel_grid_paths(fb, n_scan, horizon, leaf):
    """The per-ring VoxelGrid of feat_ring (featureExtraction.h:287-292) has three shapes: runs of equal voxel index sorted with
    32-bit keys (the usual ring), points sorted with 32-bit keys (more than 1024 runs, or a ring too small to hold the run
    tables), points sorted with 64-bit keys (voxel range >= 2^(32 - log2 H): tiny leaves) -- plus PCL's "leaf too small" copy.
    All must give the oracle's surface cloud bit for bit."""
    fr = synth.make_frame(1, 11, small=(n_scan, horizon, 2000, 8000))
    P = dict(fr["params"]); P["odometrySurfLeafSize"] = leaf
    ci = oracle.project(P, fr["scan"], fr["imu"], 0)
    want = oracle.extract_features(P, ci)
    r = fb.Registration(P, max_frames=1, max_map_corner=4096, max_map_surf=16384)
    r.set_cloud_info(0, ci)
    r.featureExtra(0, 1)
    r.sync()
    assert np.array_equal(r.get_buffer(0, "RING_SURF_COUNT_DS"), want["ring_surf_count_ds"])
    assert np.array_equal(r.get_buffer(0, "SURF"), want["surface"])
    assert np.array_equal(r.get_buffer(0, "CORNER_INDEX"), want["corner_index"])
    r.close()


def _fuzz_cloud_info(rng, n_scan, horizon):
    """A cloud_info record made to stress the selection loops: ragged column gaps (the 10-column break of the suppression
    reach), plateaus of exactly equal curvature, range steps (occlusion marks that leak across segment boundaries), rings
    with a handful of points and empty rings."""
    start, end, col, rngs = [], [], [], []
    count = 0
    for ring in range(n_scan):
        mode = rng.integers(0, 6)
        if mode == 0:
            cols = np.zeros(0, np.int64)                                         # empty ring
        elif mode == 1:
            cols = np.sort(rng.choice(horizon, rng.integers(1, 14), replace=False))   # a few points only
        else:
            keep = rng.random(horizon) < rng.choice([0.5, 0.9, 0.99])
            for _ in range(rng.integers(0, 6)):                                  # holes wider than the 10-column reach
                a = rng.integers(0, horizon); keep[a:a + rng.integers(5, 40)] = False
            cols = np.flatnonzero(keep)
        n = len(cols)
        base = rng.uniform(3, 30)
        r_ = np.full(n, base, np.float32)
        kind = rng.integers(0, 4)
        if kind == 0:
            r_ += rng.normal(0, 0.01, n).astype(np.float32)
        elif kind == 1:                                                          # quantised ranges: many exact curvature ties
            r_ = (np.round((r_ + rng.normal(0, 0.03, n)) * 8) / 8).astype(np.float32)
        elif kind == 2:                                                          # steps every few dozen points: corners + occlusions
            r_ += (np.cumsum(rng.random(n) < 0.03) % 3).astype(np.float32) * rng.choice([0.2, 0.5, 2.0]) + rng.normal(0, 0.005, n).astype(np.float32)
        else:                                                                    # perfectly flat: curvature exactly 0 everywhere
            pass
        start.append(count - 1 + 5); count += n; end.append(count - 1 - 5)
        col.append(cols); rngs.append(r_)
    col = np.concatenate(col).astype(np.int32) if count else np.zeros(0, np.int32)
    rngs = np.concatenate(rngs).astype(np.float32) if count else np.zeros(0, np.float32)
    ang = col.astype(np.float64) / horizon * 2 * np.pi
    ring_of = np.concatenate([np.full(e - s + 10, i) for i, (s, e) in enumerate(zip(start, end))]) if count else np.zeros(0)
    elev = np.deg2rad(-15 + 2.0 * ring_of)
    cloud = np.stack([rngs * np.cos(elev) * np.cos(ang), rngs * np.cos(elev) * np.sin(ang), rngs * np.sin(elev), rng.uniform(0, 255, count)], 1).astype(np.float32)
    return dict(startRingIndex=np.array(start, np.int32), endRingIndex=np.array(end, np.int32), pointColInd=col, pointRange=rngs,
                cloud_deskewed=cloud, n_valid=count)


@pytest.mark.parametrize("seed", range(24))
def test_feature_selection_fuzz_against_sequential_oracle(fb, seed):
    rng = np.random.default_rng(1000 + seed)
    n_scan = int(rng.choice([4, 16])); horizon = int(rng.choice([200, 450, 900, 1800, 2048]))
    P = synth.params_for(1); P["N_SCAN"] = n_scan; P["Horizon_SCAN"] = horizon
    P["edgeThreshold"] = float(rng.choice([0.1, 1.0])); P["surfThreshold"] = float(rng.choice([0.1, 0.02]))
    ci = _fuzz_cloud_info(rng, n_scan, horizon)
    want = oracle.extract_features(P, ci)
    r = fb.Registration(P, max_frames=1, max_map_corner=1024, max_map_surf=1024)
    r.set_cloud_info(0, ci)
    r.featureExtra(0, 1)
    r.sync()
    assert np.array_equal(r.get_buffer(0, "LABEL"), want["label"])
    assert np.array_equal(r.get_buffer(0, "PICKED"), want["picked"])
    assert np.array_equal(r.get_buffer(0, "CORNER_INDEX"), want["corner_index"])
    assert np.array_equal(r.get_buffer(0, "RING_SURF_COUNT"), want["ring_surf_count"])
    assert np.array_equal(r.get_buffer(0, "SURF"), want["surface"])
    r.close()


# ------------------------------------------------------------------ capacity overruns are cut AND reported
def test_local_map_capacity_overrun_is_flagged_not_written_out_of_bounds(fb):
    """extractCloud's VoxelGrid and registration()'s CropBox write the slot's local map; more voxels / in-box points than
    max_map_corner / max_map_surf must neither spill into the next slot's map nor go unnoticed (FBPR_FLAG_MAP_TRUNCATED)."""
    rng = np.random.default_rng(3)
    P = synth.params_for(1)
    capC, capS = 256, 512
    r = fb.Registration(P, max_frames=2, max_map_corner=capC, max_map_surf=capS, max_keyframe_points=1 << 15)
    sentinel_c = np.full((capC, 4), 7.0, np.float32); sentinel_s = np.full((capS, 4), 9.0, np.float32)
    r.set_local_map(1, sentinel_c, sentinel_s)                     # the NEXT slot's map must survive untouched
    poses = np.zeros((2, 6), np.float32)
    cf = [np.concatenate([rng.uniform(-20, 20, (4000, 3)), np.zeros((4000, 1))], 1).astype(np.float32) for _ in range(2)]
    sf = [np.concatenate([rng.uniform(-20, 20, (9000, 3)), np.zeros((9000, 1))], 1).astype(np.float32) for _ in range(2)]
    mo = oracle.MapOptimization(P)
    counts = mo.extract_cloud(poses, cf, sf, poses[0, 3:])
    assert counts[2] > capC and counts[3] > capS                   # the oracle's (untruncated) map really exceeds the capacity
    r.extractSurroundingKeyFrames(0, poses, cf, sf, poses[0, 3:])
    c = r.get_counts(0)
    assert (c["n_map_corner"], c["n_map_surf"]) == (capC, capS)
    assert np.array_equal(r.get_buffer(0, "MAP_CORNER").reshape(-1, 4), mo.get_cloud(2)[:capC])     # the first `cap` voxels, in key order
    assert np.array_equal(r.get_buffer(0, "MAP_SURF").reshape(-1, 4), mo.get_cloud(3)[:capS])
    assert np.array_equal(r.get_buffer(1, "MAP_CORNER").reshape(-1, 4), sentinel_c)
    assert np.array_equal(r.get_buffer(1, "MAP_SURF").reshape(-1, 4), sentinel_s)
    fr = synth.make_frame(1, 4, small=(16, 600, 3000, 12000))
    ci = oracle.project(fr["params"], fr["scan"], fr["imu"], 0); fe = oracle.extract_features(fr["params"], ci)
    r.set_feature_clouds(0, fe["corner"], fe["surface"])
    r.set_pose(0, fr["guess"])
    r.downsampleCurrentScan(0, 1); r.scan2MapOptimization(0, 1)
    assert r.get_pose(0)[2] & fb.FLAG_MAP_TRUNCATED
    # CropBox of registration(): 3000 + 12000 in-box points against the same small capacities
    T0 = oracle.get_transformation(fr["guess"])
    r.registration(0, fr["map_corner"], fr["map_surf"], T0)
    c = r.get_counts(0)
    assert (c["n_map_corner"], c["n_map_surf"]) == (capC, capS)
    assert r.get_pose(0)[2] & fb.FLAG_MAP_TRUNCATED
    assert np.array_equal(r.get_buffer(1, "MAP_SURF").reshape(-1, 4), sentinel_s)
    # a map that fits clears the flag again
    r.set_local_map(0, fr["map_corner"][:capC], fr["map_surf"][:capS])
    r.set_pose(0, fr["guess"])
    r.scan2MapOptimization(0, 1)
    assert not (r.get_pose(0)[2] & fb.FLAG_MAP_TRUNCATED)
    r.close()


def test_non_finite_points_are_ignored(fb):
    """NaN / Inf raw points are dropped by the projection (the reference refuses non-dense sweeps, imageProjection.cpp:250-254);
    a non-finite map point never becomes a neighbour and does not blow up the map index."""
    fr = synth.make_frame(1, 6, small=(16, 900, 4000, 20000))
    P = fr["params"]
    want = oracle.project(P, fr["scan"], fr["imu"], 0)
    raw = fb.api.pack_raw(fr["scan"])
    bad = np.zeros(3, fb.api.RAW_POINT_DTYPE)
    bad["x"] = [np.nan, np.inf, 5.0]; bad["y"] = [1.0, 2.0, np.nan]; bad["z"] = [0.0, 0.0, 1.0]; bad["ring"] = [3, 4, 5]
    r = fb.Registration(P, max_frames=1, max_map_corner=8192, max_map_surf=32768)
    r.set_raw_scan(0, np.concatenate([bad, raw]))
    r.project(0, 1); r.sync()
    assert r.get_counts(0)["n_valid"] == want["n_valid"]
    assert np.array_equal(r.get_buffer(0, "RANGE"), want["pointRange"])
    assert np.array_equal(r.get_buffer(0, "WINNER_RAW"), want["winner_raw"] + 3)
    mc = fr["map_corner"].copy(); ms = fr["map_surf"].copy()
    ms_bad = np.concatenate([ms, np.array([[np.inf, 0, 0, 0], [np.nan, np.nan, np.nan, 0], [0, -np.inf, 3, 0]], np.float32)])
    ci = oracle.project(P, fr["scan"], fr["imu"], 0); fe = oracle.extract_features(P, ci)
    mo = oracle.MapOptimization(P); mo.set_scan(fe["corner"], fe["surface"]); mo.set_map(mc, ms); mo.downsample()
    pose_w, iters_w, flags_w, _ = mo.scan2map(fr["guess"])
    r.set_feature_clouds(0, fe["corner"], fe["surface"])
    r.set_local_map(0, mc, ms_bad)
    r.set_pose(0, fr["guess"])
    r.downsampleCurrentScan(0, 1); r.scan2MapOptimization(0, 1)
    pose, iters, flags = r.get_pose(0)
    assert (iters, flags) == (iters_w, flags_w) and np.abs(pose - pose_w).max() <= POSE_TOL
    r.close()


def test_guard_zones_and_run_to_run_determinism(fb, monkeypatch):
    """Stand-in for compute-sanitizer (closed on this B200 pool): a handle created with FBPR_GUARD=1 keeps every device array
    between two pattern-filled guard zones.  The whole path runs with the frames in the LAST slot (so a write past a slot's
    capacity lands in a guard), every operator family is exercised, the guards must be intact -- and the same batch run three
    times must give bit-identical buffers (atomic-rank map index, warp-local fixed points and staged rows included: a race
    would show as a differing bit)."""
    monkeypatch.setenv("FBPR_GUARD", "1")
    frames = [synth.make_frame(3, 70 + i, small=(16, 450, 1500, 9000)) for i in range(2)]
    P = frames[0]["params"]
    F = 3
    r = fb.Registration(P, max_frames=F, max_map_corner=2048, max_map_surf=9216, max_keyframe_points=8192)
    snaps = []
    for rep in range(3):
        for k, fr in enumerate(frames):
            s = F - 1 - k                                        # slots 2 and 1
            r.set_raw_scan(s, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"])
            r.set_local_map(s, fr["map_corner"], fr["map_surf"])
            r.set_pose(s, fr["guess"])
        r.run_frames(1, 2)
        res = r.get_results(1, 2)
        snaps.append((res["pose"].copy(), res["iters"].copy(), r.get_buffer(F - 1, "SURF_DS").copy(), r.get_buffer(F - 1, "CORNER_DS").copy(),
                      r.get_buffer(F - 1, "CORNER_INDEX").copy(), r.get_buffer(F - 1, "LABEL").copy()))
    for sn in snaps[1:]:
        for a, b in zip(snaps[0], sn):
            assert np.array_equal(a, b)
    assert np.all(snaps[0][1] > 0)
    fr = frames[0]
    T0 = np.eye(4, dtype=np.float32)[:3].copy(); T0[:, 3] = fr["guess"][3:]
    r.registration(F - 1, fr["map_corner"], fr["map_surf"], T0)                      # CropBox into the last slot
    for k in range(4):
        r.keyframe_push(np.array([0, 0, 0.01 * k, 0.5 * k, 0, 0], np.float32), 0.3 * k, fr["map_corner"][100 * k:100 * k + 90], fr["map_surf"][500 * k:500 * k + 450])
    r.extractSurroundingKeyFramesResident(F - 1, 1.0, 2.0)
    r.sync()
    assert r.check_guards() == 0
    r.close()

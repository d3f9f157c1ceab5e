"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical
seeded synthetic inputs.  Bars (BASELINE.json north_star): voxel keys, selected-feature indices
and kNN index sets bit-exact; per-point residual coefficients within 1e-5 relative; final pose
within 1e-4 m / 1e-4 rad; iteration counts equal."""
import os

import numpy as np
import pytest

import oracle
import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

POSE_TOL_T = 1e-4      # metres
POSE_TOL_R = 1e-4      # radians
RES_REL_TOL = 1e-5     # per-point residual / coefficient, relative


@pytest.fixture(scope="module")
def fb():
    import feature_base_pointcloud_registration_b200 as m
    return m


def _reg(fb, params, **extra):
    kw = dict(max_frames=1, max_map_corner=65536, max_map_surf=262144)
    kw.update(extra)
    return fb.Registration(params, **kw)


# ------------------------------------------------------------------ VoxelGrid
@pytest.mark.parametrize("n,leaf", [(0, 0.2), (1, 0.2), (7, 0.4), (5000, 0.4), (70000, 0.2), (300000, 0.4)])
def test_voxel_grid_bit_exact(fb, n, leaf):
    rng = np.random.default_rng(n + 1)
    pts = np.concatenate([rng.uniform(-30, 30, (n, 3)), rng.uniform(0, 255, (n, 1))], 1).astype(np.float32)
    if n > 100:
        pts[: n // 2, 2] = rng.normal(0, 0.02, n // 2)            # dense floor: long runs per voxel
        pts[n // 2: n // 2 + 50] = pts[:50]                       # exact duplicates
    r = _reg(fb, synth.params_for(1))
    got = r.voxel_grid(pts, leaf)
    want = oracle.voxel_grid(pts, leaf)
    assert np.array_equal(got["point_keys"], want["point_keys"])
    assert np.array_equal(got["out_keys"], want["out_keys"])
    assert np.array_equal(got["points"], want["points"])          # in-index-order f32 sums: bit-exact


def test_voxel_grid_overflow_path(fb):
    far = np.array([[0, 0, 0, 1], [1e6, 1e6, 1e6, 2], [5, 5, 5, 3]], np.float32)
    r = _reg(fb, synth.params_for(1))
    got = r.voxel_grid(far, 0.01)
    assert np.array_equal(got["points"], far)                     # PCL copies input to output


# ------------------------------------------------------------------ exact 5-NN on the grid index
@pytest.mark.parametrize("first_radius", [1, 2, 5])
@pytest.mark.parametrize("cell", [0.25, 0.5, 1.0])
def test_knn5_index_sets_bit_exact(fb, cell, first_radius):
    fr = synth.make_frame(1, 5)
    m = fr["map_surf"]
    rng = np.random.default_rng(3)
    q = (m[rng.choice(len(m), 20000), :3] + rng.normal(0, 0.08, (20000, 3))).astype(np.float32)
    q[:500] += rng.uniform(-3, 3, (500, 3)).astype(np.float32)    # some far from any surface
    q[500:700] = m[1000:1200, :3]                                 # exactly on map points
    r = _reg(fb, fr["params"])
    idx, d2 = r.knn5(m, q, cell=cell, first_radius=first_radius)
    ridx, rd2 = oracle.knn5(m, q)
    accept = rd2[:, 4] < 1.0
    assert accept.sum() > 15000
    assert np.array_equal(idx[accept], ridx[accept])
    assert np.array_equal(d2[accept], rd2[accept])
    assert np.all(idx[~accept] == -1)


@pytest.mark.parametrize("first_radius", [1, 3])
def test_knn5_ties_and_tiny_maps(fb, first_radius):
    rng = np.random.default_rng(4)
    m = np.concatenate([rng.uniform(-2, 2, (3000, 3)), np.zeros((3000, 1))], 1).astype(np.float32)
    m[1500:1700] = m[100:300]                                     # duplicates -> distance ties broken by index
    q = m[100:300, :3].copy()
    r = _reg(fb, synth.params_for(1))
    idx, d2 = r.knn5(m, q, cell=0.25, first_radius=first_radius)
    ridx, rd2 = oracle.knn5(m, q, brute=True)
    assert np.array_equal(idx, ridx) and np.array_equal(d2, rd2)
    idx, _ = r.knn5(m[:4], q[:10], cell=0.25, first_radius=first_radius)                     # fewer than 5 map points: reject
    assert np.all(idx == -1)


@pytest.mark.parametrize("cell", [0.3, 0.5])
def test_knn5_against_committed_flann_vectors(fb, cell):
    """The DEVICE search against the outputs of a real FLANN KDTreeSingleIndex (tests/golden/flann_cv2.npz, made with cv2.flann by
    tests/golden/make_golden.py) -- no oracle in between: index order and squared distances bit for bit inside the 1 m gate."""
    g = np.load(os.path.join(GOLDEN, "flann_cv2.npz"))
    r = _reg(fb, synth.params_for(1))
    for name in ("map_corner", "map_surf"):
        idx, d2 = r.knn5(g[name], g[name + "_q"], cell=cell, first_radius=1)
        fi, fd = g[name + "_idx"], g[name + "_d2"]
        accept = fd[:, 4] < 1.0
        distinct = np.all(np.diff(fd, axis=1) > 0, axis=1)
        assert accept.sum() > 1000
        assert np.array_equal(d2[accept].view(np.uint32), fd[accept].view(np.uint32))
        assert np.array_equal(idx[accept & distinct], fi[accept & distinct])
        assert np.all(idx[~accept] == -1)
    m = np.concatenate([g["lattice"], np.zeros((len(g["lattice"]), 1), np.float32)], 1)
    idx, d2 = r.knn5(m, g["lattice_q"], cell=cell, first_radius=1)
    assert np.array_equal(d2, g["lattice_d2"])                    # ties: FLANN's order follows its traversal, the distances are pinned


# ------------------------------------------------------------------ projection
@pytest.mark.parametrize("config", [1, 3])
def test_projection_matches_oracle(fb, config):
    fr = synth.make_frame(config, 2)
    P = fr["params"]
    want = oracle.project(P, fr["scan"], fr["imu"], fr["imu_available"])
    r = _reg(fb, P)
    r.set_raw_scan(0, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"])
    r.project(0, 1)
    r.sync()
    assert r.get_counts(0)["n_valid"] == want["n_valid"]
    assert np.array_equal(r.get_buffer(0, "START_RING"), want["startRingIndex"])
    assert np.array_equal(r.get_buffer(0, "END_RING"), want["endRingIndex"])
    assert np.array_equal(r.get_buffer(0, "COL_IND"), want["pointColInd"])
    assert np.array_equal(r.get_buffer(0, "WINNER_RAW"), want["winner_raw"])       # first hit wins
    assert np.array_equal(r.get_buffer(0, "RANGE"), want["pointRange"])
    cloud = r.get_buffer(0, "CLOUD")
    if fr["imu_available"]:
        assert np.allclose(cloud, want["cloud_deskewed"], rtol=0, atol=2e-5)       # device f64 trig vs glibc: <= 1 ulp of f32 trig
        assert np.mean(cloud == want["cloud_deskewed"]) > 0.99
    else:
        assert np.array_equal(cloud, want["cloud_deskewed"])


def test_projection_unsorted_imu_ramp_uses_the_reference_scan(fb):
    """findRotation (imageProjection.cpp:494-526) walks the IMU ramp linearly; the device bisects only when the stamps
    ascend.  A ramp with two stamps swapped (and one repeated) must still give the linear-scan answer."""
    fr = synth.make_frame(3, 5, small=(16, 900, 2000, 8000))
    P = fr["params"]
    imu = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in fr["imu"].items()}
    n = int(imu["imuPointerCur"])
    i = n // 2
    imu["imuTime"][[i, i + 1]] = imu["imuTime"][[i + 1, i]]
    imu["imuTime"][i + 5] = imu["imuTime"][i + 4]
    want = oracle.project(P, fr["scan"], imu, 1)
    r = _reg(fb, P)
    r.set_raw_scan(0, fb.api.pack_raw(fr["scan"]), imu=imu, imu_available=1)
    r.project(0, 1); r.sync()
    cloud = r.get_buffer(0, "CLOUD")
    assert np.array_equal(r.get_buffer(0, "COL_IND"), want["pointColInd"])
    assert np.allclose(cloud, want["cloud_deskewed"], rtol=0, atol=2e-5) and np.mean(cloud == want["cloud_deskewed"]) > 0.99
    r.close()


# ------------------------------------------------------------------ features
@pytest.mark.parametrize("config,frame", [(1, 0), (1, 1), (2, 0), (3, 0), (0, 0)])
def test_feature_extraction_bit_exact(fb, config, frame):
    fr = synth.make_frame(config, frame)
    P = fr["params"]
    ci = oracle.project(P, fr["scan"], fr["imu"], 0)              # no deskew: identical bytes on both sides
    want = oracle.extract_features(P, ci)
    r = _reg(fb, P)
    r.set_cloud_info(0, ci)
    r.featureExtra(0, 1)
    r.sync()
    assert np.array_equal(r.get_buffer(0, "CURVATURE"), want["curvature"])
    assert np.array_equal(r.get_buffer(0, "LABEL"), want["label"])
    assert np.array_equal(r.get_buffer(0, "PICKED"), want["picked"])
    assert np.array_equal(r.get_buffer(0, "CORNER_INDEX"), want["corner_index"])   # order: ring, segment, pick order
    assert np.array_equal(r.get_buffer(0, "CORNER"), want["corner"])
    assert np.array_equal(r.get_buffer(0, "RING_SURF_COUNT"), want["ring_surf_count"])
    assert np.array_equal(r.get_buffer(0, "RING_SURF_COUNT_DS"), want["ring_surf_count_ds"])
    assert np.array_equal(r.get_buffer(0, "SURF"), want["surface"])               # per-ring VoxelGrid centroids


def test_feature_extraction_ragged_and_empty(fb):
    fr = synth.make_frame(1, 7, small=(16, 600, 2000, 8000))
    P = fr["params"]
    scan = dict(fr["scan"])
    keep = ~np.isin(scan["ring"], [3, 4])                         # two empty rings
    keep &= ~((scan["ring"] == 7) & (np.arange(scan["n"]) % 40 != 0))   # one ring with ~15 points
    for k in ("x", "y", "z", "intensity", "ring", "time"):
        scan[k] = scan[k][:scan["n"]][keep]
    scan["n"] = int(keep.sum())
    ci = oracle.project(P, scan, fr["imu"], 0)
    want = oracle.extract_features(P, ci)
    r = _reg(fb, P)
    r.set_cloud_info(0, ci)
    r.featureExtra(0, 1)
    r.sync()
    assert np.array_equal(r.get_buffer(0, "LABEL"), want["label"])
    assert np.array_equal(r.get_buffer(0, "CORNER_INDEX"), want["corner_index"])
    assert np.array_equal(r.get_buffer(0, "SURF"), want["surface"])
    # completely empty frame
    empty = dict(startRingIndex=np.full(16, 4, np.int32), endRingIndex=np.full(16, -6, np.int32),
                 pointColInd=np.zeros(0, np.int32), pointRange=np.zeros(0, np.float32), cloud_deskewed=np.zeros((0, 4), np.float32))
    r.set_cloud_info(0, empty)
    r.featureExtra(0, 1)
    r.sync()
    c = r.get_counts(0)
    assert c["n_corner"] == 0 and c["n_surf"] == 0


# ------------------------------------------------------------------ scan-to-map
def _oracle_scan2map(fr, debug_iter=0):
    P = fr["params"]
    ci = oracle.project(P, fr["scan"], fr["imu"], 0)
    fe = oracle.extract_features(P, ci)
    mo = oracle.MapOptimization(P)
    mo.set_scan(fe["corner"], fe["surface"])
    mo.set_map(fr["map_corner"], fr["map_surf"])
    mo.downsample()
    pose, iters, flags, _ = mo.scan2map(fr["guess"], debug_iter=debug_iter)
    return fe, mo, pose, iters, flags


@pytest.mark.parametrize("config,frame,debug_iter", [(1, 0, 0), (1, 1, 2), (2, 0, 1), (3, 1, 0), (3, 2, 3), (4, 5, 4), (0, 0, 0)])
def test_scan2map_matches_oracle(fb, config, frame, debug_iter):
    fr = synth.make_frame(config, frame)
    fe, mo, pose_w, iters_w, flags_w = _oracle_scan2map(fr, debug_iter)
    dbg = mo.debug()
    r = _reg(fb, fr["params"])
    r.set_feature_clouds(0, fe["corner"], fe["surface"])
    r.set_local_map(0, fr["map_corner"], fr["map_surf"])
    r.set_pose(0, fr["guess"])
    r.set_debug_iteration(debug_iter)
    r.downsampleCurrentScan(0, 1)
    r.scan2MapOptimization(0, 1)
    r.sync()
    # VoxelGrid of the current scan: bit-exact
    assert np.array_equal(r.get_buffer(0, "CORNER_DS"), mo.get_cloud(0))
    assert np.array_equal(r.get_buffer(0, "SURF_DS"), mo.get_cloud(1))
    pose, iters, flags = r.get_pose(0)
    if dbg["iter"] == debug_iter:
        for kind, K in (("CORNER", "corner"), ("SURF", "surf")):
            knn = r.get_buffer(0, "KNN_" + kind); d2 = r.get_buffer(0, "KNN_D2_" + kind)
            accept = dbg[K + "D2"][:, 4] < 1.0
            assert np.array_equal(knn[accept], dbg[K + "Knn"][accept]), kind      # kNN index sets bit-exact
            assert np.array_equal(d2[accept], dbg[K + "D2"][accept])
            flag = r.get_buffer(0, "FLAG_" + kind)
            assert np.array_equal(flag, dbg[K + "Flag"]), kind
            sel = flag.astype(bool)
            co = r.get_buffer(0, "COEFF_" + kind)[sel]; cw = dbg[K + "Coeff"][sel]
            err = np.abs(co - cw) / np.maximum(np.abs(cw), 1e-3)
            assert err.max() <= RES_REL_TOL, (kind, err.max())
        assert np.allclose(r.get_buffer(0, "ATA"), dbg["AtA"], rtol=1e-6, atol=0)
        assert np.allclose(r.get_buffer(0, "ATB"), dbg["AtB"], rtol=1e-5, atol=1e-6)
    assert iters == iters_w                                          # iteration-count parity
    assert flags == flags_w
    assert np.max(np.abs(pose[3:] - pose_w[3:])) <= POSE_TOL_T
    assert np.max(np.abs(pose[:3] - pose_w[:3])) <= POSE_TOL_R


def test_degenerate_corridor_reproduces_matP_quirk(fb):
    fr = synth.make_frame(0, 0)
    fe, mo, pose_w, iters_w, flags_w = _oracle_scan2map(fr, 0)
    assert flags_w & oracle.FLAG_DEGENERATE and iters_w == 2          # zero step at iteration 1 (mapOptmization.h:1278)
    r = _reg(fb, fr["params"])
    r.set_feature_clouds(0, fe["corner"], fe["surface"])
    r.set_local_map(0, fr["map_corner"], fr["map_surf"])
    r.set_pose(0, fr["guess"])
    r.downsampleCurrentScan(0, 1); r.scan2MapOptimization(0, 1)
    pose, iters, flags = r.get_pose(0)
    assert (iters, flags) == (iters_w, flags_w)
    assert np.allclose(pose, pose_w, atol=1e-4)


def test_not_enough_features_and_too_few_correspondences(fb):
    fr = synth.make_frame(1, 2, small=(16, 600, 2000, 8000))
    P = fr["params"]
    ci = oracle.project(P, fr["scan"], fr["imu"], 0)
    fe = oracle.extract_features(P, ci)
    r = _reg(fb, P)
    # (a) gate: <= edgeFeatureMinValidNum corners -> pose unchanged, flag set
    r.set_feature_clouds(0, fe["corner"][:5], fe["surface"])
    r.set_local_map(0, fr["map_corner"], fr["map_surf"])
    r.set_pose(0, fr["guess"])
    r.downsampleCurrentScan(0, 1); r.scan2MapOptimization(0, 1)
    pose, iters, flags = r.get_pose(0)
    assert flags == fb.FLAG_NOT_ENOUGH_FEATURES and iters == 0 and np.array_equal(pose, fr["guess"])
    # (b) map far away: no neighbours within 1 m -> < 50 rows -> 30 idle iterations, pose unchanged
    far_c = fr["map_corner"].copy(); far_c[:, :3] += 500
    far_s = fr["map_surf"].copy(); far_s[:, :3] += 500
    mo = oracle.MapOptimization(P)
    mo.set_scan(fe["corner"], fe["surface"]); mo.set_map(far_c, far_s); mo.downsample()
    pose_w, iters_w, flags_w, _ = mo.scan2map(fr["guess"])
    r.set_feature_clouds(0, fe["corner"], fe["surface"])
    r.set_local_map(0, far_c, far_s)
    r.set_pose(0, fr["guess"])
    r.downsampleCurrentScan(0, 1); r.scan2MapOptimization(0, 1)
    pose, iters, flags = r.get_pose(0)
    assert (iters, flags) == (iters_w, flags_w) == (30, fb.FLAG_TOO_FEW_CORRESPONDENCES)
    assert np.array_equal(pose, pose_w)


def test_transform_update_imu_slerp_and_clamps(fb):
    P = dict(synth.params_for(1)); P["z_tollerance"] = 0.5; P["rotation_tollerance"] = 0.3
    mo = oracle.MapOptimization(P)
    mo.set_imu(1, 0.12, -0.07)
    pose = np.array([0.05, 0.02, 1.0, 3.0, 4.0, 0.9], np.float32)
    want = mo.transform_update(pose)
    r = _reg(fb, P)
    ci = dict(startRingIndex=np.full(16, 4, np.int32), endRingIndex=np.full(16, -6, np.int32), pointColInd=np.zeros(0, np.int32),
              pointRange=np.zeros(0, np.float32), cloud_deskewed=np.zeros((0, 4), np.float32))
    r.set_cloud_info(0, ci, imu_available=1, imu_roll_init=0.12, imu_pitch_init=-0.07)
    r.set_pose(0, pose)
    r.transformUpdate(0, 1)
    got, _, _ = r.get_pose(0)
    assert np.allclose(got, want, atol=1e-6)
    assert got[5] == np.float32(0.5)


# ------------------------------------------------------------------ whole path, single frame and batch
@pytest.mark.parametrize("config", [1, 3])
def test_end_to_end_pose_and_batch_equals_single(fb, config):
    frames = [synth.make_frame(config, f) for f in range(3)]
    P = frames[0]["params"]
    want = []
    for fr in frames:
        ci = oracle.project(P, fr["scan"], fr["imu"], fr["imu_available"])
        fe = oracle.extract_features(P, ci)
        mo = oracle.MapOptimization(P)
        mo.set_imu(fr["imu_available"], 0.0, 0.0)                    # transformUpdate slerps towards the IMU attitude
        mo.set_scan(fe["corner"], fe["surface"]); mo.set_map(fr["map_corner"], fr["map_surf"]); mo.downsample()
        want.append(mo.scan2map(fr["guess"])[:3])
    r = _reg(fb, P, max_frames=3)
    for s, fr in enumerate(frames):
        r.set_raw_scan(s, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"])
        r.set_local_map(s, fr["map_corner"], fr["map_surf"])
    r.set_poses(0, np.stack([fr["guess"] for fr in frames]))
    r.run_frames(0, 3)                                               # batched: all three frames in one set of launches
    res = r.get_results(0, 3)
    for s in range(3):
        pw, iw, fw = want[s]
        assert int(res[s]["iters"]) == iw and int(res[s]["flags"]) == fw
        assert np.max(np.abs(res[s]["pose"][3:] - pw[3:])) <= POSE_TOL_T
        assert np.max(np.abs(res[s]["pose"][:3] - pw[:3])) <= POSE_TOL_R
    # the same frames one at a time (and through CUDA graphs) give identical bytes
    r.use_graphs(True)
    for rep in range(2):
        for s in range(3):
            r.set_pose(s, frames[s]["guess"])
            r.run_frames(s, 1)
        again = r.get_results(0, 3)
        assert np.array_equal(again["pose"], res["pose"]) and np.array_equal(again["iters"], res["iters"])


def test_pipelined_register_frames_equals_plain_batch(fb):
    """fbpr_register_frames (chunked uploads on a second stream) must give exactly what set_frames + run_frames gives."""
    F = 5
    frames = [synth.make_frame(1, 20 + i) for i in range(F)]
    r = fb.Registration(frames[0]["params"], max_frames=F, max_map_corner=16384, max_map_surf=65536)
    raws = [fb.api.pack_raw(fr["scan"]) for fr in frames]
    fin = r.make_frame_inputs([dict(raw_ptr=raw.ctypes.data, n_raw=len(raw), imu=fr["imu"], imu_available=fr["imu_available"],
                                    map_corner_ptr=fr["map_corner"].ctypes.data, n_map_corner=len(fr["map_corner"]),
                                    map_surf_ptr=fr["map_surf"].ctypes.data, n_map_surf=len(fr["map_surf"]), pose=fr["guess"])
                               for fr, raw in zip(frames, raws)])
    r.set_frames(0, fin); r.run_frames(0, F)
    want = r.get_results(0, F)
    for chunk in (0, 1, 2, 32):
        got = r.register_frames(0, fin, chunk)
        assert np.array_equal(got["pose"], want["pose"]) and np.array_equal(got["iters"], want["iters"]) and np.array_equal(got["flags"], want["flags"])
    # the same frames out of ONE pinned arena (all sweeps back to back, then all maps): a densely packed group of buffers is
    # uploaded as one copy + a scatter kernel, i.e. 2 extra kernel launches per chunk and identical results
    import torch
    arrays = list(raws) + [m for fr in frames for m in (fr["map_corner"], fr["map_surf"])]
    offs = np.concatenate([[0], np.cumsum([(a_.nbytes + 15) // 16 * 16 for a_ in arrays])])
    arena = torch.empty(int(offs[-1]) + 16, dtype=torch.uint8).pin_memory()
    ptrs = []
    for a_, off in zip(arrays, offs[:-1]):
        arena[int(off):int(off) + a_.nbytes] = torch.from_numpy(np.ascontiguousarray(a_).view(np.uint8).reshape(-1))
        ptrs.append(arena.data_ptr() + int(off))
    fin_a = r.make_frame_inputs([dict(raw_ptr=ptrs[i], n_raw=len(raw), imu=fr["imu"], imu_available=fr["imu_available"],
                                      map_corner_ptr=ptrs[F + 2 * i], n_map_corner=len(fr["map_corner"]),
                                      map_surf_ptr=ptrs[F + 2 * i + 1], n_map_surf=len(fr["map_surf"]), pose=fr["guess"])
                                 for i, (fr, raw) in enumerate(zip(frames, raws))])
    for chunk, nchunks in ((5, 1), (2, 3)):
        l0 = r.kernel_launches(); r.register_frames(0, fin, chunk); l1 = r.kernel_launches()
        got = r.register_frames(0, fin_a, chunk); l2 = r.kernel_launches()
        assert np.array_equal(got["pose"], want["pose"]) and np.array_equal(got["iters"], want["iters"]) and np.array_equal(got["flags"], want["flags"])
        merged = (l2 - l1) - (l1 - l0)
        assert merged == 2 * nchunks - (1 if chunk == 2 else 0), merged      # the last chunk of (2, 2, 1) has a single sweep: plain copy
    # split form, two batches in flight on disjoint slot ranges (double buffering), three rounds
    r2 = fb.Registration(frames[0]["params"], max_frames=2 * F, max_map_corner=16384, max_map_surf=65536)
    fin2 = r2.make_frame_inputs([dict(raw_ptr=raw.ctypes.data, n_raw=len(raw), imu=fr["imu"], imu_available=fr["imu_available"],
                                      map_corner_ptr=fr["map_corner"].ctypes.data, n_map_corner=len(fr["map_corner"]),
                                      map_surf_ptr=fr["map_surf"].ctypes.data, n_map_surf=len(fr["map_surf"]), pose=fr["guess"])
                                 for fr, raw in zip(frames, raws)])
    t = r2.register_frames_begin(0, fin2, 2)
    for rnd in range(3):
        t_next = r2.register_frames_begin(F * ((rnd + 1) % 2), fin2, 0)
        got = r2.register_frames_end(t)
        assert np.array_equal(got["pose"], want["pose"]) and np.array_equal(got["iters"], want["iters"]) and np.array_equal(got["flags"], want["flags"])
        t = t_next
    with pytest.raises(fb.FbprError, match="overlaps"):
        r2.register_frames_begin(F * (3 % 2), fin2, 0)           # the same slots while that batch is still in flight
    got = r2.register_frames_end(t)
    assert np.array_equal(got["pose"], want["pose"])
    with pytest.raises(fb.FbprError, match="no such batch"):
        r2.lib.fbpr_register_frames_end.restype = int
        r2._ck(r2.lib.fbpr_register_frames_end(r2.h, 0, None))
    r2.set_frames(0, fin2); r2.run_frames(0, F)                   # ordinary operators work again after _end
    assert np.array_equal(r2.get_results(0, F)["iters"], want["iters"])
    r2.close()
    for s_, fr in enumerate(frames[:2]):                      # and both equal the oracle
        ci = oracle.project(fr["params"], fr["scan"], fr["imu"], fr["imu_available"]); fe = oracle.extract_features(fr["params"], ci)
        mo = oracle.MapOptimization(fr["params"]); mo.set_imu(fr["imu_available"], 0.0, 0.0)
        mo.set_scan(fe["corner"], fe["surface"]); mo.set_map(fr["map_corner"], fr["map_surf"]); mo.downsample()
        pw, iw, fw, _ = mo.scan2map(fr["guess"])
        assert iw == int(want[s_]["iters"]) and fw == int(want[s_]["flags"])
        assert np.max(np.abs(pw - want[s_]["pose"])) <= 1e-4
    r.close()


# ------------------------------------------------------------------ operators around the path
def _registration_case():
    fr = synth.make_frame(1, 4, small=(16, 600, 3000, 12000))
    P = fr["params"]
    ci = oracle.project(P, fr["scan"], fr["imu"], 0)
    fe = oracle.extract_features(P, ci)
    rng = np.random.default_rng(0)
    clutter = np.concatenate([rng.uniform(100, 200, (500, 3)), np.zeros((500, 1))], 1).astype(np.float32)
    gc = np.concatenate([fr["map_corner"], clutter]); gs = np.concatenate([clutter, fr["map_surf"]])
    T0 = oracle.get_transformation(fr["guess"])
    mo = oracle.MapOptimization(P)
    mo.set_scan(fe["corner"], fe["surface"])
    T, iters, flags = mo.registration(gc, gs, T0)
    return fr, P, ci, fe, gc, gs, T0, T, iters, flags


def test_registration_entry_cropbox_matches_oracle(fb):
    """fbpr_registration = mapOptimization::registration (mapOptmization.h:263-343): CropBox, pose decompose, downsample, LM, recompose."""
    fr, P, ci, fe, gc, gs, T0, T, iters, flags = _registration_case()
    r = _reg(fb, P)
    r.set_feature_clouds(0, fe["corner"], fe["surface"])
    got = r.registration(0, gc, gs, T0)
    pose, it, fl = r.get_pose(0)
    assert (it, fl) == (iters, flags)
    assert np.max(np.abs(got[:, 3] - T[:, 3])) <= POSE_TOL_T and np.max(np.abs(got[:, :3] - T[:, :3])) <= POSE_TOL_R
    c = r.get_counts(0)
    keep_c = oracle.crop_box(gc, T0[:, 3] - [30, 30, 10], T0[:, 3] + [30, 30, 10])
    keep_s = oracle.crop_box(gs, T0[:, 3] - [30, 30, 10], T0[:, 3] + [30, 30, 10])
    assert c["n_map_corner"] == len(keep_c) and c["n_map_surf"] == len(keep_s)
    assert np.array_equal(r.get_buffer(0, "MAP_CORNER").reshape(-1, 4), keep_c)      # inclusive AABB, order preserved
    # resident global map: same answer with NULL maps
    r.set_global_map(gc, gs)
    r.set_feature_clouds(0, fe["corner"], fe["surface"])
    got2 = r.registration(0, None, None, T0)
    assert np.array_equal(got2, got)
    r.close()


def test_extract_surrounding_keyframes_matches_oracle(fb):
    """extractCloud (mapOptmization.h:909-955): distance re-check, per-keyframe rigid transform, concat in list order, VoxelGrid x2."""
    rng = np.random.default_rng(1)
    P = synth.params_for(1)
    K = 4
    poses = np.concatenate([rng.uniform(-0.2, 0.2, (K, 3)), rng.uniform(-5, 5, (K, 3))], 1).astype(np.float32)
    poses[3, 3:] += 100.0                                        # farther than surroundingKeyframeSearchRadius from the last key pose
    cf = [np.concatenate([rng.uniform(-10, 10, (n, 3)), np.full((n, 1), k)], 1).astype(np.float32) for k, n in enumerate((50, 70, 30, 40))]
    sf = [np.concatenate([rng.uniform(-10, 10, (n, 3)), np.full((n, 1), k)], 1).astype(np.float32) for k, n in enumerate((500, 700, 300, 400))]
    mo = oracle.MapOptimization(P)
    counts = mo.extract_cloud(poses, cf, sf, poses[0, 3:])
    r = _reg(fb, P, max_keyframe_points=4096)
    r.extractSurroundingKeyFrames(0, poses, cf, sf, poses[0, 3:])
    c = r.get_counts(0)
    assert (c["n_map_corner"], c["n_map_surf"]) == (int(counts[2]), int(counts[3]))
    assert np.array_equal(r.get_buffer(0, "MAP_CORNER").reshape(-1, 4), mo.get_cloud(2))
    assert np.array_equal(r.get_buffer(0, "MAP_SURF").reshape(-1, 4), mo.get_cloud(3))
    r.close()


@pytest.mark.parametrize("share", [0, 1])
def test_cpp_host_classes_cloud_handler(fb, share, tmp_path):
    """The C++ host layer (host/feature_matching.hpp: FeatureExtraction::featureExtra + mapOptimization::registration with the
    reference's names) driven the way ImageProjection::cloudHandler drives the reference (imageProjection.cpp:203, :218)."""
    import ctypes as C
    import os
    fr, P, ci, fe, gc, gs, T0, T, iters, flags = _registration_case()
    so = os.path.join(os.path.dirname(fb.__file__), "host", "libfeature_matching_b200.so")
    lib = C.CDLL(so)
    y = tmp_path / "params.yaml"
    y.write_text("".join(f"{k}: {P[k]}\n" for k in ("edgeThreshold", "surfThreshold", "edgeFeatureMinValidNum", "surfFeatureMinValidNum",
                                                      "odometrySurfLeafSize", "mappingCornerLeafSize", "mappingSurfLeafSize",
                                                      "z_tollerance", "rotation_tollerance", "numberOfCores")))
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    sr, er, col, rng_, cloud = i32(ci["startRingIndex"]), i32(ci["endRingIndex"]), i32(ci["pointColInd"]), f32(ci["pointRange"]), f32(ci["cloud_deskewed"])
    gcf, gsf = f32(gc), f32(gs)
    pose = f32(T0).reshape(-1).copy()
    it = C.c_int(0); fl = C.c_uint(0); counts = (C.c_int * 4)(); err = C.create_string_buffer(512)
    rc = lib.fm_cloud_handler(str(y).encode(), int(P["N_SCAN"]), int(P["Horizon_SCAN"]), vp(sr), vp(er), vp(col), vp(rng_), vp(cloud), len(col),
                              vp(gcf), len(gcf), vp(gsf), len(gsf), vp(pose), int(share), C.byref(it), C.byref(fl), counts, err, 512)
    assert rc == 0, err.value.decode()
    assert (it.value, fl.value) == (iters, flags)
    got = pose.reshape(3, 4)
    assert np.max(np.abs(got[:, 3] - T[:, 3])) <= POSE_TOL_T and np.max(np.abs(got[:, :3] - T[:, :3])) <= POSE_TOL_R


# ------------------------------------------------------------------ BASELINE configs[4]: 128-beam scan vs 2 M-point map
def test_dense_map_stress_config5(fb):
    """OS1-128-like 128 x 2048 scan against a raw-dense 2 M-point map (0.2 m surface leaf): feature indices, the scan
    VoxelGrids, the kNN sets / coefficients of iterations 0 and 3 and the final pose against the oracle."""
    fr = synth.make_frame(5, 0)
    P = fr["params"]
    cap = dict(max_map_corner=len(fr["map_corner"]) + 64, max_map_surf=len(fr["map_surf"]) + 64)
    ci = oracle.project(P, fr["scan"], fr["imu"], fr["imu_available"])
    fe = oracle.extract_features(P, ci)
    r = _reg(fb, P, **cap)
    r.set_raw_scan(0, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"])
    r.project(0, 1); r.featureExtra(0, 1); r.sync()
    assert np.array_equal(r.get_buffer(0, "CORNER_INDEX"), fe["corner_index"])
    assert np.array_equal(r.get_buffer(0, "SURF"), fe["surface"])
    for debug_iter in (0, 3):
        mo = oracle.MapOptimization(P)
        mo.set_imu(fr["imu_available"], 0.0, 0.0)
        mo.set_scan(fe["corner"], fe["surface"]); mo.set_map(fr["map_corner"], fr["map_surf"]); mo.downsample()
        pose_w, iters_w, flags_w, _ = mo.scan2map(fr["guess"], debug_iter)
        dbg = mo.debug()
        r.set_local_map(0, fr["map_corner"], fr["map_surf"])
        r.set_pose(0, fr["guess"])
        r.set_debug_iteration(debug_iter)
        r.downsampleCurrentScan(0, 1); r.scan2MapOptimization(0, 1); r.sync()
        assert np.array_equal(r.get_buffer(0, "SURF_DS"), mo.get_cloud(1))
        pose, iters, flags = r.get_pose(0)
        assert (iters, flags) == (iters_w, flags_w)
        assert np.max(np.abs(pose[3:] - pose_w[3:])) <= POSE_TOL_T and np.max(np.abs(pose[:3] - pose_w[:3])) <= POSE_TOL_R
        if dbg["iter"] == debug_iter:
            for kind, K in (("CORNER", "corner"), ("SURF", "surf")):
                knn = r.get_buffer(0, "KNN_" + kind); d2 = r.get_buffer(0, "KNN_D2_" + kind)
                accept = dbg[K + "D2"][:, 4] < 1.0
                assert np.array_equal(knn[accept], dbg[K + "Knn"][accept]), kind
                assert np.array_equal(d2[accept], dbg[K + "D2"][accept])
                assert np.array_equal(r.get_buffer(0, "FLAG_" + kind), dbg[K + "Flag"]), kind
    r.close()


# ------------------------------------------------------------------ keyframe selection, wire formats
def _keyframe_store(rng, n, spread, npts=(40, 300)):
    poses = np.concatenate([rng.uniform(-0.2, 0.2, (n, 3)), np.cumsum(rng.uniform(0, spread, (n, 3)) * [1, 0.3, 0.02], 0)], 1).astype(np.float32)
    cf = [np.concatenate([rng.uniform(-10, 10, (npts[0] + k, 3)), np.full((npts[0] + k, 1), k)], 1).astype(np.float32) for k in range(n)]
    sf = [np.concatenate([rng.uniform(-10, 10, (npts[1] + 3 * k, 3)), np.full((npts[1] + 3 * k, 1), k)], 1).astype(np.float32) for k in range(n)]
    return poses, cf, sf


def test_extract_cloud_with_averaged_key_poses_matches_oracle(fb):
    """fbpr_extract_cloud with the re-check positions of extractNearby's VoxelGrid-averaged key poses (mapOptmization.h:924-927)."""
    rng = np.random.default_rng(5)
    P = dict(synth.params_for(1)); P["surroundingKeyframeSearchRadius"] = 12.0
    n = 40
    poses, cf, sf = _keyframe_store(rng, n, 1.2)
    times = np.arange(n) * 0.7
    mo = oracle.MapOptimization(P)
    ds, counts = mo.extract_surrounding(poses, times, 2.0, times[-1] + 0.1, cf, sf)
    sel = ds[:, 3].astype(np.int32)                              # (int)intensity (:927)
    assert np.any(ds[:, 3] != sel)                               # averaged indices really occur
    r = _reg(fb, P, max_keyframe_points=1 << 16, max_map_corner=1 << 15, max_map_surf=1 << 16)
    r.extractCloud(0, poses[sel], ds[:, :3], [cf[i] for i in sel], [sf[i] for i in sel], poses[-1, 3:])
    c = r.get_counts(0)
    assert (c["n_map_corner"], c["n_map_surf"]) == (int(counts[2]), int(counts[3]))
    assert np.array_equal(r.get_buffer(0, "MAP_CORNER").reshape(-1, 4), mo.get_cloud(2))
    assert np.array_equal(r.get_buffer(0, "MAP_SURF").reshape(-1, 4), mo.get_cloud(3))
    r.close()


def test_cpp_host_extract_nearby_matches_oracle(fb, tmp_path):
    """mapOptimization::extractSurroundingKeyFrames of the C++ host layer with its own extractNearby (radius search, key-pose
    VoxelGrid on the device, last-10 s append) against the oracle's restatement of mapOptmization.h:872-955."""
    import ctypes as C
    import os
    rng = np.random.default_rng(9)
    P = dict(synth.params_for(1)); P["surroundingKeyframeSearchRadius"] = 15.0
    n = 60
    poses, cf, sf = _keyframe_store(rng, n, 1.0)
    times = np.arange(n) * 0.5
    tlast = float(times[-1] + 0.05)
    mo = oracle.MapOptimization(P)
    ds_w, counts_w = mo.extract_surrounding(poses, times, 2.0, tlast, cf, sf)
    y = tmp_path / "params.yaml"
    y.write_text("".join(f"{k}: {P[k]}\n" for k in ("mappingCornerLeafSize", "mappingSurfLeafSize", "surroundingKeyframeSearchRadius")) + "surroundingKeyframeDensity: 2.0\n")
    lib = C.CDLL(os.path.join(os.path.dirname(fb.__file__), "host", "libfeature_matching_b200.so"))
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    coff = np.zeros(n + 1, np.int32); soff = np.zeros(n + 1, np.int32)
    coff[1:] = np.cumsum([len(c) for c in cf]); soff[1:] = np.cumsum([len(s) for s in sf])
    call, sall = f32(np.concatenate(cf)), f32(np.concatenate(sf))
    ds = np.zeros((4 * n, 4), np.float32); nds = C.c_int(0); counts = (C.c_int * 2)(); err = C.create_string_buffer(512)
    capC, capS = int(coff[-1]) * 2 + 64, int(soff[-1]) * 2 + 64
    mc = np.zeros((capC, 4), np.float32); ms = np.zeros((capS, 4), np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    kt = np.ascontiguousarray(times, np.float64); kp = f32(poses)
    rc = lib.fm_extract_surrounding(str(y).encode(), int(P["N_SCAN"]), int(P["Horizon_SCAN"]), vp(kp), vp(kt), n, C.c_double(tlast),
                                    vp(call), vp(coff), vp(sall), vp(soff), vp(ds), len(ds), C.byref(nds), vp(mc), capC, vp(ms), capS, counts, err, 512, 0, -1)
    assert rc == 0, err.value.decode()
    assert nds.value == len(ds_w) and np.array_equal(ds[:nds.value], ds_w)
    assert (counts[0], counts[1]) == (int(counts_w[2]), int(counts_w[3]))
    assert np.array_equal(mc[:counts[0]], mo.get_cloud(2)) and np.array_equal(ms[:counts[1]], mo.get_cloud(3))
    # loopClosureEnableFlag: extractForLoopClosure (mapOptmization.h:857-870) with surroundingKeyframeSize = 7
    ds_w, counts_w = mo.extract_surrounding(poses, times, 2.0, tlast, cf, sf, loop_closure=True, keyframe_size=7)
    assert len(ds_w) == 8
    rc = lib.fm_extract_surrounding(str(y).encode(), int(P["N_SCAN"]), int(P["Horizon_SCAN"]), vp(kp), vp(kt), n, C.c_double(tlast),
                                    vp(call), vp(coff), vp(sall), vp(soff), vp(ds), len(ds), C.byref(nds), vp(mc), capC, vp(ms), capS, counts, err, 512, 1, 7)
    assert rc == 0, err.value.decode()
    assert nds.value == len(ds_w) and np.array_equal(ds[:nds.value], ds_w)
    assert np.array_equal(mc[:counts[0]], mo.get_cloud(2)) and np.array_equal(ms[:counts[1]], mo.get_cloud(3))


@pytest.mark.parametrize("mode", ["nearby", "loop_closure", "nearby_all_recent"])
def test_resident_keyframe_store_matches_oracle(fb, mode):
    """fbpr_extract_surrounding_keyframes_resident: extractSurroundingKeyFrames (mapOptmization.h:964-978) end to end on the device
    over >= 1000 resident key poses -- radius search sorted by (d^2, index), key-pose VoxelGrid with averaged indices, last-10-s
    rule / extractForLoopClosure, distance re-check, transform + concat + VoxelGrid x2 -- against the oracle: selection list and
    both local maps bit-exact.  The store grows past its first capacity and has its poses corrected in place (correctPoses)."""
    rng = np.random.default_rng(21)
    P = dict(synth.params_for(1)); P["surroundingKeyframeSearchRadius"] = 18.0
    n = 1300
    # a trajectory that winds back on itself, so the radius holds old and new key poses and voxels average far-apart indices
    t = np.linspace(0, 6 * np.pi, n)
    xyz = np.stack([30 * np.cos(t) + 3 * np.sin(7 * t), 30 * np.sin(t) + 2 * np.cos(5 * t), 0.3 * np.sin(3 * t)], 1)
    poses = np.concatenate([rng.uniform(-0.2, 0.2, (n, 3)), xyz + rng.normal(0, 0.3, (n, 3))], 1).astype(np.float32)
    cf = [np.concatenate([rng.uniform(-10, 10, (20 + k % 7, 3)), np.full((20 + k % 7, 1), k)], 1).astype(np.float32) for k in range(n)]
    sf = [np.concatenate([rng.uniform(-10, 10, (60 + k % 11, 3)), np.full((60 + k % 11, 1), k)], 1).astype(np.float32) for k in range(n)]
    times = np.arange(n) * (0.002 if mode == "nearby_all_recent" else 0.4)        # all_recent: every key pose is younger than 10 s
    tlast = float(times[-1] + 0.05)
    lc = mode == "loop_closure"
    mo = oracle.MapOptimization(P)
    r = _reg(fb, P, max_keyframe_points=1 << 18, max_map_corner=1 << 16, max_map_surf=1 << 17)
    wrong = poses.copy(); wrong[:500, 3:] += 1.0
    for k in range(n):                                           # pushed with stale poses for the first 500, corrected below
        assert r.keyframe_push(wrong[k], times[k], cf[k], sf[k]) == k
        if k == 399:                                             # an extraction in the middle of the sequence, before the store grows
            ds_w, counts_w = mo.extract_surrounding(wrong[:400], times[:400], 2.0, float(times[399] + 0.05), cf[:400], sf[:400], loop_closure=lc, keyframe_size=25)
            r.extractSurroundingKeyFramesResident(0, float(times[399] + 0.05), 2.0, loop_closure=lc, keyframe_size=25)
            lst, idx = r.keyframe_selection()
            assert np.array_equal(lst, ds_w)
            assert np.array_equal(r.get_buffer(0, "MAP_SURF").reshape(-1, 4), mo.get_cloud(3))
    assert r.keyframes_count() == n
    r.keyframes_set_poses(0, poses[:500])
    ds_w, counts_w = mo.extract_surrounding(poses, times, 2.0, tlast, cf, sf, loop_closure=lc, keyframe_size=25)
    r.extractSurroundingKeyFramesResident(0, tlast, 2.0, loop_closure=lc, keyframe_size=25)
    lst, idx = r.keyframe_selection()
    assert len(lst) == len(ds_w) and np.array_equal(lst, ds_w)
    if mode == "nearby":
        assert np.any(lst[:, 3] != np.floor(lst[:, 3])) and len(lst) > 60          # averaged indices and a real radius result
    if lc:
        assert len(lst) == 26
    keep = idx >= 0
    assert np.array_equal(idx[keep], lst[keep, 3].astype(np.int32))
    c = r.get_counts(0)
    assert (c["n_map_corner"], c["n_map_surf"]) == (int(counts_w[2]), int(counts_w[3]))
    assert np.array_equal(r.get_buffer(0, "MAP_CORNER").reshape(-1, 4), mo.get_cloud(2))
    assert np.array_equal(r.get_buffer(0, "MAP_SURF").reshape(-1, 4), mo.get_cloud(3))
    _, _, flags = r.get_pose(0)
    # an empty store leaves the local map alone (:966-967)
    r.keyframes_clear()
    r.extractSurroundingKeyFramesResident(0, tlast, 2.0)
    assert r.get_counts(0)["n_map_surf"] == int(counts_w[3])
    r.close()


@pytest.mark.parametrize("layout", ["velodyne22", "pcl32", "no_time_ring8"])
def test_pointcloud2_wire_format_projection(fb, layout):
    """fbpr_set_raw_scan_pc2: PointCloud2 payload bytes of the layouts cachePointCloud accepts (imageProjection.cpp:252-297)
    -> the same projection as the packed records / the oracle; a cloud without a time field disables deskew (deskewFlag -1)."""
    fr = synth.make_frame(3, 1, small=(16, 900, 2000, 8000))
    P = fr["params"]; sc = fr["scan"]; n = int(sc["n"])
    if layout == "velodyne22":
        dt = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"], "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                       "offsets": [0, 4, 8, 12, 16, 18], "itemsize": 22})
        L = dict(point_step=22, off_x=0, off_y=4, off_z=8, off_intensity=12, off_ring=16, ring_bytes=2, off_time=18)
    elif layout == "pcl32":                                      # pcl::toROSMsg<PointXYZIRT>: EIGEN_ALIGN16 struct, 32 bytes
        dt = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"], "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                       "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})
        L = dict(point_step=32, off_x=0, off_y=4, off_z=8, off_intensity=16, off_ring=20, ring_bytes=2, off_time=24)
    else:
        dt = np.dtype({"names": ["x", "y", "z", "intensity", "ring"], "formats": ["<f4", "<f4", "<f4", "<f4", "u1"],
                       "offsets": [0, 4, 8, 12, 17], "itemsize": 19})
        L = dict(point_step=19, off_x=0, off_y=4, off_z=8, off_intensity=12, off_ring=17, ring_bytes=1, off_time=-1)
    msg = np.zeros(n, dt)
    for k in dt.names:
        msg[k] = sc[k][:n]
    deskew = 1 if L["off_time"] >= 0 else -1
    ci = oracle.project(P, sc, fr["imu"], fr["imu_available"], deskew_flag=deskew)
    r = _reg(fb, P)
    r.set_raw_scan_pc2(0, msg.tobytes(), n, L, imu=fr["imu"], imu_available=fr["imu_available"])
    r.project(0, 1); r.sync()
    assert r.get_counts(0)["n_valid"] == len(ci["pointRange"])
    assert np.array_equal(r.get_buffer(0, "COL_IND"), ci["pointColInd"])
    assert np.array_equal(r.get_buffer(0, "RANGE"), ci["pointRange"])
    assert np.array_equal(r.get_buffer(0, "CLOUD").reshape(-1, 4), ci["cloud_deskewed"])
    # refused layouts: no ring channel (the reference shuts down, :262-280), offsets outside the record
    bad = dict(L); bad["ring_bytes"] = 0
    with pytest.raises(fb.FbprError, match="ring"):
        r.set_raw_scan_pc2(0, msg.tobytes(), n, bad)
    bad = dict(L); bad["off_z"] = L["point_step"] - 2
    with pytest.raises(fb.FbprError, match="point_step"):
        r.set_raw_scan_pc2(0, msg.tobytes(), n, bad)
    r.close()


def test_pcl_xyzi32_wire_roundtrip(fb):
    """32-byte pcl::PointXYZI records (the PointCloud2 payload of cloud_corner / cloud_surface and of a loaded PCD map,
    utility.h:255-264, mapOptmization.h:272-273): device repack in both directions, then the same registration."""
    fr = synth.make_frame(1, 2, small=(16, 900, 4000, 20000))
    P = fr["params"]
    ci = oracle.project(P, fr["scan"], fr["imu"], fr["imu_available"]); fe = oracle.extract_features(P, ci)
    wide = lambda a: np.concatenate([a[:, :3], np.ones((len(a), 1), np.float32), a[:, 3:4], np.zeros((len(a), 3), np.float32)], 1).astype(np.float32)
    r = _reg(fb, P, max_map_corner=8192, max_map_surf=32768)
    r.set_clouds_xyzi32(0, 0, wide(fe["corner"]), wide(fe["surface"]))
    r.set_clouds_xyzi32(0, 1, wide(fr["map_corner"]), wide(fr["map_surf"]))
    assert np.array_equal(r.get_buffer(0, "CORNER").reshape(-1, 4), fe["corner"]) and np.array_equal(r.get_buffer(0, "MAP_SURF").reshape(-1, 4), fr["map_surf"])
    assert np.array_equal(r.get_buffer_xyzi32(0, "SURF"), wide(fe["surface"]))
    assert np.array_equal(r.get_buffer_xyzi32(0, "MAP_CORNER"), wide(fr["map_corner"]))
    r.set_pose(0, fr["guess"])
    r.downsampleCurrentScan(0, 1); r.scan2MapOptimization(0, 1); r.sync()
    mo = oracle.MapOptimization(P)
    mo.set_imu(0, 0.0, 0.0)
    mo.set_scan(fe["corner"], fe["surface"]); mo.set_map(fr["map_corner"], fr["map_surf"]); mo.downsample()
    pose_w, iters_w, flags_w, _ = mo.scan2map(fr["guess"])
    pose, iters, flags = r.get_pose(0)
    assert (iters, flags) == (iters_w, flags_w)
    assert np.max(np.abs(pose[3:] - pose_w[3:])) <= POSE_TOL_T and np.max(np.abs(pose[:3] - pose_w[:3])) <= POSE_TOL_R
    assert np.array_equal(r.get_buffer_xyzi32(0, "SURF_DS")[:, [0, 1, 2, 4]], mo.get_cloud(1))
    r.close()


# ------------------------------------------------------------------ full-size properties that need no oracle (BASELINE configs[3] sizes)
def test_full_size_properties_config4(fb):
    """A 64 x 2048 sweep against its 200 k-point map, checked through properties that hold at any size: VoxelGrid keys strictly
    ascending and covering every input key with counts that add up, centroids inside their voxel; 5-NN distances ascending, equal
    to the f32 distance recomputed from the coordinates, never beaten by a brute-force scan on a sample; idempotent re-registration
    from the converged pose; the final pose within centimetres of the synthetic ground truth."""
    fr = synth.make_frame(4, 7)
    P = fr["params"]
    r = _reg(fb, P, max_map_corner=len(fr["map_corner"]) + 64, max_map_surf=len(fr["map_surf"]) + 64)
    # VoxelGrid of the 160 k surface map at the mapping leaf size
    m = fr["map_surf"]
    leaf = float(P["mappingSurfLeafSize"])
    vg = r.voxel_grid(m, leaf)
    ok_, pk = vg["out_keys"], vg["point_keys"]
    assert np.all(np.diff(ok_) > 0)                                           # ascending, one output per voxel
    uniq, cnt = np.unique(pk, return_counts=True)
    assert np.array_equal(uniq, ok_) and cnt.sum() == len(m)
    inv = np.float32(1.0) / np.float32(leaf)
    cell_of_centroid = np.floor(vg["points"][:, :3] * inv)
    order = np.argsort(pk, kind="stable")
    starts = np.searchsorted(pk[order], ok_, side="left")
    member_cell = np.floor(m[order[starts], :3] * inv)
    assert np.max(np.abs(cell_of_centroid - member_cell)) <= 1.0              # a centroid may round onto the voxel face, never farther
    # exact 5-NN of 4000 perturbed map points
    rng = np.random.default_rng(11)
    q = (m[rng.choice(len(m), 4000), :3] + rng.normal(0, 0.05, (4000, 3))).astype(np.float32)
    idx, d2 = r.knn5(m, q, cell=0.33, first_radius=1)
    acc = idx[:, 0] >= 0
    assert acc.mean() > 0.95
    assert np.all(np.diff(d2[acc], axis=1) >= 0)
    nb = m[idx[acc]][:, :, :3]
    dx = q[acc][:, None, 0] - nb[:, :, 0]; dy = q[acc][:, None, 1] - nb[:, :, 1]; dz = q[acc][:, None, 2] - nb[:, :, 2]
    rec = (dx * dx + dy * dy) + dz * dz                                       # the oracle's f32 expression order
    assert np.array_equal(rec.astype(np.float32), d2[acc])
    for i in np.flatnonzero(acc)[:200]:                                       # brute force on a sample: nothing closer was missed
        d = ((q[i, 0] - m[:, 0]) ** 2 + (q[i, 1] - m[:, 1]) ** 2) + (q[i, 2] - m[:, 2]) ** 2
        best = np.lexsort((np.arange(len(m)), d))[:5]
        assert np.array_equal(best, idx[i]), i
    # whole path, then a second registration started from the converged pose changes nothing measurable
    r.set_raw_scan(0, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"])
    r.set_local_map(0, fr["map_corner"], fr["map_surf"])
    r.set_pose(0, fr["guess"])
    r.run_frames(0, 1)
    pose, iters, flags = r.get_pose(0)
    assert flags & 8 and iters <= 12                                          # FBPR_FLAG_CONVERGED
    assert np.max(np.abs(pose[3:] - fr["gt"][3:])) < 0.05 and np.max(np.abs(pose[:3] - fr["gt"][:3])) < 0.01
    r.scan2MapOptimization(0, 1)
    pose2, iters2, flags2 = r.get_pose(0)
    assert flags2 & 8 and iters2 <= 2 and np.max(np.abs(pose2 - pose)) < 2e-3
    r.close()


def test_wire_format_batched_uploads(fb):
    """fbpr_set_frames / fbpr_register_frames with the 22-byte Velodyne wire records (imageProjection.cpp:8-21) and 12-byte XYZ
    maps: what crosses PCIe is the wire layout, the device repacks it -- projection, features and poses are bit-identical to the
    packed 24-byte / 16-byte upload (a map point's intensity is never read by the registration).  Pageable buffers (one copy per
    piece into the landing area), one pinned arena (one merged copy per chunk) and device-resident sources."""
    import torch
    F = 5
    frames = [synth.make_frame(3, 80 + i, small=(16, 900, 2000, 8000)) for i in range(F)]
    P = frames[0]["params"]
    r = fb.Registration(P, max_frames=F, max_map_corner=4096, max_map_surf=16384)
    raws = [fb.api.pack_raw(fr["scan"]) for fr in frames]
    base = [dict(imu=fr["imu"], imu_available=fr["imu_available"], pose=fr["guess"], n_raw=len(raw),
                 n_map_corner=len(fr["map_corner"]), n_map_surf=len(fr["map_surf"])) for fr, raw in zip(frames, raws)]
    fin = r.make_frame_inputs([dict(b, raw_ptr=raw.ctypes.data, map_corner_ptr=fr["map_corner"].ctypes.data, map_surf_ptr=fr["map_surf"].ctypes.data)
                               for b, fr, raw in zip(base, frames, raws)])
    r.set_frames(0, fin); r.run_frames(0, F)
    want = r.get_results(0, F)
    want_cloud = [r.get_buffer(s, "CLOUD").copy() for s in range(F)]
    want_surf = [r.get_buffer(s, "SURF_DS").copy() for s in range(F)]
    assert np.all(want["iters"] > 0)
    # wire copies: odd record counts make the 22-byte pieces end off any 4-byte boundary
    wires = []
    for raw in raws:
        w = np.zeros(len(raw), fb.api.VELODYNE22_DTYPE)
        for k in ("x", "y", "z", "intensity", "time"):
            w[k] = raw[k]
        w["ring"] = raw["ring"]
        wires.append(w)
    xyzc = [np.ascontiguousarray(fr["map_corner"][:, :3]) for fr in frames]
    xyzs = [np.ascontiguousarray(fr["map_surf"][:, :3]) for fr in frames]
    fmt = dict(raw_format=fb.api.RAW_VELODYNE22, map_format=fb.api.MAP_XYZ12)

    def check(got):
        assert np.array_equal(got["pose"], want["pose"]) and np.array_equal(got["iters"], want["iters"]) and np.array_equal(got["flags"], want["flags"])
        for s in range(F):
            assert np.array_equal(r.get_buffer(s, "CLOUD"), want_cloud[s])
            assert np.array_equal(r.get_buffer(s, "SURF_DS"), want_surf[s])
            m = r.get_buffer(s, "MAP_SURF").reshape(-1, 4)
            assert np.array_equal(m[:, :3], xyzs[s]) and not m[:, 3].any()

    # (1) pageable host buffers
    fin_w = r.make_frame_inputs([dict(b, raw_ptr=w.ctypes.data, map_corner_ptr=c.ctypes.data, map_surf_ptr=s_.ctypes.data, **fmt)
                                 for b, w, c, s_ in zip(base, wires, xyzc, xyzs)])
    r.set_frames(0, fin_w); r.run_frames(0, F)
    check(r.get_results(0, F))
    for chunk in (0, 2):
        check(r.register_frames(0, fin_w, chunk))
    # mixed formats inside one batch
    fin_m = r.make_frame_inputs([dict(b, raw_ptr=(w if i % 2 else raw).ctypes.data, map_corner_ptr=c.ctypes.data, map_surf_ptr=s_.ctypes.data,
                                      raw_format=fb.api.RAW_VELODYNE22 if i % 2 else fb.api.RAW_PACKED24, map_format=fb.api.MAP_XYZ12)
                                 for i, (b, w, raw, c, s_) in enumerate(zip(base, wires, raws, xyzc, xyzs))])
    check(r.register_frames(0, fin_m, 2))
    # (2) one pinned arena, pieces 2-byte aligned back to back
    arrays = list(wires) + [m for c, s_ in zip(xyzc, xyzs) for m in (c, s_)]
    offs, o = [], 0
    for a_ in arrays:
        o = (o + 3) // 4 * 4 if a_.dtype == np.float32 else (o + 1) // 2 * 2
        offs.append(o); o += a_.nbytes
    arena = torch.empty(o + 16, dtype=torch.uint8).pin_memory()
    for a_, off in zip(arrays, offs):
        arena[off:off + a_.nbytes] = torch.from_numpy(np.ascontiguousarray(a_).view(np.uint8).reshape(-1))
    ptrs = [arena.data_ptr() + off for off in offs]
    fin_a = r.make_frame_inputs([dict(b, raw_ptr=ptrs[i], map_corner_ptr=ptrs[F + 2 * i], map_surf_ptr=ptrs[F + 2 * i + 1], **fmt) for i, b in enumerate(base)])
    for chunk in (5, 2):
        check(r.register_frames(0, fin_a, chunk))
    # (3) device-resident sources
    dev = torch.empty(o + 16, dtype=torch.uint8, device="cuda"); dev.copy_(arena); torch.cuda.synchronize()
    dptrs = [dev.data_ptr() + off for off in offs]
    fin_d = r.make_frame_inputs([dict(b, raw_ptr=dptrs[i], map_corner_ptr=dptrs[F + 2 * i], map_surf_ptr=dptrs[F + 2 * i + 1], **fmt) for i, b in enumerate(base)])
    r.set_frames(0, fin_d, mem=fb.api.MEM_DEVICE); r.run_frames(0, F)
    check(r.get_results(0, F))
    with pytest.raises(fb.FbprError, match="raw_format"):
        r.set_frames(0, r.make_frame_inputs([dict(base[0], raw_ptr=raws[0].ctypes.data, raw_format=7)]))
    r.close()


def test_batched_registration_against_resident_global_map(fb):
    """map_format = FROM_GLOBAL: no local map crosses PCIe; every frame of a batch crops the resident global maps around its own
    pose guess on the device (the fork's live registration(), mapOptmization.h:284-304, for many independent frames) and registers
    against that crop.  Cropped maps bit-exact vs the oracle's CropBox, iterations / flags equal, poses within tolerance; the
    pipelined call gives the same as set_frames + run_frames."""
    F = 4
    frames = [synth.make_frame(3, 90 + i, small=(16, 900, 3000, 12000)) for i in range(F)]
    P = frames[0]["params"]
    # the global maps: two frames' maps of the same scene, plus far-away points that every crop must drop
    far = np.array([[500.0, 0, 0, 1], [0, -500.0, 0, 2], [20, 15, 40.0, 3]], np.float32)
    gc = np.concatenate([frames[0]["map_corner"], far, frames[1]["map_corner"]]).astype(np.float32)
    gs = np.concatenate([frames[0]["map_surf"], far, frames[1]["map_surf"]]).astype(np.float32)
    r = fb.Registration(P, max_frames=2 * F, max_map_corner=len(gc) + 64, max_map_surf=len(gs) + 64)
    r.set_global_map(gc, gs)
    raws = [fb.api.pack_wire22(fr["scan"]) for fr in frames]
    fin = r.make_frame_inputs([dict(raw_ptr=raw.ctypes.data, n_raw=len(raw), imu=fr["imu"], imu_available=fr["imu_available"], pose=fr["guess"],
                                    raw_format=fb.api.RAW_VELODYNE22, map_format=fb.api.MAP_FROM_GLOBAL) for fr, raw in zip(frames, raws)])
    r.set_frames(F, fin); r.run_frames(F, F)
    got = r.get_results(F, F)
    for i, fr in enumerate(frames):
        g = fr["guess"]
        keep_c = oracle.crop_box(gc, g[3:] - np.float32([30, 30, 10]), g[3:] + np.float32([30, 30, 10]))
        keep_s = oracle.crop_box(gs, g[3:] - np.float32([30, 30, 10]), g[3:] + np.float32([30, 30, 10]))
        assert len(keep_c) < len(gc) and len(keep_s) < len(gs)
        assert np.array_equal(r.get_buffer(F + i, "MAP_CORNER").reshape(-1, 4), keep_c)
        assert np.array_equal(r.get_buffer(F + i, "MAP_SURF").reshape(-1, 4), keep_s)
        ci = oracle.project(P, fr["scan"], fr["imu"], fr["imu_available"])
        fe = oracle.extract_features(P, ci)
        mo = oracle.MapOptimization(P); mo.set_imu(fr["imu_available"], 0.0, 0.0)
        mo.set_scan(fe["corner"], fe["surface"]); mo.set_map(keep_c, keep_s); mo.downsample()
        pose_w, iters_w, flags_w, _ = mo.scan2map(g)
        assert (int(got["iters"][i]), int(got["flags"][i])) == (iters_w, flags_w)
        assert np.max(np.abs(got["pose"][i][3:] - pose_w[3:])) <= POSE_TOL_T and np.max(np.abs(got["pose"][i][:3] - pose_w[:3])) <= POSE_TOL_R
    for chunk in (0, 3):
        again = r.register_frames(0, fin, chunk)
        assert np.array_equal(again["pose"], got["pose"]) and np.array_equal(again["iters"], got["iters"])
    # a mixed batch is refused
    mixed = r.make_frame_inputs([dict(raw_ptr=raws[0].ctypes.data, n_raw=len(raws[0]), pose=frames[0]["guess"], raw_format=fb.api.RAW_VELODYNE22, map_format=fb.api.MAP_FROM_GLOBAL),
                                 dict(raw_ptr=raws[1].ctypes.data, n_raw=len(raws[1]), pose=frames[1]["guess"], raw_format=fb.api.RAW_VELODYNE22,
                                      map_corner_ptr=gc.ctypes.data, n_map_corner=len(gc), map_surf_ptr=gs.ctypes.data, n_map_surf=len(gs))])
    with pytest.raises(fb.FbprError, match="FROM_GLOBAL"):
        r.set_frames(0, mixed)
    r.close()


def test_run_frames_pipelined_equals_run_frames(fb):
    """fbpr_run_frames_pipelined (front-end, map index and LM loop of consecutive batches on three streams) must give exactly what
    fbpr_run_frames gives batch by batch, and later work on the handle's stream must see the finished slots."""
    F = 7
    frames = [synth.make_frame(3, 120 + i, small=(16, 900, 2000, 8000)) for i in range(F)]
    r = fb.Registration(frames[0]["params"], max_frames=F, max_map_corner=4096, max_map_surf=16384)
    guesses = np.stack([fr["guess"] for fr in frames])
    for s, fr in enumerate(frames):
        r.set_raw_scan(s, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"])
        r.set_local_map(s, fr["map_corner"], fr["map_surf"])
    r.set_poses(0, guesses)
    r.run_frames(0, F)
    want = r.get_results(0, F)
    want_ds = [r.get_buffer(s, "SURF_DS").copy() for s in range(F)]
    assert np.all(want["iters"] > 0)
    for batch in (3, 1, 0, 7):
        r.set_poses(0, guesses)
        r.run_frames_pipelined(0, F, batch)
        got = r.get_results(0, F)                                # queued on the handle's stream right behind the pipelined call
        assert np.array_equal(got["pose"], want["pose"]) and np.array_equal(got["iters"], want["iters"]) and np.array_equal(got["flags"], want["flags"])
        for s in range(F):
            assert np.array_equal(r.get_buffer(s, "SURF_DS"), want_ds[s])
    r.set_poses(0, guesses); r.run_frames_pipelined(2, 4, 2)     # a sub-range
    assert np.array_equal(r.get_results(2, 4)["pose"], want["pose"][2:6])
    r.close()

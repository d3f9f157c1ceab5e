"""Diagnostic run on the GPU box: compares every stage against the oracle and prints mismatch
statistics instead of stopping at the first failure.  Not a test; see tests/test_gpu_parity.py."""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle, synth
import feature_base_pointcloud_registration_b200 as fb


def cmp(name, got, want, exact=True):
    got = np.asarray(got); want = np.asarray(want)
    if got.shape != want.shape:
        print(f"  [{name}] SHAPE got {got.shape} want {want.shape}"); return False
    if got.size == 0:
        print(f"  [{name}] ok (empty)"); return True
    neq = got != want
    if neq.any():
        idx = np.argwhere(neq)
        d = np.abs(got.astype(np.float64) - want.astype(np.float64))
        print(f"  [{name}] MISMATCH {neq.sum()}/{neq.size} first at {idx[0].tolist()} got {got[tuple(idx[0])]} want {want[tuple(idx[0])]} maxabs {d.max():.3e}")
        return False
    print(f"  [{name}] ok ({got.size})"); return True


def stage(title):
    print(f"== {title}", flush=True)


def run(config, frame):
    fr = synth.make_frame(config, frame)
    P = fr["params"]
    stage(f"config {config} frame {frame}: projection")
    want = oracle.project(P, fr["scan"], fr["imu"], fr["imu_available"])
    r = fb.Registration(P, max_frames=2, max_map_corner=max(65536, len(fr["map_corner"]) + 16), max_map_surf=max(262144, len(fr["map_surf"]) + 16))
    r.set_raw_scan(0, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"])
    r.project(0, 1); r.sync()
    print("  counts", r.get_counts(0), "want n_valid", want["n_valid"])
    cmp("startRing", r.get_buffer(0, "START_RING"), want["startRingIndex"])
    cmp("endRing", r.get_buffer(0, "END_RING"), want["endRingIndex"])
    cmp("colInd", r.get_buffer(0, "COL_IND"), want["pointColInd"])
    cmp("winner", r.get_buffer(0, "WINNER_RAW"), want["winner_raw"])
    cmp("range", r.get_buffer(0, "RANGE"), want["pointRange"])
    cmp("cloud", r.get_buffer(0, "CLOUD"), want["cloud_deskewed"])

    stage("features (on oracle cloud_info)")
    fe = oracle.extract_features(P, want)
    r.set_cloud_info(0, want)
    r.featureExtra(0, 1); r.sync()
    print("  counts", r.get_counts(0), "want corners", len(fe["corner"]), "surf", len(fe["surface"]))
    cmp("curvature", r.get_buffer(0, "CURVATURE"), fe["curvature"])
    cmp("picked", r.get_buffer(0, "PICKED"), fe["picked"])
    cmp("label", r.get_buffer(0, "LABEL"), fe["label"])
    cmp("cornerIndex", r.get_buffer(0, "CORNER_INDEX"), fe["corner_index"])
    cmp("ringSurf", r.get_buffer(0, "RING_SURF_COUNT"), fe["ring_surf_count"])
    cmp("ringSurfDS", r.get_buffer(0, "RING_SURF_COUNT_DS"), fe["ring_surf_count_ds"])
    cmp("surface", r.get_buffer(0, "SURF"), fe["surface"])

    stage("downsample + scan2map")
    mo = oracle.MapOptimization(P)
    mo.set_scan(fe["corner"], fe["surface"]); mo.set_map(fr["map_corner"], fr["map_surf"]); mo.downsample()
    t0 = time.time(); pose_w, iters_w, flags_w, secs = mo.scan2map(fr["guess"], debug_iter=0); t1 = time.time()
    dbg = mo.debug()
    r.set_feature_clouds(0, fe["corner"], fe["surface"])
    r.set_local_map(0, fr["map_corner"], fr["map_surf"])
    r.set_pose(0, fr["guess"]); r.set_debug_iteration(0)
    r.downsampleCurrentScan(0, 1); r.sync()
    cmp("cornerDS", r.get_buffer(0, "CORNER_DS"), mo.get_cloud(0))
    cmp("surfDS", r.get_buffer(0, "SURF_DS"), mo.get_cloud(1))
    t2 = time.time(); r.scan2MapOptimization(0, 1); r.sync(); t3 = time.time()
    pose, iters, flags = r.get_pose(0)
    print(f"  oracle: iters {iters_w} flags {flags_w} pose {pose_w}  ({(t1-t0)*1e3:.1f} ms)")
    print(f"  gpu   : iters {iters} flags {flags} pose {pose}  ({(t3-t2)*1e3:.1f} ms incl. launch)")
    print(f"  |dpose| t {np.abs(pose[3:]-pose_w[3:]).max():.2e} r {np.abs(pose[:3]-pose_w[:3]).max():.2e}")
    for kind, K in (("CORNER", "corner"), ("SURF", "surf")):
        knn = r.get_buffer(0, "KNN_" + kind); d2 = r.get_buffer(0, "KNN_D2_" + kind)
        acc = dbg[K + "D2"][:, 4] < 1.0
        cmp(kind + " knn(accepted)", knn[acc], dbg[K + "Knn"][acc])
        cmp(kind + " d2(accepted)", d2[acc], dbg[K + "D2"][acc])
        print(f"  {kind} rejected agree: {np.all(knn[~acc] == -1)} ({(~acc).sum()} rejected)")
        flag = r.get_buffer(0, "FLAG_" + kind)
        cmp(kind + " flag", flag, dbg[K + "Flag"])
        sel = flag.astype(bool) & dbg[K + "Flag"].astype(bool)
        co = r.get_buffer(0, "COEFF_" + kind)[sel]; cw = dbg[K + "Coeff"][sel]
        if len(co):
            err = np.abs(co - cw) / np.maximum(np.abs(cw), 1e-3)
            print(f"  {kind} coeff rel err max {err.max():.2e} exact {np.mean(co == cw):.4f}")
    cmp("AtA", r.get_buffer(0, "ATA"), dbg["AtA"])
    cmp("AtB", r.get_buffer(0, "ATB"), dbg["AtB"])
    cmp("X", r.get_buffer(0, "X"), dbg["X"])
    tr = r.get_buffer(0, "POSE_TRACE")[:iters]; trw = mo.pose_trace()
    if len(tr) == len(trw):
        print("  pose trace max diff per iter", np.abs(tr - trw).max(1))

    stage("whole path through run_frames (projection -> features -> downsample -> LM)")
    r.set_raw_scan(1, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"])
    r.set_local_map(1, fr["map_corner"], fr["map_surf"]); r.set_pose(1, fr["guess"])
    r.run_frames(1, 1); r.sync()
    pose2, it2, fl2 = r.get_pose(1)
    print(f"  e2e: iters {it2} flags {fl2} pose {pose2} counts {r.get_counts(1)}")
    print(f"  |dpose| vs oracle t {np.abs(pose2[3:]-pose_w[3:]).max():.2e} r {np.abs(pose2[:3]-pose_w[:3]).max():.2e}")
    # timing, warm
    for graphs in (False, True):
        r.use_graphs(graphs)
        for rep in range(3):
            r.set_pose(1, fr["guess"]); r.sync()
            t = time.time(); r.run_frames(1, 1); r.sync(); dt = time.time() - t
        print(f"  run_frames wall (graphs={graphs}): {dt*1e3:.3f} ms, launches so far {r.kernel_launches()}")
    r.close()


if __name__ == "__main__":
    cfgs = [int(c) for c in sys.argv[1:]] or [1, 3]
    for c in cfgs:
        try:
            run(c, 0)
        except Exception:
            traceback.print_exc()

import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth
import feature_base_pointcloud_registration_b200 as fb
F = int(sys.argv[1]); cl = 8; cell = float(sys.argv[2])
frames = [synth.make_frame(4, i) for i in range(F)]
r = fb.Registration(frames[0]["params"], max_frames=F, max_map_corner=40064, max_map_surf=160064, lm_cluster_size=cl, knn_cell_surf=cell)
for s, fr in enumerate(frames):
    r.set_raw_scan(s, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"])
    r.set_local_map(s, fr["map_corner"], fr["map_surf"])
for rep in range(3):
    r.set_poses(0, np.stack([fr["guess"] for fr in frames])); r.run_frames(0, F); r.sync()
ctas = 148 if F == 1 else cl
buf = np.zeros(ctas * 512 * 8, np.int64)
r.lib.fbpr_debug_lm_profile(r.h, buf.ctypes.data_as(C.c_void_p), ctas)
v = buf.reshape(ctas, 512, 8)[:, ::32, :].reshape(-1, 8).astype(np.float64)   # lane 0 of each warp
nq = v[:, 7].sum(); passes = v[:, 5].sum()
names = ["bounds", "scan", "cand-load-issue", "cand-consume", "merge", "passes", "pre(pOri+T)", "queries"]
print("F", F, "cell", cell, "queries", nq, "passes/query %.2f" % (passes / nq))
for i in (6, 0, 1, 2, 3, 4):
    print("  %-16s %.0f cycles/query" % (names[i], v[:, i].sum() / nq))
print("  total %.0f cycles/query" % (v[:, [0, 1, 2, 3, 4, 6]].sum() / nq))

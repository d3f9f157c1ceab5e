import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth
import feature_base_pointcloud_registration_b200 as fb
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cl = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cell = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
frames = [synth.make_frame(4, i) for i in range(F)]
r = fb.Registration(frames[0]["params"], max_frames=F, max_map_corner=40064, max_map_surf=160064, lm_cluster_size=cl, knn_cell_surf=cell)
for s, fr in enumerate(frames):
    r.set_raw_scan(s, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"])
    r.set_local_map(s, fr["map_corner"], fr["map_surf"])
r.enable_stage_timing(True)
for rep in range(3):
    r.set_poses(0, np.stack([fr["guess"] for fr in frames])); r.run_frames(0, F); r.sync()
ctas = 148 if F == 1 else cl
buf = np.zeros(ctas * 512 * 8, np.int64)
r.lib.fbpr_debug_lm_profile(r.h, buf.ctypes.data_as(C.c_void_p), ctas)
v = buf.reshape(ctas, 512, 8) / 1.9e3      # us at ~1.9 GHz
names = ["trig+T", "phaseA", "syncA", "phaseB", "reduce", "barrier", "sum", "solve+sync"]
print("stage ms (3 reps):", {k: round(v[0], 3) for k, v in r.get_stage_ms().items()})
print("F", F, "cell", cell, "ctas", ctas, "iters", r.get_pose(0)[1], "counts", r.get_counts(0))
print("thread 0 of CTA 0 (us, all iterations):", dict(zip(names, np.round(v[0, 0], 1))))
print("max over threads:", dict(zip(names, np.round(v.max((0, 1)), 1))))
print("mean over threads:", dict(zip(names, np.round(v.mean((0, 1)), 1))))
print("total per thread mean %.1f us" % v.sum(2).mean())

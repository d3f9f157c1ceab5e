"""Small driver for ncu: one config-3 frame (or a batch) through the whole path a few times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth
import feature_base_pointcloud_registration_b200 as fb

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cluster = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cell = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
frames = [synth.make_frame(4, i) for i in range(F)]
extra = dict(max_frames=F, max_map_corner=40064, max_map_surf=160064)
if cluster:
    extra["lm_cluster_size"] = cluster
if cell:
    extra["knn_cell_surf"] = cell
r = fb.Registration(frames[0]["params"], **extra)
for s, fr in enumerate(frames):
    r.set_raw_scan(s, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"])
    r.set_local_map(s, fr["map_corner"], fr["map_surf"])
for rep in range(reps):
    r.set_poses(0, np.stack([fr["guess"] for fr in frames]))
    r.enable_stage_timing(True)
    r.run_frames(0, F)
    r.sync()
    print(rep, {k: round(v[0], 3) for k, v in r.get_stage_ms().items()}, r.get_results(0, F)["iters"][:8])

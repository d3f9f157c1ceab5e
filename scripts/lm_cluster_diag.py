"""LM stage time per cluster size for a batch of F config-4 frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth
import feature_base_pointcloud_registration_b200 as fb

F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
sizes = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 4, 6, 7, 8, 9, 10, 12, 16]
frames = [synth.make_frame(4, i) for i in range(F)]
guess = np.stack([fr["guess"] for fr in frames])
for c in sizes:
    r = fb.Registration(frames[0]["params"], max_frames=F, max_map_corner=40064, max_map_surf=160064, lm_cluster_size=c)
    for s, fr in enumerate(frames):
        r.set_raw_scan(s, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"])
        r.set_local_map(s, fr["map_corner"], fr["map_surf"])
    best = None
    for rep in range(4):
        r.set_poses(0, guess); r.enable_stage_timing(True); r.run_frames(0, F); r.sync()
        ms = r.get_stage_ms()
        if rep > 0: best = ms["lm"][0] if best is None else min(best, ms["lm"][0])
    print("cluster", c, "lm ms", round(best, 3), "iters", r.get_results(0, F)["iters"][:6], flush=True)
    r.close()

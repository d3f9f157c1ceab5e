"""map_index + LM stage time for F config-4 frames over kNN grid cell sizes / first search radius."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth
import feature_base_pointcloud_registration_b200 as fb

F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
frames = [synth.make_frame(4, i) for i in range(F)]
guess = np.stack([fr["guess"] for fr in frames])
raws = [fb.api.pack_raw(fr["scan"]) for fr in frames]
combos = [(0.25, 0.5, 0.5)] + [(cs, cc, fr0) for cs in (0.28, 0.3, 0.33) for cc in (0.4, 0.5) for fr0 in (0.25, 0.35, 0.45)]
if len(sys.argv) > 2:
    combos = [tuple(float(x) for x in c.split(",")) for c in sys.argv[2:]]
ref = None
for cs, cc, fr0 in combos:
    r = fb.Registration(frames[0]["params"], max_frames=F, max_map_corner=40064, max_map_surf=160064, knn_cell_surf=cs, knn_cell_corner=cc, knn_first_radius=fr0)
    for s, fr in enumerate(frames):
        r.set_raw_scan(s, raws[s], imu=fr["imu"], imu_available=fr["imu_available"])
        r.set_local_map(s, fr["map_corner"], fr["map_surf"])
    best = None
    for rep in range(4):
        r.set_poses(0, guess); r.enable_stage_timing(True); r.run_frames(0, F); r.sync()
        ms = r.get_stage_ms()
        if rep > 0:
            v = (ms["map_index"][0], ms["lm"][0])
            best = v if best is None or sum(v) < sum(best) else best
    res = r.get_results(0, F)
    if ref is None: ref = res
    same = np.array_equal(res["pose"], ref["pose"]) and np.array_equal(res["iters"], ref["iters"])
    print(f"cell_surf {cs} cell_corner {cc} first_radius {fr0}: map_index {best[0]:.3f} lm {best[1]:.3f} sum {sum(best):.3f} {'same' if same else 'DIFFERENT'}", flush=True)
    r.close()

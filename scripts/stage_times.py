"""F config-4 frames through the whole path: per-stage CUDA-event times.
python scripts/lm_tile_diag.py [F] [reps] [cluster]      (FBPR_LM_NO_MORTON=1 switches the spatial ordering of the LM kernel off)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_base_pointcloud_registration_b200 as fb  # noqa: E402
import synth  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cluster = int(sys.argv[3]) if len(sys.argv) > 3 else 0
frames = [synth.make_frame(4, i) for i in range(F)]
if os.environ.get("FBPR_MAP_ORDER") == "voxel":
    # the order pcl::VoxelGrid leaves a local map in (mapOptmization.h:948-954: ascending voxel index, x fastest) instead of the
    # random order synth.c draws the map points in
    for fr in frames:
        for name, leaf in (("map_corner", 0.2), ("map_surf", 0.4)):
            m = fr[name]; ijk = np.floor(m[:, :3] / np.float32(leaf)).astype(np.int64); ijk -= ijk.min(0)
            d = ijk.max(0) + 1
            fr[name] = np.ascontiguousarray(m[np.argsort(ijk[:, 0] + d[0] * (ijk[:, 1] + d[1] * ijk[:, 2]), kind="stable")])
cfg = synth.CONFIGS[4]
r = fb.Registration(frames[0]["params"], max_frames=F, max_map_corner=cfg["map_corner"] + 64, max_map_surf=cfg["map_surf"] + 64, lm_cluster_size=cluster)
raws = [fb.api.pack_raw(fr["scan"]) for fr in frames]
fin = r.make_frame_inputs([dict(raw_ptr=raw.ctypes.data, n_raw=len(raw), imu=fr["imu"], imu_available=fr["imu_available"],
                                map_corner_ptr=fr["map_corner"].ctypes.data, n_map_corner=len(fr["map_corner"]),
                                map_surf_ptr=fr["map_surf"].ctypes.data, n_map_surf=len(fr["map_surf"]), pose=fr["guess"]) for fr, raw in zip(frames, raws)])
guesses = np.stack([fr["guess"] for fr in frames])
r.set_frames(0, fin)
r.run_frames(0, F); r.sync()
res = r.get_results(0, F)
print("iters", res["iters"].tolist()[:16], "flags", sorted(set(res["flags"].tolist())))
r.enable_stage_timing(True); r.get_stage_ms(reset=True)
for _ in range(reps):
    r.set_poses(0, guesses); r.run_frames(0, F)
r.sync()
ms = r.get_stage_ms(reset=True)
print("stage ms per %d frames:" % F, {k: round(v[0] / max(v[1], 1), 3) for k, v in ms.items()})

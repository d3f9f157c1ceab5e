"""Diagnostics build only (make EXTRA=-DFBPR_KNN_STATS): distribution of search passes / candidates of the LM kernel's 5-NN.
python scripts/knn_stats.py [F]"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_base_pointcloud_registration_b200 as fb  # noqa: E402
import synth  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
frames = [synth.make_frame(4, i) for i in range(F)]
cfg = synth.CONFIGS[4]
r = fb.Registration(frames[0]["params"], max_frames=F, max_map_corner=cfg["map_corner"] + 64, max_map_surf=cfg["map_surf"] + 64)
raws = [fb.api.pack_raw(fr["scan"]) for fr in frames]
fin = r.make_frame_inputs([dict(raw_ptr=raw.ctypes.data, n_raw=len(raw), imu=fr["imu"], imu_available=fr["imu_available"],
                                map_corner_ptr=fr["map_corner"].ctypes.data, n_map_corner=len(fr["map_corner"]),
                                map_surf_ptr=fr["map_surf"].ctypes.data, n_map_surf=len(fr["map_surf"]), pose=fr["guess"]) for fr, raw in zip(frames, raws)])
r.set_frames(0, fin)
lib = fb.load_library()
out = (ctypes.c_ulonglong * 32)()
lib.fbpr_debug_knn_stats(out, 1)
r.run_frames(0, F); r.sync()
lib.fbpr_debug_knn_stats(out, 0)
s = np.array(list(out), dtype=np.float64)
print("frames", F, "point-iterations", s[19], "re-ranks", s[17], "certified", s[18])
print("searches", s[0], "per frame", s[0] / F, "passes per search", s[1] / max(s[0], 1))
print("candidates per pass", s[3] / max(s[1], 1), "rows per pass", s[4] / max(s[1], 1))
print("passes by candidates <=8,16,32,64,128,256,512,more:", (s[5:13] / max(s[1], 1)).round(3).tolist())
print("passes by rows <=9,25,32,more:", (s[13:17] / max(s[1], 1)).round(3).tolist())
print("searches by iteration 0..7+ (per frame):", (s[20:28] / F).round(0).tolist())

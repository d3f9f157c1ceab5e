import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, synth
import feature_base_pointcloud_registration_b200 as fb
F=128
frames=[synth.make_frame(4,i) for i in range(F)]
cfg=synth.CONFIGS[4]
r=fb.Registration(frames[0]["params"],max_frames=F,max_map_corner=cfg["map_corner"]+64,max_map_surf=cfg["map_surf"]+64)
for s,fr in enumerate(frames):
    r.set_raw_scan(s, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"])
r.project(0,F); r.sync()
r.enable_stage_timing(True); r.get_stage_ms(reset=True)
for _ in range(5): r.featureExtra(0,F)
r.sync(); ms=r.get_stage_ms(reset=True)
print("features ms per %d frames: %.3f"%(F, ms["features"][0]/ms["features"][1]))

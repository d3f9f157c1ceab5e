import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, synth
import feature_base_pointcloud_registration_b200 as fb
n=16
frames=[synth.make_frame(4,i) for i in range(n)]
cfg=synth.CONFIGS[4]
reg=fb.Registration(frames[0]["params"],max_frames=n,max_map_corner=cfg["map_corner"]+64,max_map_surf=cfg["map_surf"]+64)
stream=torch.cuda.ExternalStream(reg.stream(),device=torch.device("cuda",0))
for s,fr in enumerate(frames):
    reg.set_raw_scan(s, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"]); reg.set_local_map(s, fr["map_corner"], fr["map_surf"])
reg.use_graphs(True)
for name,fn in (("whole",lambda s: reg.run_frames(s,1)),("s2m",lambda s: reg.scan2MapOptimization(s,1))):
    ts=[]
    for rep in range(6):
        for s,fr in enumerate(frames):
            reg.set_pose(s, fr["guess"]); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
            reg.sync(); a.record(stream); fn(s); b.record(stream); reg.sync()
            if rep>0: ts.append(a.elapsed_time(b))
    print(name, "median %.4f p95 %.4f"%(np.median(ts), np.percentile(ts,95)), reg.get_results(0,4)["iters"])

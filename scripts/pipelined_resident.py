"""1024 resident config-4 frames: fbpr_run_frames batch by batch against fbpr_run_frames_pipelined (CUDA events, ms per 1024 frames)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, synth
import feature_base_pointcloud_registration_b200 as fb
F = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
CL = int(sys.argv[3]) if len(sys.argv) > 3 else 0
frames = [synth.make_frame(4, i % 128) for i in range(F)]
cfg = synth.CONFIGS[4]
r = fb.Registration(frames[0]["params"], max_frames=F, max_map_corner=cfg["map_corner"] + 64, max_map_surf=cfg["map_surf"] + 64, lm_cluster_size=CL)
stream = torch.cuda.ExternalStream(r.stream(), device=torch.device("cuda", 0))
for s, fr in enumerate(frames):
    r.set_raw_scan(s, fb.api.pack_raw(fr["scan"]), imu=fr["imu"], imu_available=fr["imu_available"]); r.set_local_map(s, fr["map_corner"], fr["map_surf"])
guess = np.stack([fr["guess"] for fr in frames])
def seq():
    for b in range(0, F, B): r.run_frames(b, min(B, F - b))
def pipe(): r.run_frames_pipelined(0, F, B)
res = {}
for name, fn in (("sequential", seq), ("pipelined", pipe), ("sequential", seq), ("pipelined", pipe)):
    r.set_poses(0, guess); fn(); r.sync()
    ts = []
    for rep in range(4):
        r.set_poses(0, guess); r.sync()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); r.sync(); ts.append(a.elapsed_time(b))
    out = r.get_results(0, F)
    if "ref" not in res: res["ref"] = out.copy()
    same = np.array_equal(out["pose"], res["ref"]["pose"]) and np.array_equal(out["iters"], res["ref"]["iters"])
    print(f"{name}: {np.median(ts):.2f} ms per {F} frames ({F / np.median(ts) * 1e3:.0f} frames/s), identical results: {same}")

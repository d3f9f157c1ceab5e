"""How much does the oracle's tie-break rule matter?  (VERDICT r01 "missing" 4)

The oracle makes the path's two unstable std::sort calls total orders by point index; the reference sorts by curvature only
(featureExtraction.h:13-17, :203) and pcl::VoxelGrid by voxel index only.  This script runs N synthetic frames through the CPU
oracle in both modes (oracle.set_literal_sort) and counts the frames whose outputs differ:
    python scripts/literal_sort_study.py [config=3] [frames=200]
"""
import multiprocessing as mp
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(args):
    cfg, idx = args
    import oracle
    import synth
    fr = synth.make_frame(cfg, idx)
    P = dict(fr["params"]); P["numberOfCores"] = 1
    out = []
    for mode in (0, 1):
        oracle.set_literal_sort(mode)
        ci = oracle.project(P, fr["scan"], fr["imu"], fr["imu_available"])
        fe = oracle.extract_features(P, ci)
        mo = oracle.MapOptimization(P); mo.set_imu(fr["imu_available"], 0.0, 0.0)
        mo.set_scan(fe["corner"], fe["surface"]); mo.set_map(fr["map_corner"], fr["map_surf"]); mo.downsample()
        pose, iters, flags, _ = mo.scan2map(fr["guess"])
        out.append((fe, mo.get_cloud(0).copy(), mo.get_cloud(1).copy(), pose.copy(), iters, flags))
    oracle.set_literal_sort(0)
    a, b = out
    curv = a[0]["curvature"]; nv = int(ci["n_valid"])
    return dict(
        curvature_ties=int((nv - 10) - len(np.unique(curv[5:nv - 5]))) if nv > 10 else 0,
        corner_index=not np.array_equal(a[0]["corner_index"], b[0]["corner_index"]),
        label=not np.array_equal(a[0]["label"], b[0]["label"]),
        surf_count=len(a[0]["surface"]) != len(b[0]["surface"]),
        surf_bits=not np.array_equal(a[0]["surface"], b[0]["surface"]),
        surf_maxdiff=float(np.abs(a[0]["surface"][:, :3] - b[0]["surface"][:, :3]).max()) if len(a[0]["surface"]) == len(b[0]["surface"]) else float("nan"),
        surf_maxdiff_i=float(np.abs(a[0]["surface"][:, 3] - b[0]["surface"][:, 3]).max()) if len(a[0]["surface"]) == len(b[0]["surface"]) else float("nan"),
        ds_count=len(a[2]) != len(b[2]) or len(a[1]) != len(b[1]),
        ds_bits=not (np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])),
        iters=a[4] != b[4], flags=a[5] != b[5],
        pose_bits=not np.array_equal(a[3], b[3]),
        pose_dt=float(np.abs(a[3][3:] - b[3][3:]).max()), pose_dr=float(np.abs(a[3][:3] - b[3][:3]).max()))


def main():
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        rows = pool.map(one, [(cfg, 1000 + i) for i in range(n)], chunksize=4)
    cnt = lambda k: sum(1 for r in rows if r[k])
    print(f"config {cfg}, {n} frames, oracle total-order sorts vs literal std::sort (libstdc++ introsort, {sys.version.split()[0]})")
    print(f"  frames with tied curvature values inside a ring window : {sum(1 for r in rows if r['curvature_ties'])} (ties per frame, max {max(r['curvature_ties'] for r in rows)})")
    for k, label in (("corner_index", "selected corner indices differ"), ("label", "cloudLabel differs"), ("surf_count", "surface cloud size differs"),
                     ("surf_bits", "surface cloud differs in any bit (voxel centroid summation order)"), ("ds_count", "downsampled scan sizes differ"),
                     ("ds_bits", "downsampled scan clouds differ in any bit"), ("iters", "LM iteration count differs"), ("flags", "outcome flags differ"),
                     ("pose_bits", "final pose differs in any bit")):
        print(f"  {label:75s}: {cnt(k)} / {n}")
    print(f"  max |centroid difference| over all frames                                  : xyz {np.nanmax([r['surf_maxdiff'] for r in rows]):.3e} m, intensity {np.nanmax([r['surf_maxdiff_i'] for r in rows]):.3e}")
    print(f"  max |pose difference|: translation {max(r['pose_dt'] for r in rows):.3e} m, rotation {max(r['pose_dr'] for r in rows):.3e} rad (tolerance 1e-4 / 1e-4)")


if __name__ == "__main__":
    main()

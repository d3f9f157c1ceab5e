"""Generates the committed golden fixtures (run here, in the authoring container):

  smallmat_cv2.npz    inputs and the outputs of cv2 4.13 (OpenCV's own JacobiImpl_ / QR32f / LU32f): pins the
                      oracle's restated small-matrix routines to the library the reference calls.
  flann_cv2.npz       inputs and the outputs of cv2.flann's KDTreeSingleIndex (OpenCV 4.13 vendors the FLANN library: the same
                      kd-tree class, result sets and L2 functor PCL's KdTreeFLANN instantiates): pins the oracle's exact 5-NN
                      (index order and squared distances, bit for bit) and the strict radius rule of extractNearby.
  pipeline_small.npz  a small seeded frame and the oracle's outputs for every stage: a regression pin of the
                      oracle itself and a fixed-input case for the GPU parity tests.

The reference has no tests, fixtures or golden vectors of its own and cannot be
built or imported here, so these are the only vectors there are.  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cv2  # noqa: E402

import oracle  # noqa: E402
import synth  # noqa: E402


def smallmat():
    rng = np.random.default_rng(20201018)
    A3, W3, V3, A6, W6, V6, B6, X6, I6 = [], [], [], [], [], [], [], [], []
    for _ in range(64):
        pts = rng.normal(size=(5, 3)).astype(np.float32) * rng.uniform(0.01, 1.0); pts[:, 0] *= rng.uniform(0.1, 20)
        d = pts - pts.mean(0); A = (d.T @ d / 5).astype(np.float32); A = ((A + A.T) * np.float32(0.5)).astype(np.float32)
        _, w, v = cv2.eigen(A); A3.append(A); W3.append(w.reshape(-1)); V3.append(v)
    for _ in range(32):
        J = rng.normal(size=(400, 6)).astype(np.float32); J[:, :3] *= rng.uniform(1, 30)
        A = (J.astype(np.float64).T @ J.astype(np.float64)).astype(np.float32); A = np.ascontiguousarray((A + A.T) * np.float32(0.5))
        b = (rng.normal(size=(6, 1)) * 50).astype(np.float32)
        _, w, v = cv2.eigen(A); _, x = cv2.solve(A, b, flags=cv2.DECOMP_QR); _, vi = cv2.invert(v, flags=cv2.DECOMP_LU)
        A6.append(A); W6.append(w.reshape(-1)); V6.append(v); B6.append(b.reshape(-1)); X6.append(x.reshape(-1)); I6.append(vi)
    np.savez_compressed(os.path.join(HERE, "smallmat_cv2.npz"), A3=np.array(A3), W3=np.array(W3), V3=np.array(V3), A6=np.array(A6),
                        W6=np.array(W6), V6=np.array(V6), B6=np.array(B6), X6=np.array(X6), I6=np.array(I6), cv2_version=cv2.__version__)


def flann():
    """pcl::KdTreeFLANN = flann::KDTreeSingleIndex, leaf_max_size 15, reorder, exact search (eps 0), sorted result set
    (SURVEY Appendix B-2); nearestKSearch(k = 5) at mapOptmization.h:1020 / :1143, radiusSearch at :880."""
    sp = dict(checks=-1, eps=0.0, sorted=True)
    fr = synth.make_frame(1, 7, small=(16, 600, 3000, 12000))
    rng = np.random.default_rng(20261019)
    out = {}
    for name in ("map_corner", "map_surf"):
        m = fr[name]; xyz = np.ascontiguousarray(m[:, :3], dtype=np.float32)
        q = xyz[rng.integers(0, len(xyz), 1500)] + rng.normal(scale=0.12, size=(1500, 3))
        q[:100] += rng.uniform(-2, 2, (100, 3))                           # some far from the map: the 1 m gate rejects them later
        q = np.ascontiguousarray(q, dtype=np.float32)
        index = cv2.flann_Index(xyz, dict(algorithm=4, leaf_max_size=15, reorder=True))       # 4 = FLANN_INDEX_KDTREE_SINGLE
        idx, d2 = index.knnSearch(q, 5, params=sp)
        out[name] = m.astype(np.float32); out[name + "_q"] = q; out[name + "_idx"] = idx.astype(np.int32); out[name + "_d2"] = d2.astype(np.float32)
    # lattice + exact duplicates: many exactly equal distances.  FLANN's order among equal distances follows its tree
    # traversal; the distances themselves are what the test pins there.
    g = np.stack(np.meshgrid(np.arange(10), np.arange(10), np.arange(5), indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * np.float32(0.25)
    g = g[rng.permutation(len(g))]; g = np.ascontiguousarray(np.concatenate([g, g[:40]]))
    q = np.ascontiguousarray(g[rng.integers(0, len(g), 400)] + np.float32(0.125) * rng.integers(0, 2, size=(400, 3)).astype(np.float32))
    idx, d2 = cv2.flann_Index(g, dict(algorithm=4, leaf_max_size=15, reorder=True)).knnSearch(q, 5, params=sp)
    out["lattice"] = g; out["lattice_q"] = q; out["lattice_idx"] = idx.astype(np.int32); out["lattice_d2"] = d2.astype(np.float32)
    # radius search over key poses (a trajectory that winds back on itself), query = the newest pose, radius 50 m; PCL hands
    # FLANN radius * radius.  Three poses sit exactly ON the sphere (3-4-5 triangles): FLANN's RadiusResultSet is strict.
    t = np.linspace(0, 6 * np.pi, 900)
    kp = np.stack([40 * np.cos(t) + 0.02 * t, 30 * np.sin(2 * t), 0.1 * t], 1).astype(np.float32)
    last = kp[-1].copy()
    kp[100] = last + np.array([30, 40, 0], np.float32); kp[200] = last + np.array([0, -30, 40], np.float32); kp[300] = last + np.array([48, 14, 0], np.float32)
    r2 = np.float32(50.0 * 50.0)
    n, ri, rd = cv2.flann_Index(np.ascontiguousarray(kp), dict(algorithm=4, leaf_max_size=15, reorder=True)).radiusSearch(last[None, :], float(r2), len(kp), params=sp)
    out["keyposes"] = kp; out["radius"] = np.float32(50.0); out["radius_idx"] = ri[0, :n].astype(np.int32); out["radius_d2"] = rd[0, :n].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "flann_cv2.npz"), cv2_version=cv2.__version__, **out)


def pipeline():
    small = (16, 450, 1500, 9000)
    fr = synth.make_frame(3, 11, small=small)            # config 3 flavour: IMU ramp -> deskew on
    P = fr["params"]
    ci = oracle.project(P, fr["scan"], fr["imu"], fr["imu_available"])
    fe = oracle.extract_features(P, ci)
    mo = oracle.MapOptimization(P)
    mo.set_imu(fr["imu_available"], 0.0, 0.0)
    mo.set_scan(fe["corner"], fe["surface"]); mo.set_map(fr["map_corner"], fr["map_surf"])
    nC, nS = mo.downsample()
    pose, iters, flags, _ = mo.scan2map(fr["guess"], debug_iter=0)
    d = mo.debug()
    np.savez_compressed(
        os.path.join(HERE, "pipeline_small.npz"), small=np.array(small), config=3, frame=11,
        n_raw=fr["scan"]["n"], raw_x=fr["scan"]["x"], raw_y=fr["scan"]["y"], raw_z=fr["scan"]["z"], raw_i=fr["scan"]["intensity"],
        raw_ring=fr["scan"]["ring"], raw_time=fr["scan"]["time"], guess=fr["guess"], gt=fr["gt"],
        map_corner=fr["map_corner"], map_surf=fr["map_surf"],
        startRing=ci["startRingIndex"], endRing=ci["endRingIndex"], colInd=ci["pointColInd"].astype(np.int16), rng=ci["pointRange"],
        cloud=ci["cloud_deskewed"], winner=ci["winner_raw"],
        label=fe["label"].astype(np.int8), picked=fe["picked"].astype(np.int8), corner_index=fe["corner_index"], surface=fe["surface"],
        cornerDS=mo.get_cloud(0), surfDS=mo.get_cloud(1),
        pose=pose, iters=iters, flags=flags, pose_trace=mo.pose_trace(), AtA=d["AtA"], AtB=d["AtB"], X=d["X"], nSel=d["nSel"],
        cornerKnn=d["cornerKnn"], surfKnn=d["surfKnn"], cornerFlag=d["cornerFlag"], surfFlag=d["surfFlag"])


if __name__ == "__main__":
    smallmat()
    flann()
    if "--flann-only" not in sys.argv:
        pipeline()
    for f in ("smallmat_cv2.npz", "flann_cv2.npz", "pipeline_small.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")

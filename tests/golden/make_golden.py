"""Generates the committed golden fixtures (run here, in the authoring container):

  smallmat_cv2.npz    inputs and the outputs of cv2 4.13 (OpenCV's own JacobiImpl_ / QR32f / LU32f): pins the
                      oracle's restated small-matrix routines to the library the reference calls.
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
    pipeline()
    for f in ("smallmat_cv2.npz", "pipeline_small.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")

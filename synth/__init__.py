"""Deterministic synthetic workloads -- ctypes front-end of synth.c.

`make_frame(config, frame)` returns the byte-identical inputs that the CPU oracle and the CUDA
path both consume for BASELINE.json's configs 1-5:

  1  VLP-16    16 x 1800, map  50 k (10 k corner + 40 k surf)
  2  HDL-32E   32 x 1800, map 200 k (40 k + 160 k)
  3  HDL-64E   64 x 2048, map 200 k, IMU rotation ramp (deskew on)        <- headline ms/frame
  4  = config 3 geometry, one independent (scan, map) pair per frame      <- headline frames/s
  5  OS1-128  128 x 2048, map 2 M (0.2 M + 1.8 M), surf leaf 0.2 m
  0  degenerate corridor (VLP-16) exercising the matP quirk (mapOptmization.h:1278)
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libsynth.so")
    src = os.path.join(_HERE, "synth.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fvisibility=hidden",
                               "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.synth_scan.restype = C.c_int
        _LIB.synth_imu_ramp.restype = C.c_int
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


CONFIGS = {
    0: dict(name="corridor-vlp16", n_scan=16, horizon=1800, map_corner=4000, map_surf=40000, scene=1, imu=False,
            surf_leaf=0.4, odo_leaf=0.4, jitter_surf=0.002),
    1: dict(name="vlp16-50k", n_scan=16, horizon=1800, map_corner=10000, map_surf=40000, scene=0, imu=False,
            surf_leaf=0.4, odo_leaf=0.4),
    2: dict(name="hdl32-200k", n_scan=32, horizon=1800, map_corner=40000, map_surf=160000, scene=0, imu=False,
            surf_leaf=0.4, odo_leaf=0.4),
    3: dict(name="hdl64-200k", n_scan=64, horizon=2048, map_corner=40000, map_surf=160000, scene=0, imu=True,
            surf_leaf=0.4, odo_leaf=0.4),
    4: dict(name="hdl64-200k-batch", n_scan=64, horizon=2048, map_corner=40000, map_surf=160000, scene=0, imu=True,
            surf_leaf=0.4, odo_leaf=0.4),
    5: dict(name="os1-128-2M", n_scan=128, horizon=2048, map_corner=200000, map_surf=1800000, scene=0, imu=False,
            surf_leaf=0.2, odo_leaf=0.2),
}

SCENE_SEED = 20201018
IMU_RATES = (0.02, 0.02, 0.3)        # rad/s roll, pitch, yaw during the sweep (config 3/4)
TIME_SCAN_CUR = 1000.0


def frame_seed(config, frame):
    return 20201018 + 1000 * config + frame


def params_for(config, number_of_cores=4):
    """The params.yaml knob values (reference config/params.yaml) adjusted per workload."""
    c = CONFIGS[config]
    return dict(N_SCAN=c["n_scan"], Horizon_SCAN=c["horizon"], edgeThreshold=1.0, surfThreshold=0.1,
                edgeFeatureMinValidNum=10, surfFeatureMinValidNum=100,
                odometrySurfLeafSize=c["odo_leaf"], mappingCornerLeafSize=0.2, mappingSurfLeafSize=c["surf_leaf"],
                z_tollerance=1000.0, rotation_tollerance=1000.0, numberOfCores=number_of_cores,
                surroundingKeyframeSearchRadius=50.0)


def make_pose(config, frame, dt_max=0.15, dr_max=np.deg2rad(1.5)):
    c = CONFIGS[config]
    gt = np.zeros(6, np.float64)
    guess = np.zeros(6, np.float64)
    _lib().synth_pose(C.c_int(c["scene"]), C.c_uint64(frame_seed(config, frame)), C.c_double(dt_max), C.c_double(dr_max),
                      _p(gt, C.c_double), _p(guess, C.c_double))
    return gt, guess


def make_scan(config, frame, pose, n_scan=None, horizon=None, rates=None, range_sigma=0.01, dropout=0.01, dup_frac=0.01):
    c = CONFIGS[config]
    n_scan = n_scan or c["n_scan"]
    horizon = horizon or c["horizon"]
    if rates is None:
        rates = IMU_RATES if c["imu"] else (0.0, 0.0, 0.0)
    cap = int(n_scan * horizon * (1.0 + dup_frac)) + 16
    x = np.zeros(cap, np.float32); y = np.zeros(cap, np.float32); z = np.zeros(cap, np.float32)
    inten = np.zeros(cap, np.float32); ring = np.zeros(cap, np.int32); t = np.zeros(cap, np.float32)
    pose = np.ascontiguousarray(pose, np.float64)
    rates_a = np.asarray(rates, np.float64)
    n = _lib().synth_scan(C.c_int(c["scene"]), C.c_uint64(SCENE_SEED), C.c_uint64(frame_seed(config, frame)),
                          C.c_int(n_scan), C.c_int(horizon), _p(pose, C.c_double), _p(rates_a, C.c_double),
                          C.c_double(range_sigma), C.c_double(dropout), C.c_double(dup_frac),
                          _p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), _p(inten, C.c_float),
                          _p(ring, C.c_int32), _p(t, C.c_float), C.c_int(cap))
    return dict(x=x[:n].copy(), y=y[:n].copy(), z=z[:n].copy(), intensity=inten[:n].copy(), ring=ring[:n].copy(),
                time=t[:n].copy(), n=n, rates=tuple(rates))


def make_imu_ramp(rates, hz=500.0):
    cap = 256
    it = np.zeros(cap, np.float64); rx = np.zeros(cap, np.float64); ry = np.zeros(cap, np.float64); rz = np.zeros(cap, np.float64)
    rates_a = np.asarray(rates, np.float64)
    n = _lib().synth_imu_ramp(C.c_double(TIME_SCAN_CUR), _p(rates_a, C.c_double), C.c_double(hz), C.c_int(cap),
                              _p(it, C.c_double), _p(rx, C.c_double), _p(ry, C.c_double), _p(rz, C.c_double))
    return dict(imuTime=it[:n].copy(), imuRotX=rx[:n].copy(), imuRotY=ry[:n].copy(), imuRotZ=rz[:n].copy(),
                imuPointerCur=n - 1, timeScanCur=TIME_SCAN_CUR)


def make_map(config, frame, n_corner=None, n_surf=None, jitter_corner=0.005, jitter_surf=None, centre=(0.0, 0.0, 0.0)):
    c = CONFIGS[config]
    if jitter_surf is None:
        jitter_surf = c.get("jitter_surf", 0.02)
    n_corner = n_corner or c["map_corner"]
    n_surf = n_surf or c["map_surf"]
    corner = np.zeros((n_corner, 4), np.float32)
    surf = np.zeros((n_surf, 4), np.float32)
    centre_a = np.asarray(centre, np.float64)
    _lib().synth_map(C.c_int(c["scene"]), C.c_uint64(SCENE_SEED), C.c_uint64(frame_seed(config, frame)),
                     C.c_int(n_corner), C.c_int(n_surf), C.c_double(jitter_corner), C.c_double(jitter_surf),
                     _p(centre_a, C.c_double), C.c_double(60.0), _p(corner, C.c_float), _p(surf, C.c_float))
    return corner, surf


def make_frame(config, frame, small=None):
    """One (scan, map, pose) workload.  `small` = (n_scan, horizon, map_corner, map_surf) shrinks it for CPU tests."""
    c = CONFIGS[config]
    gt, guess = make_pose(config, frame)
    n_scan, horizon = (small[0], small[1]) if small else (c["n_scan"], c["horizon"])
    mc, ms = (small[2], small[3]) if small else (c["map_corner"], c["map_surf"])
    rates = IMU_RATES if c["imu"] else (0.0, 0.0, 0.0)
    scan = make_scan(config, frame, gt, n_scan=n_scan, horizon=horizon, rates=rates)
    imu = make_imu_ramp(rates)
    corner, surf = make_map(config, frame, n_corner=mc, n_surf=ms, centre=gt[3:6])
    params = params_for(config)
    params["N_SCAN"] = n_scan
    params["Horizon_SCAN"] = horizon
    return dict(config=config, frame=frame, params=params, scan=scan, imu=imu, imu_available=1 if c["imu"] else 0,
                map_corner=corner, map_surf=surf, gt=gt.astype(np.float32), guess=guess.astype(np.float32))

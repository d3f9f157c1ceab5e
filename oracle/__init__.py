"""TEST INFRASTRUCTURE -- ctypes front-end of the CPU oracle (liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FLAG_NOT_ENOUGH_FEATURES = 1
FLAG_TOO_FEW_CORRESPONDENCES = 2
FLAG_DEGENERATE = 4
FLAG_CONVERGED = 8


class OrcParams(C.Structure):
    _fields_ = [("N_SCAN", C.c_int), ("Horizon_SCAN", C.c_int),
                ("edgeThreshold", C.c_float), ("surfThreshold", C.c_float),
                ("edgeFeatureMinValidNum", C.c_int), ("surfFeatureMinValidNum", C.c_int),
                ("odometrySurfLeafSize", C.c_float), ("mappingCornerLeafSize", C.c_float), ("mappingSurfLeafSize", C.c_float),
                ("z_tollerance", C.c_float), ("rotation_tollerance", C.c_float),
                ("numberOfCores", C.c_int), ("surroundingKeyframeSearchRadius", C.c_float)]


def make_params(d):
    p = OrcParams()
    for k, _ in OrcParams._fields_:
        setattr(p, k, d[k])
    return p


def build(force=False):
    if os.environ.get("ORACLE_SANITIZE"):          # scripts/oracle_sanitize.sh: the ASan + UBSan build (needs libasan preloaded)
        subprocess.check_call(["make", "-C", _HERE, "liboracle_asan.so"], stdout=subprocess.DEVNULL)
        return os.path.join(_HERE, "liboracle_asan.so")
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "ref_pipeline.hpp", "ref_cloud.hpp", "ref_smallmat.hpp")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_mo_create.restype = C.c_void_p
    return _LIB


def _f(a): return a.ctypes.data_as(C.POINTER(C.c_float))
def _i(a): return a.ctypes.data_as(C.POINTER(C.c_int))
def _d(a): return a.ctypes.data_as(C.POINTER(C.c_double))
def _u8(a): return a.ctypes.data_as(C.POINTER(C.c_uint8))


def f32(a): return np.ascontiguousarray(a, dtype=np.float32)
def i32(a): return np.ascontiguousarray(a, dtype=np.int32)


# ------------------------------------------------------------------ small matrices
def eigen_sym(A):
    A = f32(A); n = A.shape[0]
    W = np.zeros(n, np.float32); V = np.zeros((n, n), np.float32)
    lib().orc_eigen_sym(n, _f(A), _f(W), _f(V))
    return W, V


def qr_solve(A, b):
    A = f32(A); b = f32(b); n = A.shape[0]
    x = np.zeros(n, np.float32)
    ok = lib().orc_qr_solve(n, _f(A), _f(b), _f(x))
    return ok, x


def lu_invert(A):
    A = f32(A); n = A.shape[0]
    B = np.zeros((n, n), np.float32)
    ok = lib().orc_lu_invert(n, _f(A), _f(B))
    return ok, B


def matmul_f64acc(A, B):
    A = f32(A); B = f32(B)
    r, k = A.shape; c = B.shape[1]
    out = np.zeros((r, c), np.float32)
    lib().orc_matmul_f64acc(r, k, c, _f(A), _f(B), _f(out))
    return out


def colpiv_solve_5x3(A, b):
    A = f32(A); b = f32(b)
    x = np.zeros(3, np.float32)
    lib().orc_colpiv_solve_5x3(_f(A), _f(b), _f(x))
    return x


def get_transformation(pose6):
    p = f32(pose6); T = np.zeros(12, np.float32)
    lib().orc_get_transformation(_f(p), _f(T))
    return T.reshape(3, 4)


def get_translation_and_euler(T):
    T = f32(T).reshape(-1); p = np.zeros(6, np.float32)
    lib().orc_get_translation_and_euler(_f(T), _f(p))
    return p


# ------------------------------------------------------------------ cloud primitives
def set_literal_sort(on):
    """1: featureExtraction's std::sort compares the curvature only and VoxelGrid's the voxel index only, exactly as the reference
    does (ties land where libstdc++'s introsort leaves them); 0 (default): ties broken by point index.  Returns the old value."""
    return int(lib().orc_set_literal_sort(1 if on else 0))


def set_kdtree_flann_split(on):
    """1 (default): the kd-tree is built with FLANN's middleSplit_ rule (CPU-timing fidelity); 0: median split.  Same results."""
    return int(lib().orc_set_kdtree_flann_split(1 if on else 0))


def voxel_grid(xyzi, leaf):
    xyzi = f32(xyzi).reshape(-1, 4); n = xyzi.shape[0]
    out = np.zeros((max(n, 1), 4), np.float32)
    pk = np.zeros(max(n, 1), np.int32); ok = np.zeros(max(n, 1), np.int32); ov = C.c_int(0)
    m = lib().orc_voxel_grid(_f(xyzi), n, C.c_float(leaf), _f(out), _i(pk), _i(ok), C.byref(ov))
    return dict(points=out[:m].copy(), point_keys=pk[:n].copy(), out_keys=ok[:m].copy() if not ov.value else None,
                overflow=bool(ov.value))


def crop_box(xyzi, mn, mx):
    xyzi = f32(xyzi).reshape(-1, 4); n = xyzi.shape[0]
    out = np.zeros((max(n, 1), 4), np.float32)
    mn = f32(mn); mx = f32(mx)
    m = lib().orc_crop_box(_f(xyzi), n, _f(mn), _f(mx), _f(out))
    return out[:m].copy()


def knn5(map_xyzi, q_xyz, brute=False, threads=8):
    m = f32(map_xyzi).reshape(-1, 4); q = f32(q_xyz).reshape(-1, 3); nq = q.shape[0]
    idx = np.zeros((nq, 5), np.int32); d2 = np.zeros((nq, 5), np.float32)
    fn = lib().orc_brute_knn5 if brute else lib().orc_kdtree_knn5
    fn(_f(m), m.shape[0], _f(q), nq, _i(idx), _f(d2), threads)
    return idx, d2


# ------------------------------------------------------------------ projection / features
def project(params, scan, imu, imu_available, deskew_flag=1):
    p = make_params(params); cap = params["N_SCAN"] * params["Horizon_SCAN"]
    sr = np.zeros(params["N_SCAN"], np.int32); er = np.zeros(params["N_SCAN"], np.int32)
    col = np.zeros(cap, np.int32); rng = np.zeros(cap, np.float32); cloud = np.zeros((cap, 4), np.float32); win = np.zeros(cap, np.int32)
    x, y, z, it = f32(scan["x"]), f32(scan["y"]), f32(scan["z"]), f32(scan["intensity"])
    ring = i32(scan["ring"]); t = f32(scan["time"])
    nv = lib().orc_project(C.byref(p), _f(x), _f(y), _f(z), _f(it), _i(ring), _f(t), int(scan["n"]),
                           C.c_int64(imu_available), deskew_flag, C.c_double(imu["timeScanCur"]),
                           _d(imu["imuTime"]), _d(imu["imuRotX"]), _d(imu["imuRotY"]), _d(imu["imuRotZ"]), int(imu["imuPointerCur"]),
                           _i(sr), _i(er), _i(col), _f(rng), _f(cloud), _i(win))
    return dict(startRingIndex=sr, endRingIndex=er, pointColInd=col[:nv].copy(), pointRange=rng[:nv].copy(),
                cloud_deskewed=cloud[:nv].copy(), winner_raw=win[:nv].copy(), n_valid=nv)


def extract_features(params, ci):
    p = make_params(params); nv = int(ci["n_valid"]); cap = max(nv, 1)
    corner = np.zeros((cap, 4), np.float32); cidx = np.zeros(cap, np.int32)
    surf = np.zeros((cap, 4), np.float32); sidx = np.zeros(cap, np.int32)
    rc = np.zeros(params["N_SCAN"], np.int32); rcd = np.zeros(params["N_SCAN"], np.int32)
    curv = np.zeros(cap, np.float32); picked = np.zeros(cap, np.int32); label = np.zeros(cap, np.int32)
    counts = np.zeros(3, np.int32)
    lib().orc_extract_features(C.byref(p), _i(i32(ci["startRingIndex"])), _i(i32(ci["endRingIndex"])),
                               _i(i32(ci["pointColInd"])), _f(f32(ci["pointRange"])), _f(f32(ci["cloud_deskewed"])), nv,
                               _f(corner), _i(cidx), _f(surf), _i(sidx), _i(rc), _i(rcd), _f(curv), _i(picked), _i(label), _i(counts))
    return dict(corner=corner[:counts[0]].copy(), corner_index=cidx[:counts[0]].copy(),
                surface=surf[:counts[1]].copy(), surface_raw_index=sidx[:counts[2]].copy(),
                ring_surf_count=rc, ring_surf_count_ds=rcd, curvature=curv[:nv], picked=picked[:nv], label=label[:nv])


# ------------------------------------------------------------------ mapOptimization
def imu_deskew_info(queue8, time_scan_cur, time_scan_next, queue_length=2000):
    """ImageProjection::imuDeskewInfo (imageProjection.cpp:323-393).  queue8: [n,8] f64 rows (stamp, gyro xyz, orientation xyzw).
    Returns dict(popped, imuAvailable, imuPointerCur, imuRollInit, imuPitchInit, imuYawInit, imuTime, imuRotX, imuRotY, imuRotZ)."""
    q = np.ascontiguousarray(queue8, np.float64).reshape(-1, 8)
    t, rx, ry, rz = (np.zeros(queue_length, np.float64) for _ in range(4))
    out = np.zeros(5, np.float64)
    fn = lib().orc_imu_deskew_info
    fn.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double, C.c_double, C.c_int] + [C.POINTER(C.c_double)] * 5
    fn.restype = C.c_int
    popped = fn(_d(q), len(q), float(time_scan_cur), float(time_scan_next), int(queue_length), _d(t), _d(rx), _d(ry), _d(rz), _d(out))
    return dict(popped=popped, imuAvailable=int(out[0]), imuPointerCur=int(out[1]), imuRollInit=np.float32(out[2]),
                imuPitchInit=np.float32(out[3]), imuYawInit=np.float32(out[4]), imuTime=t, imuRotX=rx, imuRotY=ry, imuRotZ=rz)


class MapOptimization:
    """Oracle counterpart of the reference's mapOptimization operator surface."""

    def __init__(self, params):
        self.params = dict(params)
        self._p = make_params(params)
        self.h = C.c_void_p(lib().orc_mo_create(C.byref(self._p)))
        self.nC = self.nS = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_mo_destroy(self.h); self.h = None

    def set_threads(self, n): lib().orc_mo_set_threads(self.h, int(n))

    def set_scan(self, corner, surf):
        c = f32(corner).reshape(-1, 4); s = f32(surf).reshape(-1, 4)
        lib().orc_mo_set_scan(self.h, _f(c), c.shape[0], _f(s), s.shape[0])

    def set_map(self, corner, surf):
        c = f32(corner).reshape(-1, 4); s = f32(surf).reshape(-1, 4)
        lib().orc_mo_set_map(self.h, _f(c), c.shape[0], _f(s), s.shape[0])

    def set_imu(self, available, roll, pitch):
        lib().orc_mo_set_imu(self.h, C.c_int64(available), C.c_float(roll), C.c_float(pitch))

    def extract_cloud(self, key_poses6, corner_frames, surf_frames, last_key_xyz):
        K = len(corner_frames)
        kp = f32(key_poses6).reshape(K, 6)
        coff = np.zeros(K + 1, np.int32); soff = np.zeros(K + 1, np.int32)
        coff[1:] = np.cumsum([len(c) for c in corner_frames]); soff[1:] = np.cumsum([len(s) for s in surf_frames])
        call = f32(np.concatenate(corner_frames)).reshape(-1, 4); sall = f32(np.concatenate(surf_frames)).reshape(-1, 4)
        counts = np.zeros(4, np.int32); lk = f32(last_key_xyz)
        lib().orc_mo_extract_cloud(self.h, _f(kp), K, _f(call), _i(coff), _f(sall), _i(soff), _f(lk), _i(counts))
        return counts

    def extract_surrounding(self, key_poses6_all, key_times, density, time_last, corner_frames, surf_frames,
                            loop_closure=False, keyframe_size=50):
        """extractSurroundingKeyFrames over the whole keyframe store (mapOptmization.h:964-978): extractNearby (:872-907) or, with
        loop_closure, extractForLoopClosure (:857-870), then extractCloud (:909-955).
        Returns (cloudToExtract [m,4], counts[4])."""
        n = len(corner_frames)
        kp = f32(key_poses6_all).reshape(n, 6); kt = np.ascontiguousarray(key_times, np.float64)
        coff = np.zeros(n + 1, np.int32); soff = np.zeros(n + 1, np.int32)
        coff[1:] = np.cumsum([len(c) for c in corner_frames]); soff[1:] = np.cumsum([len(s) for s in surf_frames])
        call = f32(np.concatenate(corner_frames)).reshape(-1, 4); sall = f32(np.concatenate(surf_frames)).reshape(-1, 4)
        counts = np.zeros(4, np.int32); ds = np.zeros((2 * n + 8, 4), np.float32)
        fn = lib().orc_mo_extract_surrounding
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_int, C.c_float, C.c_double,
                       C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_int),
                       C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int]
        fn.restype = C.c_int
        m = fn(self.h, _f(kp), _d(kt), n, float(density), float(time_last), _f(call), _i(coff), _f(sall), _i(soff), _f(ds), len(ds), _i(counts),
               1 if loop_closure else 0, int(keyframe_size))
        return ds[:m].copy(), counts

    def downsample(self):
        counts = np.zeros(2, np.int32)
        lib().orc_mo_downsample(self.h, _i(counts))
        self.nC, self.nS = int(counts[0]), int(counts[1])
        return self.nC, self.nS

    def get_cloud(self, which):
        n = lib().orc_mo_get_cloud(self.h, which, None, 0)
        out = np.zeros((max(n, 1), 4), np.float32)
        lib().orc_mo_get_cloud(self.h, which, _f(out), n)
        return out[:n].copy()

    def scan2map(self, pose6, debug_iter=-1):
        pose = f32(pose6).copy(); iters = C.c_int(0); flags = C.c_uint(0); secs = np.zeros(2, np.float64)
        lib().orc_mo_scan2map(self.h, _f(pose), int(debug_iter), C.byref(iters), C.byref(flags), _d(secs))
        return pose, iters.value, flags.value, secs

    def transform_update(self, pose6):
        pose = f32(pose6).copy()
        lib().orc_mo_transform_update(self.h, _f(pose))
        return pose

    def registration(self, corner_global, surf_global, pose12):
        c = f32(corner_global).reshape(-1, 4); s = f32(surf_global).reshape(-1, 4)
        T = f32(pose12).reshape(-1).copy(); iters = C.c_int(0); flags = C.c_uint(0)
        lib().orc_mo_registration(self.h, _f(c), c.shape[0], _f(s), s.shape[0], _f(T), C.byref(iters), C.byref(flags))
        return T.reshape(3, 4), iters.value, flags.value

    def pose_trace(self):
        out = np.zeros((30, 6), np.float32)
        n = lib().orc_mo_pose_trace(self.h, _f(out), 30)
        return out[:n].copy()

    def debug(self):
        nC, nS = self.nC, self.nS
        d = dict(cornerKnn=np.zeros((nC, 5), np.int32), cornerD2=np.zeros((nC, 5), np.float32),
                 cornerCoeff=np.zeros((nC, 4), np.float32), cornerFlag=np.zeros(nC, np.uint8),
                 surfKnn=np.zeros((nS, 5), np.int32), surfD2=np.zeros((nS, 5), np.float32),
                 surfCoeff=np.zeros((nS, 4), np.float32), surfFlag=np.zeros(nS, np.uint8),
                 AtA=np.zeros((6, 6), np.float32), AtB=np.zeros(6, np.float32), X=np.zeros(6, np.float32))
        nsel = C.c_int(0)
        it = lib().orc_mo_debug(self.h, _i(d["cornerKnn"]), _f(d["cornerD2"]), _f(d["cornerCoeff"]), _u8(d["cornerFlag"]),
                                _i(d["surfKnn"]), _f(d["surfD2"]), _f(d["surfCoeff"]), _u8(d["surfFlag"]),
                                _f(d["AtA"]), _f(d["AtB"]), _f(d["X"]), C.byref(nsel))
        d["iter"] = it; d["nSel"] = nsel.value
        return d

"""ctypes binding of include/fbpr_b200.h.  Mirrors the reference's operator surface
(featureExtra / extractSurroundingKeyFrames / downsampleCurrentScan / scan2MapOptimization /
transformUpdate / registration) on frame SLOTS of one device-resident handle."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FLAG_NOT_ENOUGH_FEATURES = 1
FLAG_TOO_FEW_CORRESPONDENCES = 2
FLAG_DEGENERATE = 4
FLAG_CONVERGED = 8
FLAG_MAP_TRUNCATED = 16
MEM_HOST, MEM_DEVICE = 0, 1

RAW_POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4"), ("ring", "<i4"), ("time", "<f4")])
RESULT_DTYPE = np.dtype([("pose", "<f4", (6,)), ("iters", "<i4"), ("flags", "<u4")])

_BUF_NAMES = ["START_RING", "END_RING", "COL_IND", "RANGE", "CLOUD", "CURVATURE", "PICKED", "LABEL",
              "CORNER", "CORNER_INDEX", "SURF", "RING_SURF_COUNT", "RING_SURF_COUNT_DS",
              "CORNER_DS", "SURF_DS", "MAP_CORNER", "MAP_SURF",
              "KNN_CORNER", "KNN_SURF", "KNN_D2_CORNER", "KNN_D2_SURF",
              "COEFF_CORNER", "COEFF_SURF", "FLAG_CORNER", "FLAG_SURF", "ATA", "ATB", "X", "POSE_TRACE", "WINNER_RAW"]
BUF = {n: i for i, n in enumerate(_BUF_NAMES)}
_BUF_DTYPE = dict(START_RING=(np.int32, 1), END_RING=(np.int32, 1), COL_IND=(np.int32, 1), RANGE=(np.float32, 1),
                  CLOUD=(np.float32, 4), CURVATURE=(np.float32, 1), PICKED=(np.int32, 1), LABEL=(np.int32, 1),
                  CORNER=(np.float32, 4), CORNER_INDEX=(np.int32, 1), SURF=(np.float32, 4),
                  RING_SURF_COUNT=(np.int32, 1), RING_SURF_COUNT_DS=(np.int32, 1),
                  CORNER_DS=(np.float32, 4), SURF_DS=(np.float32, 4), MAP_CORNER=(np.float32, 4), MAP_SURF=(np.float32, 4),
                  KNN_CORNER=(np.int32, 5), KNN_SURF=(np.int32, 5), KNN_D2_CORNER=(np.float32, 5), KNN_D2_SURF=(np.float32, 5),
                  COEFF_CORNER=(np.float32, 4), COEFF_SURF=(np.float32, 4), FLAG_CORNER=(np.uint8, 1), FLAG_SURF=(np.uint8, 1),
                  ATA=(np.float32, 6), ATB=(np.float32, 1), X=(np.float32, 1), POSE_TRACE=(np.float32, 6), WINNER_RAW=(np.int32, 1))


class FbprError(RuntimeError):
    pass


class Params(C.Structure):
    """fbpr_params: the params.yaml knobs the path reads + device capacities."""
    _fields_ = [("N_SCAN", C.c_int32), ("Horizon_SCAN", C.c_int32),
                ("edgeThreshold", C.c_float), ("surfThreshold", C.c_float),
                ("edgeFeatureMinValidNum", C.c_int32), ("surfFeatureMinValidNum", C.c_int32),
                ("odometrySurfLeafSize", C.c_float), ("mappingCornerLeafSize", C.c_float), ("mappingSurfLeafSize", C.c_float),
                ("z_tollerance", C.c_float), ("rotation_tollerance", C.c_float),
                ("numberOfCores", C.c_int32), ("surroundingKeyframeSearchRadius", C.c_float),
                ("max_frames", C.c_int32), ("max_raw_points", C.c_int32),
                ("max_map_corner", C.c_int32), ("max_map_surf", C.c_int32), ("max_keyframe_points", C.c_int32),
                ("knn_cell_corner", C.c_float), ("knn_cell_surf", C.c_float),
                ("grid_cells_corner", C.c_int32), ("grid_cells_surf", C.c_int32),
                ("lm_cluster_size", C.c_int32), ("lm_single_frame_mode", C.c_int32), ("knn_first_radius", C.c_float)]

    @classmethod
    def from_dict(cls, d, **extra):
        p = cls()
        names = {n for n, _ in cls._fields_}
        for k, v in {**d, **extra}.items():
            if k in names:
                setattr(p, k, v)
        if p.max_frames <= 0:
            p.max_frames = 1
        return p


class CloudInfoView(C.Structure):
    _fields_ = [("startRingIndex", C.POINTER(C.c_int32)), ("endRingIndex", C.POINTER(C.c_int32)),
                ("pointColInd", C.POINTER(C.c_int32)), ("pointRange", C.POINTER(C.c_float)),
                ("cloud_deskewed", C.POINTER(C.c_float)), ("n_valid", C.c_int32),
                ("imuAvailable", C.c_int64), ("imuRollInit", C.c_float), ("imuPitchInit", C.c_float), ("imuYawInit", C.c_float)]


class Pc2Layout(C.Structure):                                # fbpr_pc2_layout
    _fields_ = [(k, C.c_int32) for k in ("point_step", "off_x", "off_y", "off_z", "off_intensity", "off_ring", "ring_bytes", "off_time")]


class FrameInput(C.Structure):
    _fields_ = [("raw", C.c_void_p), ("n_raw", C.c_int32), ("deskewFlag", C.c_int32), ("imuAvailable", C.c_int64),
                ("timeScanCur", C.c_double), ("imuTime", C.c_void_p), ("imuRotX", C.c_void_p), ("imuRotY", C.c_void_p), ("imuRotZ", C.c_void_p),
                ("imuPointerCur", C.c_int32), ("imuRollInit", C.c_float), ("imuPitchInit", C.c_float),
                ("map_corner_xyzi", C.c_void_p), ("n_map_corner", C.c_int32), ("map_surf_xyzi", C.c_void_p), ("n_map_surf", C.c_int32),
                ("pose", C.c_float * 6), ("raw_format", C.c_int32), ("map_format", C.c_int32)]


RAW_PACKED24, RAW_VELODYNE22 = 0, 1
MAP_XYZI16, MAP_XYZ12, MAP_FROM_GLOBAL = 0, 1, 2
VELODYNE22_DTYPE = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"], "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                             "offsets": [0, 4, 8, 12, 16, 18], "itemsize": 22})


STAGES = ["project", "features", "downsample", "map_index", "lm"]


def library_path():
    # FBPR_B200_LIB lets a developer A/B two builds of the same library; it is still the CUDA library
    return os.environ.get("FBPR_B200_LIB") or os.path.join(_HERE, "libfbpr_b200.so")


def load_library():
    """Load libfbpr_b200.so; raise loudly if it has not been built (there is no fallback path)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise FbprError(f"{path} is missing: build it with `make` or `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(the CUDA library is the only implementation; there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.fbpr_last_error.restype = C.c_char_p
    lib.fbpr_stream.restype = C.c_void_p
    lib.fbpr_stream.argtypes = [C.c_void_p]
    lib.fbpr_kernel_launches.restype = C.c_int64
    lib.fbpr_kernel_launches.argtypes = [C.c_void_p]
    lib.fbpr_get_buffer.restype = C.c_int64
    lib.fbpr_get_buffer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64]
    lib.fbpr_destroy.argtypes = [C.c_void_p]
    lib.fbpr_destroy.restype = None
    _LIB = lib
    return lib


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def pack_raw(scan):
    """synth scan dict (SoA) -> array of fbpr_raw_point records."""
    n = int(scan["n"])
    raw = np.zeros(n, RAW_POINT_DTYPE)
    for k in ("x", "y", "z", "intensity", "ring", "time"):
        raw[k] = scan[k][:n]
    return raw


def pack_wire22(scan):
    """synth scan dict (SoA) -> the Velodyne driver's 22-byte PointXYZIRT records (RAW_VELODYNE22)."""
    n = int(scan["n"])
    w = np.zeros(n, VELODYNE22_DTYPE)
    for k in ("x", "y", "z", "intensity", "ring", "time"):
        w[k] = scan[k][:n]
    return w


class Registration:
    """One device-resident handle = `max_frames` frame slots on one GPU and one stream."""

    def __init__(self, params, device=0, **extra):
        self.lib = load_library()
        self.params = params if isinstance(params, Params) else Params.from_dict(params, **extra)
        self.h = C.c_void_p()
        rc = self.lib.fbpr_create(C.byref(self.params), int(device), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise FbprError(self.lib.fbpr_last_error().decode())
        self.F = self.params.max_frames
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.fbpr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise FbprError(self.lib.fbpr_last_error().decode())
        return rc

    # ---- inputs
    def set_raw_scan(self, slot, raw, imu=None, imu_available=0, deskew_flag=1, imu_roll_init=0.0, imu_pitch_init=0.0):
        raw = np.ascontiguousarray(raw, dtype=RAW_POINT_DTYPE)
        if imu is not None and imu_available:
            args = (C.c_double(imu["timeScanCur"]), _vp(imu["imuTime"]), _vp(imu["imuRotX"]), _vp(imu["imuRotY"]), _vp(imu["imuRotZ"]),
                    C.c_int(int(imu["imuPointerCur"])))
        else:
            args = (C.c_double(0.0), None, None, None, None, C.c_int(0))
        self._ck(self.lib.fbpr_set_raw_scan(self.h, slot, _vp(raw), len(raw), MEM_HOST, C.c_int64(imu_available), C.c_int(deskew_flag),
                                            *args, C.c_float(imu_roll_init), C.c_float(imu_pitch_init)))
        self._keep = [raw]

    def set_raw_scan_pc2(self, slot, data, n, layout, imu=None, imu_available=0, imu_roll_init=0.0, imu_pitch_init=0.0):
        """sensor_msgs/PointCloud2 payload bytes (any point_step / field offsets) -> the slot's raw scan, repacked on the device.
        layout: dict(point_step, off_x, off_y, off_z, off_intensity, off_ring, ring_bytes, off_time)."""
        buf = np.ascontiguousarray(np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8).reshape(-1))
        L = Pc2Layout(*[int(layout[k]) for k, _ in Pc2Layout._fields_])
        if imu is not None and imu_available:
            args = (C.c_double(imu["timeScanCur"]), _vp(imu["imuTime"]), _vp(imu["imuRotX"]), _vp(imu["imuRotY"]), _vp(imu["imuRotZ"]),
                    C.c_int(int(imu["imuPointerCur"])))
        else:
            args = (C.c_double(0.0), None, None, None, None, C.c_int(0))
        self._ck(self.lib.fbpr_set_raw_scan_pc2(self.h, slot, _vp(buf), int(n), C.byref(L), MEM_HOST, C.c_int64(imu_available), *args,
                                                C.c_float(imu_roll_init), C.c_float(imu_pitch_init)))

    def set_clouds_xyzi32(self, slot, kind, corner32, surf32):
        """32-byte pcl::PointXYZI records ([n,8] f32) as the slot's feature clouds (kind 0) or local map (kind 1)."""
        c = _f32(corner32).reshape(-1, 8); s = _f32(surf32).reshape(-1, 8)
        self._ck(self.lib.fbpr_set_clouds_xyzi32(self.h, slot, int(kind), _vp(c), len(c), _vp(s), len(s)))

    def get_buffer_xyzi32(self, slot, name):
        cap = max(self.params.N_SCAN * self.params.Horizon_SCAN, self.params.max_map_corner, self.params.max_map_surf, 65536) * 32 + 1024
        buf = np.zeros(cap, np.uint8)
        self.lib.fbpr_get_buffer_xyzi32.restype = C.c_int64
        self.lib.fbpr_get_buffer_xyzi32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64]
        nb = self._ck(self.lib.fbpr_get_buffer_xyzi32(self.h, slot, BUF[name], _vp(buf), cap))
        return buf[:nb].view(np.float32).reshape(-1, 8).copy()

    def extractCloud(self, slot, key_poses6, check_xyz, corner_frames, surf_frames, last_key_xyz):
        """fbpr_extract_cloud: extractCloud with explicit re-check positions (entry i re-checked at check_xyz[i])."""
        K = len(corner_frames)
        kp = _f32(key_poses6).reshape(K, 6); ck = _f32(check_xyz).reshape(K, 3) if check_xyz is not None else None
        coff = np.zeros(K + 1, np.int32); soff = np.zeros(K + 1, np.int32)
        coff[1:] = np.cumsum([len(c) for c in corner_frames]); soff[1:] = np.cumsum([len(s) for s in surf_frames])
        call = _f32(np.concatenate(corner_frames)).reshape(-1, 4); sall = _f32(np.concatenate(surf_frames)).reshape(-1, 4)
        lk = _f32(last_key_xyz)
        self._ck(self.lib.fbpr_extract_cloud(self.h, slot, K, _vp(kp), _vp(ck), _vp(call), _vp(coff), _vp(sall), _vp(soff), _vp(lk), MEM_HOST))
        self.sync()

    def set_raw_scan_device(self, slot, dev_ptr, n, imu=None, imu_available=0, deskew_flag=1):
        if imu is not None and imu_available:
            args = (C.c_double(imu["timeScanCur"]), _vp(imu["imuTime"]), _vp(imu["imuRotX"]), _vp(imu["imuRotY"]), _vp(imu["imuRotZ"]),
                    C.c_int(int(imu["imuPointerCur"])))
        else:
            args = (C.c_double(0.0), None, None, None, None, C.c_int(0))
        self._ck(self.lib.fbpr_set_raw_scan(self.h, slot, C.c_void_p(dev_ptr), int(n), MEM_DEVICE, C.c_int64(imu_available),
                                            C.c_int(deskew_flag), *args, C.c_float(0.0), C.c_float(0.0)))

    def set_cloud_info(self, slot, ci, imu_available=0, imu_roll_init=0.0, imu_pitch_init=0.0):
        sr = np.ascontiguousarray(ci["startRingIndex"], np.int32); er = np.ascontiguousarray(ci["endRingIndex"], np.int32)
        col = np.ascontiguousarray(ci["pointColInd"], np.int32); rng = _f32(ci["pointRange"]); cl = _f32(ci["cloud_deskewed"])
        v = CloudInfoView(sr.ctypes.data_as(C.POINTER(C.c_int32)), er.ctypes.data_as(C.POINTER(C.c_int32)),
                          col.ctypes.data_as(C.POINTER(C.c_int32)), rng.ctypes.data_as(C.POINTER(C.c_float)),
                          cl.ctypes.data_as(C.POINTER(C.c_float)), len(col), imu_available, imu_roll_init, imu_pitch_init, 0.0)
        self._ck(self.lib.fbpr_set_cloud_info(self.h, slot, C.byref(v), MEM_HOST))

    def set_feature_clouds(self, slot, corner, surf):
        c = _f32(corner).reshape(-1, 4); s = _f32(surf).reshape(-1, 4)
        self._ck(self.lib.fbpr_set_feature_clouds(self.h, slot, _vp(c), len(c), _vp(s), len(s), MEM_HOST))

    def set_local_map(self, slot, corner, surf):
        c = _f32(corner).reshape(-1, 4); s = _f32(surf).reshape(-1, 4)
        self._ck(self.lib.fbpr_set_local_map(self.h, slot, _vp(c), len(c), _vp(s), len(s), MEM_HOST))

    def set_local_map_device(self, slot, corner_ptr, n_corner, surf_ptr, n_surf):
        self._ck(self.lib.fbpr_set_local_map(self.h, slot, C.c_void_p(corner_ptr), int(n_corner), C.c_void_p(surf_ptr), int(n_surf), MEM_DEVICE))

    def set_pose(self, slot, pose6):
        p = _f32(pose6)
        self._ck(self.lib.fbpr_set_pose(self.h, slot, _vp(p)))

    def set_poses(self, first, poses6):
        p = _f32(poses6).reshape(-1, 6)
        self._ck(self.lib.fbpr_set_poses(self.h, first, len(p), _vp(p), MEM_HOST))

    def set_poses_device(self, first, count, dev_ptr):
        self._ck(self.lib.fbpr_set_poses(self.h, first, count, C.c_void_p(dev_ptr), MEM_DEVICE))

    def make_frame_inputs(self, frames):
        """frames: list of dicts with raw_ptr/n_raw, imu (dict or None), imu_available, map_corner_ptr/n, map_surf_ptr/n, pose.
        Pointers are integers (host or device addresses); the arrays must stay alive until the copies complete."""
        arr = (FrameInput * len(frames))()
        for i, f in enumerate(frames):
            a = arr[i]
            a.raw = f.get("raw_ptr"); a.n_raw = int(f.get("n_raw", 0)); a.deskewFlag = int(f.get("deskew_flag", 1))
            a.imuAvailable = int(f.get("imu_available", 0))
            imu = f.get("imu")
            if imu is not None and a.imuAvailable:
                a.timeScanCur = float(imu["timeScanCur"]); a.imuPointerCur = int(imu["imuPointerCur"])
                a.imuTime = imu["imuTime"].ctypes.data; a.imuRotX = imu["imuRotX"].ctypes.data
                a.imuRotY = imu["imuRotY"].ctypes.data; a.imuRotZ = imu["imuRotZ"].ctypes.data
            a.imuRollInit = float(f.get("imu_roll_init", 0.0)); a.imuPitchInit = float(f.get("imu_pitch_init", 0.0))
            a.map_corner_xyzi = f.get("map_corner_ptr"); a.n_map_corner = int(f.get("n_map_corner", 0))
            a.map_surf_xyzi = f.get("map_surf_ptr"); a.n_map_surf = int(f.get("n_map_surf", 0))
            for q in range(6):
                a.pose[q] = float(f["pose"][q])
            a.raw_format = int(f.get("raw_format", RAW_PACKED24)); a.map_format = int(f.get("map_format", MAP_XYZI16))
        return arr

    def set_frames(self, first, frame_inputs, mem=MEM_HOST):
        self._ck(self.lib.fbpr_set_frames(self.h, first, len(frame_inputs), frame_inputs, mem))

    def register_frames(self, first, frame_inputs, chunk_frames=0):
        """Upload (host buffers), run the whole path and fetch the results of len(frame_inputs) independent frames,
        with the uploads of one chunk overlapping the kernels of the previous one."""
        out = np.zeros(len(frame_inputs), RESULT_DTYPE)
        self._ck(self.lib.fbpr_register_frames(self.h, first, len(frame_inputs), frame_inputs, int(chunk_frames), _vp(out)))
        return out

    def register_frames_begin(self, first, frame_inputs, chunk_frames=0):
        """Enqueue uploads + the whole path of one batch; returns a ticket.  Batches on disjoint slot ranges may overlap."""
        t = self._ck(self.lib.fbpr_register_frames_begin(self.h, first, len(frame_inputs), frame_inputs, int(chunk_frames)))
        self._tickets = getattr(self, "_tickets", {}); self._tickets[t] = (len(frame_inputs), frame_inputs)
        return t

    def register_frames_end(self, ticket):
        n, _ = self._tickets.pop(ticket)
        out = np.zeros(n, RESULT_DTYPE)
        self._ck(self.lib.fbpr_register_frames_end(self.h, int(ticket), _vp(out)))
        return out

    def enable_stage_timing(self, on=True): self._ck(self.lib.fbpr_enable_stage_timing(self.h, int(on)))

    def get_stage_ms(self, reset=True):
        ms = (C.c_float * 5)(); calls = (C.c_int32 * 5)()
        self._ck(self.lib.fbpr_get_stage_ms(self.h, ms, calls, int(reset)))
        return {n: (float(ms[i]), int(calls[i])) for i, n in enumerate(STAGES)}

    # ---- operators (reference names)
    def project(self, first=0, count=1): self._ck(self.lib.fbpr_project(self.h, first, count))
    def featureExtra(self, first=0, count=1): self._ck(self.lib.fbpr_feature_extract(self.h, first, count))
    def downsampleCurrentScan(self, first=0, count=1): self._ck(self.lib.fbpr_downsample_current_scan(self.h, first, count))
    def scan2MapOptimization(self, first=0, count=1): self._ck(self.lib.fbpr_scan2map_optimization(self.h, first, count))
    def transformUpdate(self, first=0, count=1): self._ck(self.lib.fbpr_transform_update(self.h, first, count))

    def extractSurroundingKeyFrames(self, slot, key_poses6, corner_frames, surf_frames, last_key_xyz):
        K = len(corner_frames)
        kp = _f32(key_poses6).reshape(K, 6)
        coff = np.zeros(K + 1, np.int32); soff = np.zeros(K + 1, np.int32)
        coff[1:] = np.cumsum([len(c) for c in corner_frames]); soff[1:] = np.cumsum([len(s) for s in surf_frames])
        call = _f32(np.concatenate(corner_frames)).reshape(-1, 4); sall = _f32(np.concatenate(surf_frames)).reshape(-1, 4)
        lk = _f32(last_key_xyz)
        self._ck(self.lib.fbpr_extract_surrounding_keyframes(self.h, slot, K, _vp(kp), _vp(call), _vp(coff), _vp(sall), _vp(soff), _vp(lk), MEM_HOST))
        self.sync()

    def check_guards(self):
        """FBPR_GUARD=1 handles: number of overwritten guard zones (0 = no out-of-bounds write past any device array)"""
        return self._ck(self.lib.fbpr_debug_check_guards(self.h))

    # ---- resident keyframe store (cloudKeyPoses3D/6D + corner/surfCloudKeyFrames, mapOptmization.h:84-88)
    def keyframes_clear(self): self._ck(self.lib.fbpr_keyframes_clear(self.h))
    def keyframes_count(self): return self._ck(self.lib.fbpr_keyframes_count(self.h))

    def keyframe_push(self, pose6, time, corner, surf):
        """saveKeyFramesAndFactor's push_backs (:1690-1726); returns the keyframe index"""
        p = _f32(pose6).reshape(6); c = _f32(corner).reshape(-1, 4); s = _f32(surf).reshape(-1, 4)
        self.lib.fbpr_keyframe_push.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        return self._ck(self.lib.fbpr_keyframe_push(self.h, _vp(p), float(time), _vp(c), len(c), _vp(s), len(s), MEM_HOST))

    def keyframes_set_poses(self, first, poses6):
        p = _f32(poses6).reshape(-1, 6)
        self._ck(self.lib.fbpr_keyframes_set_poses(self.h, first, len(p), _vp(p)))

    def extractSurroundingKeyFramesResident(self, slot, time_last, density=1.0, loop_closure=False, keyframe_size=50):
        """mapOptimization::extractSurroundingKeyFrames (:964-978) end to end on the device over the resident store"""
        self.lib.fbpr_extract_surrounding_keyframes_resident.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_float, C.c_int, C.c_int]
        self._ck(self.lib.fbpr_extract_surrounding_keyframes_resident(self.h, slot, float(time_last), float(density), 1 if loop_closure else 0, int(keyframe_size)))

    def keyframe_selection(self):
        """cloudToExtract of the last resident extraction: (list [K,4], keyframe index per entry, -1 = dropped by the re-check)"""
        n = max(self.keyframes_count(), 1)
        lst = np.zeros((2 * n + 8, 4), np.float32); idx = np.zeros(2 * n + 8, np.int32)
        K = self._ck(self.lib.fbpr_get_keyframe_selection(self.h, _vp(lst), _vp(idx), len(lst)))
        return lst[:K].copy(), idx[:K].copy()

    def set_global_map(self, corner_global, surf_global):
        c = _f32(corner_global).reshape(-1, 4); s = _f32(surf_global).reshape(-1, 4)
        self._ck(self.lib.fbpr_set_global_map(self.h, _vp(c), len(c), _vp(s), len(s), MEM_HOST))

    def crop_local_maps(self, first=0, count=1):
        """fbpr_crop_local_maps: CropBox of the resident global maps around every slot's pose (mapOptmization.h:284-304), batched"""
        self._ck(self.lib.fbpr_crop_local_maps(self.h, first, count))

    def registration(self, slot, corner_global, surf_global, pose12):
        T = _f32(pose12).reshape(-1).copy()
        if corner_global is None and surf_global is None:       # use the maps made resident by set_global_map
            self._ck(self.lib.fbpr_registration(self.h, slot, None, 0, None, 0, MEM_DEVICE, _vp(T)))
        else:
            c = _f32(corner_global).reshape(-1, 4); s = _f32(surf_global).reshape(-1, 4)
            self._ck(self.lib.fbpr_registration(self.h, slot, _vp(c), len(c), _vp(s), len(s), MEM_HOST, _vp(T)))
        return T.reshape(3, 4)

    def run_frames(self, first=0, count=1, with_projection=True, with_features=True):
        self._ck(self.lib.fbpr_run_frames(self.h, first, count, int(with_projection), int(with_features)))

    def run_frames_pipelined(self, first, count, batch_frames=0):
        """fbpr_run_frames_pipelined: whole path for resident frames, front-end of batch k+1 overlapping the LM loop of batch k"""
        self._ck(self.lib.fbpr_run_frames_pipelined(self.h, first, count, int(batch_frames)))

    def use_graphs(self, on=True): self._ck(self.lib.fbpr_use_graphs(self.h, int(on)))
    def sync(self): self._ck(self.lib.fbpr_sync(self.h))
    def stream(self): return self.lib.fbpr_stream(self.h)
    def kernel_launches(self): return int(self.lib.fbpr_kernel_launches(self.h))
    def set_debug_iteration(self, it): self._ck(self.lib.fbpr_set_debug_iteration(self.h, int(it)))

    # ---- results
    def get_results(self, first=0, count=1):
        out = np.zeros(count, RESULT_DTYPE)
        self._ck(self.lib.fbpr_get_results(self.h, first, count, _vp(out), MEM_HOST))
        return out

    def get_results_device(self, first, count, dev_ptr):
        self._ck(self.lib.fbpr_get_results(self.h, first, count, C.c_void_p(dev_ptr), MEM_DEVICE))

    def get_pose(self, slot=0):
        r = self.get_results(slot, 1)[0]
        return r["pose"].copy(), int(r["iters"]), int(r["flags"])

    def get_counts(self, slot=0):
        c = np.zeros(8, np.int32)
        self._ck(self.lib.fbpr_get_counts(self.h, slot, _vp(c)))
        return dict(zip(["n_raw", "n_valid", "n_corner", "n_surf", "n_corner_ds", "n_surf_ds", "n_map_corner", "n_map_surf"], map(int, c)))

    def get_buffer(self, slot, name):
        dt, w = _BUF_DTYPE[name]
        cap = max(self.params.N_SCAN * self.params.Horizon_SCAN, self.params.max_map_corner, self.params.max_map_surf, 65536) * 5 * 4 + 1024
        buf = np.zeros(cap, np.uint8)
        nb = self._ck(self.lib.fbpr_get_buffer(self.h, slot, BUF[name], _vp(buf), cap))
        out = buf[:nb].view(dt).copy()
        return out.reshape(-1, w) if w > 1 else out

    def selftest_smallmat(self, which, rows):
        """fbpr_selftest_smallmat: the device small-matrix routines on `rows` ([n, in_width] f32); which = name below."""
        names = ["JACOBI3", "JACOBI6", "QR6", "LU6", "PLANE5X3", "NOT_DEGENERATE", "QR6_WARP", "SINCOS"]
        out_w = [12, 42, 6, 36, 3, 1, 6, 2]
        k = names.index(which)
        a = _f32(rows).reshape(len(rows), -1)
        out = np.zeros((len(a), out_w[k]), np.float32)
        self._ck(self.lib.fbpr_selftest_smallmat(self.h, k, _vp(a), len(a), _vp(out)))
        return out

    # ---- stand-alone primitives
    def voxel_grid(self, xyzi, leaf):
        p = _f32(xyzi).reshape(-1, 4); n = len(p)
        out = np.zeros((max(n, 1), 4), np.float32); pk = np.zeros(max(n, 1), np.int32); ok = np.zeros(max(n, 1), np.int32)
        m = self._ck(self.lib.fbpr_voxel_grid(self.h, _vp(p), n, C.c_float(leaf), _vp(out), _vp(pk), _vp(ok), MEM_HOST))
        return dict(points=out[:m].copy(), point_keys=pk[:n].copy(), out_keys=ok[:m].copy())

    def knn5(self, map_xyzi, q_xyz, cell=0.25, first_radius=1):
        self._ck(self.lib.fbpr_knn5_first_radius(self.h, int(first_radius)))
        m = _f32(map_xyzi).reshape(-1, 4); q = _f32(q_xyz).reshape(-1, 3); nq = len(q)
        idx = np.zeros((max(nq, 1), 5), np.int32); d2 = np.zeros((max(nq, 1), 5), np.float32)
        self._ck(self.lib.fbpr_knn5(self.h, _vp(m), len(m), C.c_float(cell), _vp(q), nq, _vp(idx), _vp(d2), MEM_HOST))
        return idx[:nq], d2[:nq]

"""CPU-side checks: the C ABI library loads and exports every symbol include/fbpr_b200.h declares (no
compute without a GPU), the product fails loudly without a device, params.yaml parsing, frame sharding
and the world_size-2 gloo gather."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "fbpr_b200.h")).read()
    return sorted(set(re.findall(r"FBPR_API\s+[\w\s\*]+?\b(fbpr_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import feature_base_pointcloud_registration_b200 as fb
    lib = fb.load_library()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fbpr_b200.h but not exported"


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import feature_base_pointcloud_registration_b200 as fb
    with pytest.raises(fb.FbprError):
        fb.Registration(dict(N_SCAN=16, Horizon_SCAN=1800), max_frames=1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "feature_base_pointcloud_registration_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "liboracle" not in src and "oracle/" not in src.replace("the CPU oracle", ""), f


def test_struct_layouts_match_the_header():
    import feature_base_pointcloud_registration_b200 as fb
    assert fb.RAW_POINT_DTYPE.itemsize == 24
    assert fb.RESULT_DTYPE.itemsize == 32
    assert ctypes.sizeof(fb.api.Params) == 25 * 4
    assert ctypes.sizeof(fb.api.FrameInput) == 8 + 4 + 4 + 8 + 8 + 32 + 4 + 4 + 4 + 4 + 8 + 8 + 8 + 8 + 24 + 4 + 4
    assert ctypes.sizeof(fb.api.Pc2Layout) == 8 * 4          # fbpr_pc2_layout: eight int32 fields


def test_params_yaml_reader(tmp_path):
    from feature_base_pointcloud_registration_b200 import load_params_yaml
    y = tmp_path / "params.yaml"
    y.write_text("# Sensor\nN_SCAN: 64   # channels\nHorizon_SCAN: 2048\nedgeThreshold: 1.0\nodometrySurfLeafSize: 0.4\n"
                 "mappingSurfLeafSize: 0.4\nnumberOfCores: 4\nz_tollerance: 1000\nextrinsicRot: [0, 1, 0,\n  -1, 0, 0,\n  0, 0, 1]\n")
    p = load_params_yaml(str(y))
    assert p["N_SCAN"] == 64 and p["Horizon_SCAN"] == 2048 and p["edgeThreshold"] == 1.0
    assert p["odometrySurfLeafSize"] == 0.4 and p["numberOfCores"] == 4 and p["z_tollerance"] == 1000.0
    assert p["surfThreshold"] == 0.1 and p["mappingCornerLeafSize"] == 0.2           # code defaults (utility.h:180-186)


def test_frame_range_partitions_exactly():
    from feature_base_pointcloud_registration_b200.sharding import frame_range
    for total in (0, 1, 7, 1024, 1000):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = frame_range(r, world, total)
                assert 0 <= lo <= hi <= total
                cover += list(range(lo, hi))
            assert cover == list(range(total))
            sizes = [frame_range(r, world, total)[1] - frame_range(r, world, total)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from feature_base_pointcloud_registration_b200.sharding import frame_range, gather_results, RESULT_DTYPE
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
total = 37
lo, hi = frame_range(rank, world, total)
res = np.zeros(hi - lo, RESULT_DTYPE)
for i, f in enumerate(range(lo, hi)):          # stand-in for the per-rank registration results
    res[i]["pose"] = np.arange(6, dtype=np.float32) + f
    res[i]["iters"] = f %% 30
    res[i]["flags"] = 8 if f %% 3 else 12
allres = gather_results(res, dist)
if rank == 0:
    assert len(allres) == total
    for f in range(total):
        assert np.array_equal(allres[f]["pose"], np.arange(6, dtype=np.float32) + f)
        assert allres[f]["iters"] == f %% 30 and allres[f]["flags"] == (8 if f %% 3 else 12)
    print("GATHER_OK")
else:
    assert allres is None
dist.destroy_process_group()
"""


def test_world_size_2_gloo_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % dict(root=ROOT))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GATHER_OK" in out.stdout


# ------------------------------------------------------------------ wire formats and IMU deskew inputs (host side)
def _host_lib():
    import feature_base_pointcloud_registration_b200 as fb
    return ctypes.CDLL(os.path.join(os.path.dirname(fb.__file__), "host", "libfeature_matching_b200.so"))


def _imu_queue(rng, t0, n, hz=500.0, unnormalised=False):
    t = t0 + np.arange(n) / hz
    gyro = rng.normal(0, 0.3, (n, 3))
    q = rng.normal(0, 1, (n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    if unnormalised:
        q *= 1.5
    return np.concatenate([t[:, None], gyro, q], 1)


@pytest.mark.parametrize("case", ["normal", "all_old", "one_sample", "capacity", "unnormalised"])
def test_imu_deskew_info_host_equals_oracle(case):
    """imuDeskewInfo (imageProjection.cpp:323-393): queue pruning, attitude of the last sample before the sweep,
    gyro integration into imuTime / imuRotX,Y,Z, imuPointerCur and the imuAvailable gate -- C++ host vs the oracle."""
    import oracle
    lib = _host_lib()
    rng = np.random.default_rng(7)
    cur, nxt, cap = 100.0, 100.1, 2000
    if case == "normal":
        q = _imu_queue(rng, 99.9, 150)
    elif case == "all_old":
        q = _imu_queue(rng, 90.0, 50)
    elif case == "one_sample":
        q = _imu_queue(rng, 100.2, 5)              # first sample already past timeScanNext + 0.01 -> imuPointerCur 0, unavailable
    elif case == "capacity":
        q = _imu_queue(rng, 99.995, 200, hz=2000.0); cap = 64
    else:
        q = _imu_queue(rng, 99.9, 150, unnormalised=True)
    want = oracle.imu_deskew_info(q, cur, nxt, cap)
    t, rx, ry, rz = (np.zeros(cap) for _ in range(4)); out = np.zeros(5)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    lib.fm_imu_deskew_info.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int] + [ctypes.POINTER(ctypes.c_double)] * 5
    qq = np.ascontiguousarray(q)
    popped = lib.fm_imu_deskew_info(dp(qq), len(qq), cur, nxt, cap, dp(t), dp(rx), dp(ry), dp(rz), dp(out))
    assert popped == want["popped"]
    assert int(out[0]) == want["imuAvailable"]
    if case == "all_old":
        assert want["imuAvailable"] == 0 and popped == len(q)
        return
    assert int(out[1]) == want["imuPointerCur"]
    assert np.float32(out[2]) == want["imuRollInit"] and np.float32(out[3]) == want["imuPitchInit"] and np.float32(out[4]) == want["imuYawInit"]
    n = want["imuPointerCur"] + 1
    for a, k in ((t, "imuTime"), (rx, "imuRotX"), (ry, "imuRotY"), (rz, "imuRotZ")):
        assert np.array_equal(a[:n], want[k][:n])
    if case in ("normal", "unnormalised"):
        # independent check of the tf quaternion -> roll/pitch/yaw convention (fixed axes x, y, z) against scipy
        from scipy.spatial.transform import Rotation
        kept = q[popped:]
        last_before = kept[kept[:, 0] <= cur][-1]
        rpy = Rotation.from_quat(last_before[4:8] / np.linalg.norm(last_before[4:8])).as_euler("xyz")
        assert np.allclose([want["imuRollInit"], want["imuPitchInit"], want["imuYawInit"]], rpy, rtol=0, atol=2e-6)
    if case == "normal":
        assert want["imuAvailable"] == 1 and popped == 45        # stamps < 99.99 are dropped
        # the integrated ramp is the running sum of gyro * dt from the first kept sample
        kept = q[popped:]
        last = np.searchsorted(kept[:, 0], nxt + 0.01, side="right")          # first stamp > timeScanNext + 0.01 breaks the loop
        assert want["imuPointerCur"] == last - 1
        ref = np.concatenate([[0.0], np.cumsum(kept[1:last, 1] * np.diff(kept[:last, 0]))])
        assert np.allclose(want["imuRotX"][:last], ref, rtol=0, atol=1e-12)
    if case == "one_sample":
        assert want["imuAvailable"] == 0 and want["imuPointerCur"] == -1


def test_pcd_io_roundtrip_and_pcl_layout(tmp_path):
    """PCD map IO (pcl::io::loadPCDFile / savePCDFileASCII, mapOptmization.h:247-257, :495-519): binary round trip is
    bit-exact, ASCII keeps 8 significant digits (PCL's precision) and a file in PCL's own header layout (extra fields,
    COUNT, VIEWPOINT) is read field by field."""
    lib = _host_lib()
    rng = np.random.default_rng(3)
    pts = np.concatenate([rng.uniform(-50, 50, (1000, 3)), rng.uniform(0, 255, (1000, 1))], 1).astype(np.float32)
    fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    for mode in (0, 1):
        path = str(tmp_path / f"m{mode}.pcd").encode()
        assert lib.fm_save_pcd(path, fp(pts), len(pts), mode) == 0
        got = np.zeros_like(pts); err = ctypes.create_string_buffer(256)
        n = lib.fm_load_pcd(path, fp(got), len(got), err, 256)
        assert n == len(pts), err.value
        if mode == 1:
            assert np.array_equal(got, pts)
        else:
            assert np.allclose(got, pts, rtol=2e-8 * 10, atol=0)          # %.8g
            head = open(path.decode()).read().split("DATA ascii")[0]
            assert "FIELDS x y z intensity" in head and "SIZE 4 4 4 4" in head and "TYPE F F F F" in head and f"POINTS {len(pts)}" in head
    # a PCL-written ASCII file with a different field set (PointXYZINormal-like) and CRLF line ends
    body = "\r\n".join(["# .PCD v0.7 - Point Cloud Data file format", "VERSION 0.7", "FIELDS x y z intensity normal_x curvature", "SIZE 4 4 4 4 4 4",
                        "TYPE F F F F F F", "COUNT 1 1 1 1 3 1", "WIDTH 2", "HEIGHT 1", "VIEWPOINT 0 0 0 1 0 0 0", "POINTS 2", "DATA ascii",
                        "1.5 -2.25 3 7 0 0 1 0.5", "nan 4 5e-1 9 1 0 0 0.25", ""])
    f = tmp_path / "pcl.pcd"; f.write_bytes(body.encode())
    got = np.zeros((2, 4), np.float32); err = ctypes.create_string_buffer(256)
    assert lib.fm_load_pcd(str(f).encode(), fp(got), 2, err, 256) == 2, err.value
    assert np.array_equal(got[0], np.float32([1.5, -2.25, 3, 7])) and np.isnan(got[1, 0]) and np.array_equal(got[1, 1:], np.float32([4, 0.5, 9]))
    # errors are reported, not swallowed
    bad = tmp_path / "bad.pcd"; bad.write_text("FIELDS x y\nSIZE 4 4\nTYPE F F\nCOUNT 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA ascii\n1 2\n")
    assert lib.fm_load_pcd(str(bad).encode(), fp(got), 2, err, 256) == -1 and b"x y z" in err.value
    assert lib.fm_load_pcd(str(tmp_path / "missing.pcd").encode(), fp(got), 2, err, 256) == -1


def test_oracle_extract_nearby_semantics():
    """extractNearby (mapOptmization.h:872-907) in the oracle: radius hits sorted by distance, VoxelGrid averages the key
    index with the position, the last-10 s poses are appended newest first, and extractCloud truncates the averaged index."""
    import oracle
    import synth
    P = dict(synth.params_for(1)); P["surroundingKeyframeSearchRadius"] = 10.0
    # key poses on a line, 0.6 m apart; density 2 m => ~3 poses per voxel; the far ones (> 10 m) are outside the radius
    n = 30
    poses = np.zeros((n, 6), np.float32); poses[:, 3] = np.arange(n)[::-1] * 0.6
    times = np.arange(n) * 1.0                                   # 1 s apart; last 10 s = keys 20..29 (time_last = 29.5)
    cf = [np.float32([[0, 0, 0, k]]) for k in range(n)]; sf = [np.float32([[1, 0, 0, k]]) for k in range(n)]
    mo = oracle.MapOptimization(P)
    ds, counts = mo.extract_surrounding(poses, times, 2.0, 29.5, cf, sf)
    inside = np.where(poses[:, 3] ** 2 < 100.0)[0]              # strict
    vox = oracle.voxel_grid(np.concatenate([poses[inside][::-1][:, 3:], inside[::-1, None].astype(np.float32)], 1), 2.0)["points"]
    assert np.array_equal(ds[:len(vox)], vox)                    # hits ascending by distance = keys 29, 28, ...
    recent = ds[len(vox):]
    assert np.array_equal(recent[:, 3], np.arange(29, 19, -1, dtype=np.float32))      # 29.5 - t < 10 -> t > 19.5
    assert np.any(ds[:len(vox), 3] != np.floor(ds[:len(vox), 3]))                      # averaged (non-integer) indices exist
    assert counts[0] == len(ds) and counts[1] == len(ds)        # one corner + one surf point per selected entry, none re-check-rejected

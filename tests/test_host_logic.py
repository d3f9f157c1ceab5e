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
    assert ctypes.sizeof(fb.api.FrameInput) == 8 + 4 + 4 + 8 + 8 + 32 + 4 + 4 + 4 + 4 + 8 + 8 + 8 + 8 + 24


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

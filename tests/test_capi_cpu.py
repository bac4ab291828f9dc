"""CPU-side checks of the boundary (no GPU needed): the C-ABI library builds, loads, exports every
symbol include/nmpm.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import nuclearmpm_b200 as nm
from nuclearmpm_b200 import build as nbuild
from oracle import cpu_oracle as co

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    nbuild.build()
    return nm.load_library()


def declared_symbols():
    text = (ROOT / "include" / "nmpm.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nmpm_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 30
    raw = ctypes.CDLL(str(nm.lib_path()))
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, f"declared in include/nmpm.h but not exported: {missing}"
    assert b"sm_100a" in lib.nmpm_build_info()


def test_cubin_is_sm100a_only():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", str(nm.lib_path())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback(lib):
    """Without a GPU every compute entry point must fail loudly, never silently compute on the host."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present; the negative path is checked on the CPU box")
    with pytest.raises(nm.NmpmError, match="no CUDA device"):
        nm.MPMSimulation(nm.cube(2, 5, 0.4, 0.6), nm.MaterialModel.kSnow)
    with pytest.raises(nm.NmpmError):
        nm.svd_batch(np.eye(3, dtype=np.float32)[None])


def test_invalid_arguments_rejected(lib):
    h = ctypes.c_void_p()
    x = np.zeros((4, 2), np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    args = [x.ctypes.data_as(fp)] + [None] * 6
    assert lib.nmpm_create(4, 0, 64, 1e-4, 1e4, 0.2, -100.0, 4, *args, None, ctypes.byref(h)) == 1  # dim
    assert lib.nmpm_create(2, 7, 64, 1e-4, 1e4, 0.2, -100.0, 4, *args, None, ctypes.byref(h)) == 1  # model
    assert lib.nmpm_create_aos(2, 0, 64, 1e-4, 1e4, 0.2, -100.0, 4, x.ctypes.data, 8, None, ctypes.byref(h)) == 1
    assert lib.nmpm_advance(None, 1) == 1
    assert lib.nmpm_num_particles(None) == 0
    # batches (nmpm_create_batch): nscenes >= 1, counts / E / nu required
    counts = (ctypes.c_size_t * 2)(2, 2)
    E = np.array([1e3, 2e3], np.float32).ctypes.data_as(fp)
    assert lib.nmpm_create_batch(1, 64, 1e-4, -100.0, 0, counts, E, E, *args, None, ctypes.byref(h)) == 1
    assert lib.nmpm_create_batch(1, 64, 1e-4, -100.0, 2, None, E, E, *args, None, ctypes.byref(h)) == 1
    assert b"batch" in lib.nmpm_last_error(None)
    assert lib.nmpm_num_scenes(None) == 0 and lib.nmpm_fused(None) == 0 and lib.nmpm_tiles_active(None) == 0


def test_batch_has_no_cpu_fallback_either(lib):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present; the negative path is checked on the CPU box")
    with pytest.raises(nm.NmpmError, match="no CUDA device"):
        nm.MPMBatch([nm.cube(2, 5, 0.4, 0.6), nm.cube(2, 5, 0.2, 0.4)], nm.MaterialModel.kJelly, E=[1e3, 2e3], nu=[0.3, 0.3])


def test_python_cube_matches_oracle_bitwise():
    """Scene generator: nm.cube must reproduce nclr::cube / Eigen LinSpaced."""
    for args in [(2, 50, 0.4, 0.6), (3, 16, 0.375, 0.625), (2, 7, -0.9, 0.3), (3, 1, 0.5, 0.7), (2, 25, 0.1, 0.3),
                 (3, 9, 0.25, 0.5), (2, 126, 0.05, 0.05 + 62.5 / 256)]:
        a, b = nm.cube(*args), co.cube(*args)
        assert a.shape == b.shape and (a.view(np.uint32) == b.view(np.uint32)).all(), args


def test_material_enum_and_constants():
    assert [int(m) for m in nm.MaterialModel] == [0, 1, 2]  # src/nclr.h:57-61
    assert nm.MPMSimulation.kBoundary == 3 and nm.MPMSimulation.kSnowHardening == 10.0
    assert nm.MPMSimulation.kJellyHardening == 0.3 and nm.MPMSimulation.kLiquidHardening == 1.0

"""GPU test of the drop-in C++ header (include/nclr.h): a C++ caller written like the reference's
src/example.cpp / src/solver.cpp gets the oracle's particles() and grid() back."""
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import cpu_oracle as co
from test_dropin_cpu import ROOT, _build

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,model,res,cres,steps", [(2, co.SNOW, 64, 20, 3), (2, co.LIQUID, 64, 20, 3),
                                                      (3, co.JELLY, 32, 8, 3), (3, co.SNOW, 32, 8, 1)])
def test_cpp_caller_matches_oracle(tmp_path, dim, model, res, cres, steps):
    exe = _build(tmp_path, ROOT / "tests" / "cpp" / "dropin_main.cpp", "dropin_main")
    out = tmp_path / "state.bin"
    r = subprocess.run([str(exe), str(dim), str(model), str(res), str(cres), "0.4", "0.6", str(steps), str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "out_of_range ok" in r.stdout and "positions ok" in r.stdout
    raw = out.read_bytes()
    n, psize, cells, csize = struct.unpack("4Q", raw[:32])
    assert psize == (64 if dim == 2 else 112) and csize == 4 * (dim + 1)
    assert n == cres ** dim and cells == (res + 1) ** dim
    rec = np.frombuffer(raw, np.float32, n * psize // 4, 32).reshape(n, psize // 4)
    grid = np.frombuffer(raw, np.float32, cells * (dim + 1), 32 + n * psize).reshape(cells, dim + 1)
    d, dd = dim, dim * dim
    got = dict(x=rec[:, :d], v=rec[:, d:2 * d], F=rec[:, 2 * d:2 * d + dd], C=rec[:, 2 * d + dd:2 * d + 2 * dd],
               Jp=rec[:, 2 * d + 2 * dd])
    assert (rec[:, 2 * d + 2 * dd + 1] == 1).all() and (rec[:, 2 * d + 2 * dd + 2] == 1).all()      # mass, volume kept
    assert (rec[:, 2 * d + 2 * dd + 3].view(np.int32) == 0xED553B).all()                           # colour kept

    x0 = co.cube(dim, cres, 0.4, 0.6)
    cpu = co.CpuSim(x0, model, res)
    cpu.advance(steps)
    ref = cpu.particles()
    gv, gm = cpu.grid()
    mu0, lam0 = cpu.lame()
    assert f"mu_0 {mu0:.9g} lambda_0 {lam0:.9g}" in r.stdout
    vmax = max(1.0, float(np.abs(ref["v"]).max()))
    cmax = max(1.0, float(np.abs(ref["C"]).max()))
    k = steps  # free-running: tolerances scale with the number of steps
    assert np.abs(got["x"] - ref["x"]).max() <= 2.4e-7 * k
    assert np.abs(got["v"] - ref["v"]).max() <= 1e-5 * vmax * k
    assert np.abs(got["F"] - ref["F"].reshape(n, dd)).max() <= 2e-5 * k
    assert np.abs(got["C"] - ref["C"].reshape(n, dd)).max() <= 5e-5 * cmax * k
    assert np.abs(got["Jp"] - ref["Jp"]).max() <= 1e-4 * k
    assert np.abs(grid[:, :dim] - gv).max() <= 3e-5 * vmax * k
    assert np.abs(grid[:, dim] - gm).max() <= 1e-5 * max(1.0, float(gm.max())) * k

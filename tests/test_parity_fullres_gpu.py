"""Parity against the REFERENCE at the BASELINE configs' own grid resolution (res 128 / 256 / 512).

The golden scenes of test_parity_gpu.py live on 32^3 / 64^2 grids.  north_star asks for "results within tolerance
of the reference on every config", and the stress term grows with the resolution (dt*vol*4/dx^2*2mu, src/nclr.h:
325-336), so the one-step tolerances are re-established here where the bench runs: sub-blocks of cfg2 / cfg3 / cfg4
(same spacing, material, dt and grid as the config; particle counts the CPU reference steps through in seconds)
are advanced by oracle/_ref — the UNMODIFIED reference header (`ref_strict`) when it was built, else the C port that
is bit-exact with it — and at steps {1, 2, 3, 30} the reference state is uploaded and ONE step is compared
(teacher-forced: 3D snow is chaotic from step 4): cell keys / sort permutation bit-exact, post-P2G
grid, x, v, F, C, Jp.

Tolerances.  x, F, Jp, C keep the golden-scene values.  v and the grid momentum get an additive term for the
rounding of the polar factor R (a few ulp of fp32, whatever the implementation — the reference's own -Ofast and
strict builds differ by as much) amplified by the stress prefactor:
    amp = dt * volume * (4/dx^2) * 2 mu_0 e(Jp) * dx      velocity change per unit error of (F - R), mass = 1
    tol_v = 1e-5 max(1, |v|max) + 8 eps_f32 * amp
At res 512 snow (mu_0 e^4 ~ 2.3e5) amp ~ 1e5, i.e. tol_v ~ 0.1 = 5e-3 of the clamp speed 17.6; at res 128 jelly
amp ~ 20.  The measured errors are written to gpurun_out/parity_fullres.md (committed as profiles/r02_parity_fullres.md).
"""
import os
from pathlib import Path

import numpy as np
import pytest

import nuclearmpm_b200 as nm
from oracle import cpu_oracle as co

pytestmark = pytest.mark.gpu
FIELDS = ("x", "v", "F", "C", "Jp")
EPS = float(np.finfo(np.float32).eps)
REPORT = Path(__file__).resolve().parents[1] / "gpurun_out" / "parity_fullres.md"


def _kind():
    return "ref_strict" if co.available("ref_strict") else "port"


def _sub_block(workload, m):
    """m^3 particles of the config's block, same spacing, starting at the block's min corner."""
    if workload == "cfg2":   # cube<3>(64, 0.375, 0.625), jelly, res 128
        return nm.cube(3, m, 0.375, 0.375 + (m - 1) * (0.25 / 63)), co.JELLY, 128
    if workload == "cfg3":   # 126^3 liquid block, res 256
        return nm.cube(3, m, 0.05, 0.05 + (m - 1) * (62.5 / 256 / 125)), co.LIQUID, 256
    if workload == "cfg4":   # cube<3>(256, 0.25, 0.5), snow, res 512
        return nm.cube(3, m, 0.25, 0.25 + (m - 1) * (0.25 / 255)), co.SNOW, 512
    raise ValueError(workload)


def _amp(model, res, jp_min, dt=1e-4, E=1e4, nu=0.2):
    mu0 = E / (2 * (1 + nu))
    e = {co.SNOW: float(np.exp(10.0 * (1.0 - jp_min))), co.JELLY: 0.3, co.LIQUID: 1.0}[model]
    dx = 1.0 / res
    return dt * 1.0 * (4.0 / dx ** 2) * 2.0 * mu0 * e * dx


def _errors(got, ref):
    return dict(x=float(np.abs(got["x"] - ref["x"]).max()), v=float(np.abs(got["v"] - ref["v"]).max()),
                C=float(np.abs(got["C"] - ref["C"]).max()), F=float(np.abs(got["F"] - ref["F"]).max()),
                Jp=float(np.abs(got["Jp"] - ref["Jp"]).max()))


def _log(line):
    REPORT.parent.mkdir(parents=True, exist_ok=True)
    new = not REPORT.exists()
    with open(REPORT, "a") as f:
        if new:
            f.write("| case | step | x err / tol | v | F | C | Jp | grid momentum | momentum drift gpu / reference | scales | "
                    "reference -Ofast vs strict on the same step |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
        f.write(line + "\n")


CASES = [
    # workload, particles per axis, p2g_variant (0 = what the library picks), reference steps to teacher-force from
    ("cfg2", 64, 0, (0, 1, 2, 29)),    # the whole of cfg2 (262 144 p)
    ("cfg3", 64, 0, (0, 1, 2, 29)),
    ("cfg4", 64, 0, (0, 1, 2, 29)),
    ("cfg4", 128, 44, (0, 1, 2)),      # 2.1 M p through the three-stream P2G with 4 chunks per warp (the cfg4 default path)
    ("cfg4", 128, 42, (2,)),
]


@pytest.mark.parametrize("workload,m,variant,from_steps", CASES)
def test_one_step_against_reference_at_config_resolution(workload, m, variant, from_steps):
    x, model, res = _sub_block(workload, m)
    kind = _kind()
    cpu = co.CpuSim(x, model, res, kind=kind)
    gpu = nm.MPMSimulation(x, model, res, p2g_variant=variant)
    n = len(x)
    cells = (res + 1) ** 3
    done = 0
    for s in from_steps:
        cpu.advance(s - done)
        done = s
        ref0 = cpu.particles()
        gpu.upload(*[ref0[k] for k in FIELDS])
        # --- integer work: cell keys and the stable radix order, bit-exact ---------------------------------
        d = gpu.sort_debug()
        base, keys, bad = co.cell_keys(ref0["x"], res, mode=1, tb=d["tile_bits"])
        assert bad == 0
        assert (d["base"] == base).all() and (d["keys"] == keys).all()
        perm = np.argsort(keys, kind="stable").astype(np.uint32)
        assert (d["perm"] == perm).all()
        del d, base, keys, perm
        # --- P2G: post-P2G grid (momentum, mass) against the reference's ---------------------------------------
        cpu.phase(0)
        gpu.phase(0)
        rgv, rgm = cpu.grid()
        ggv, ggm = gpu.grid()
        assert ggm.shape == (cells,)
        jp_min = float(ref0["Jp"].min())
        amp = _amp(model, res, jp_min)
        mom_max = max(1.0, float(np.abs(rgv).max()))
        # momentum of a node: sum of <= ~64 particle terms, each with the amplified rounding of R
        tol_mom = 3e-5 * mom_max + 8 * EPS * amp * max(1.0, float(rgm.max()))
        e_mom = float(np.abs(ggv - rgv).max())
        e_mass = float(np.abs(ggm - rgm).max() / max(1e-30, float(rgm.max())))
        assert e_mom <= tol_mom and e_mass <= 1e-5, (workload, m, s, e_mom, tol_mom, e_mass)
        assert np.isclose(ggm.astype(np.float64).sum(), float(n), rtol=1e-6)       # mass conservation (pre grid_op, Q6)
        # momentum conservation: the stress terms of a particle cancel over its 27 nodes only up to their rounding
        # (~eps * amp each), a random walk over n particles — the reference's own grid shows the same drift (logged)
        mom_p = ref0["v"].astype(np.float64).sum(0)
        drift_gpu = float(np.abs(ggv.astype(np.float64).sum(0) - mom_p).max())
        drift_ref = float(np.abs(rgv.astype(np.float64).sum(0) - mom_p).max())
        assert drift_gpu <= 1e-4 * float(np.abs(mom_p).max()) + 1e-6 * n + 32 * EPS * amp * np.sqrt(n), (drift_gpu, drift_ref)
        touched = int(np.count_nonzero(rgm))
        assert np.count_nonzero(ggm) == touched                                      # same set of nodes written
        del rgv, rgm, ggv, ggm
        # --- grid_op + G2P: particle state after the step -------------------------------------------------
        cpu.phase(1), cpu.phase(2)
        gpu.phase(1), gpu.phase(2)
        done += 1
        ref1 = cpu.particles()
        got = gpu.particles()
        err = _errors(got, ref1)
        # the reference's own FP-mode noise on the same step (its CMake build is -Ofast; the oracle is strict FP)
        noise = None
        if kind == "ref_strict" and co.available("ref_fast") and n <= 300_000:
            fast = co.CpuSim(ref0["x"], model, res, v=ref0["v"], F=ref0["F"], Cm=ref0["C"], Jp=ref0["Jp"], kind="ref_fast")
            fast.advance(1)
            noise = _errors(fast.particles(), ref1)
            del fast
        vmax = max(1.0, float(np.abs(ref1["v"]).max()))
        cmax = max(1.0, float(np.abs(ref1["C"]).max()))
        tol = dict(x=2.4e-7 + 1e-4 * 8 * EPS * amp,   # x += dt v
                   v=1e-5 * vmax + 8 * EPS * amp,
                   C=5e-5 * cmax + 4 * res * 8 * EPS * amp,   # C = 4/dx * sum w v dpos^T
                   F=2e-5 + 1e-4 * 4 * res * 8 * EPS * amp,   # F' = (I + dt C) F
                   Jp=1e-4 + 1e-4 * 4 * res * 8 * EPS * amp)
        _log(f"| {workload} {m}^3 res {res} variant {variant} ({kind}) | {s}->{s + 1} | "
             + " | ".join(f"{err[k]:.2e} / {tol[k]:.2e}" for k in ("x", "v", "F", "C", "Jp"))
             + f" | {e_mom:.2e} / {tol_mom:.2e} | {drift_gpu:.3g} / {drift_ref:.3g} | |v|max {vmax:.3g} amp {amp:.3g} | "
             + ("-" if noise is None else " ".join(f"{k} {noise[k]:.1e}" for k in FIELDS)) + " |")
        bad = {k: (err[k], tol[k]) for k in FIELDS if not err[k] <= tol[k]}
        assert not bad, (workload, m, variant, s, bad)

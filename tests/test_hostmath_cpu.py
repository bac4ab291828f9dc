"""The device math header (csrc/nmpm_math.cuh) instantiated for the HOST and checked against the oracle:
lets the CPU suite validate the register algorithms (one-sided Jacobi recompose -> polar rotation and
snow projection, reference-shaped two-sided Jacobi SVD) without a GPU.  The same cases run on the device
through the C-ABI batch hooks in tests/test_parity_gpu.py."""
import ctypes as ct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from mathcases import TOL, matrices, oracle_reference
from oracle import cpu_oracle as co

ROOT = Path(__file__).resolve().parents[1]
fp = ct.POINTER(ct.c_float)


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostmath") / "libhostmath.so"
    subprocess.run(["g++", "-std=c++17", "-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared",
                    "-I/usr/local/cuda/include", "-o", str(out), str(ROOT / "tests" / "cpp" / "hostmath.cpp")],
                   check=True, capture_output=True)
    L = ct.CDLL(str(out))
    L.hm_polar3.restype = ct.c_long
    L.hm_snow_project3.restype = ct.c_long
    L.hm_polar3.argtypes = [fp, fp, ct.c_long]
    L.hm_snow_project3.argtypes = [fp, ct.c_float, ct.c_float, fp, ct.c_long]
    L.hm_svd3.argtypes = [fp, fp, fp, fp, ct.c_long]
    L.hm_affine3.argtypes = [ct.c_int, fp, fp, fp, fp, fp] + [ct.c_float] * 4 + [fp, ct.c_long]
    return L


def P(a):
    return a.ctypes.data_as(fp)


@pytest.mark.parametrize("name", sorted(TOL))
def test_fast_polar_and_snow_projection_match_oracle(hm, name):
    A = matrices()[name]
    n = A.shape[0]
    R, G = np.empty_like(A), np.empty_like(A)
    fast = hm.hm_polar3(P(A), P(R), n)
    hm.hm_snow_project3(P(A), ct.c_float(0.975), ct.c_float(1.0045), P(G), n)
    Ro, Go, well = oracle_reference(co, A)
    assert np.isfinite(R).all() and np.isfinite(G).all()
    assert well.sum() >= (0 if name == "liquid_diag" else n // 2)
    eR = np.abs(R - Ro)[well].max() if well.any() else 0.0
    eG = np.abs(G - Go)[well].max() if well.any() else 0.0
    assert eR <= TOL[name] and eG <= TOL[name], (name, eR, eG)
    # orthogonality and orientation of R where it is defined
    Rm = R[well].transpose(0, 2, 1).astype(np.float64)
    if len(Rm):
        assert np.abs(Rm @ Rm.transpose(0, 2, 1) - np.eye(3)).max() <= 5e-6
        assert (np.linalg.det(Rm) > 0.99).all() or name in ("general", "scaled")  # det R = sign(det A) in general
    newton, hestenes = divmod(fast, 1000000)
    if name != "liquid_diag":
        assert newton + hestenes == n  # the fast paths carry these regimes (rank-1 inputs fall back to the Jacobi SVD)
    if name == "snow_F":
        assert newton == n  # snow's F is always within Newton's basin (singular values clamped to [0.975, 1.0045])
    if name in ("jelly_rank2", "step0"):
        assert newton == 0  # singular input must never pass the Newton acceptance test


def test_reference_shaped_svd_matches_golden_bitwise_or_tight(hm):
    g = dict(np.load(ROOT / "tests" / "golden" / "svd_3d.npz"))
    A = np.ascontiguousarray(g["A"], np.float32)
    n = A.shape[0]
    U, V, S = np.empty_like(A), np.empty_like(A), np.empty((n, 3), np.float32)
    hm.hm_svd3(P(A), P(U), P(S), P(V), n)
    for k in range(n):
        a = A[k].T.astype(np.float64)
        u, v, s = U[k].T.astype(np.float64), V[k].T.astype(np.float64), S[k].astype(np.float64)
        scale = max(1.0, np.abs(a).max())
        assert np.abs((u * s) @ v.T - a).max() <= 3e-6 * scale
        assert np.abs(s - np.diag(g["S"][k].T)).max() <= 3e-6 * scale


@pytest.mark.parametrize("scene", ["random_3d_snow", "random_3d_jelly", "random_3d_liquid"])
def test_affine_matrix_matches_reference_golden(hm, scene):
    """first_piola_kirchoff_stress + mass*C (src/nclr.h:313-337) as the device computes it (Newton / one-sided
    Jacobi polar, fp32 hardening exponent) against the value the reference header produced (golden `affine0`)."""
    g = dict(np.load(ROOT / "tests" / "golden" / f"{scene}.npz"))
    F, C, Jp = (np.ascontiguousarray(g[k], np.float32) for k in ("F0", "C0", "Jp0"))
    n = len(Jp)
    mass, volume = (np.ascontiguousarray(g[k], np.float32) for k in ("mass", "volume"))
    E, nu, res = float(g["E"]), float(g["nu"]), int(g["res"])
    mu0 = np.float32(E) / (np.float32(2) * (np.float32(1) + np.float32(nu)))
    lam0 = np.float32(E) * np.float32(nu) / ((np.float32(1) + np.float32(nu)) * (np.float32(1) - np.float32(2) * np.float32(nu)))
    dx = np.float32(1.0 / res)
    A = np.empty_like(F)
    hm.hm_affine3(int(g["model"]), P(F), P(C), P(Jp), P(mass), P(volume),
                  ct.c_float(float(mu0)), ct.c_float(float(lam0)), ct.c_float(float(g["dt"])),
                  ct.c_float(float(np.float32(1) / dx)), P(A), n)
    ref = g["affine0"]
    scale = max(1.0, float(np.abs(ref).max()))
    assert np.abs(A - ref).max() <= 2e-5 * scale, (scene, np.abs(A - ref).max(), scale)

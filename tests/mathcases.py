"""Shared 3x3 test matrices for the register math (polar rotation, snow projection): the regimes the
solver actually meets (SURVEY.md §7 'hard parts'): well-conditioned snow F, the Q1 F' with an O(dt C)
third row, rank-2 jelly F, diagonal rank-1 liquid F, general matrices with either sign of det."""
import numpy as np


def _rot(rng, n):
    q, _ = np.linalg.qr(rng.normal(size=(n, 3, 3)))
    q[:, :, 0] *= np.sign(np.linalg.det(q))[:, None]
    return q


def matrices(n=400, seed=0):
    """-> dict name -> (n,3,3) float32 in the interchange layout (per matrix column-major: A[k, j, i] = A(i,j))."""
    rng = np.random.default_rng(seed)
    U, V = _rot(rng, n), _rot(rng, n)
    sig = rng.uniform(0.975, 1.0045, size=(n, 3))
    F = np.einsum("nij,nj,nkj->nik", U, sig, V)
    C = rng.normal(size=(n, 3, 3)) * rng.choice([1, 100, 3000, 15000], size=(n, 1, 1))
    out = {"snow_F": F, "snow_Fprime_q1": (np.diag([1, 1, 0.0]) + 1e-4 * C) @ F,
           "snow_Fprime_phys": (np.eye(3) + 1e-4 * C) @ F}
    J = F.copy()
    J[:, :, 2] = 0
    out["jelly_rank2"] = J
    Jl = np.zeros((n, 3, 3))
    Jl[:, 1, 1] = 1
    Jl[:, 0, 0] = rng.choice([0, 0, 1, 0.5, 2], size=n)
    out["liquid_diag"] = Jl
    out["general"] = rng.normal(size=(n, 3, 3))
    out["step0"] = np.tile(np.diag([1, 1, 0.0]), (n, 1, 1))
    out["scaled"] = rng.normal(size=(n, 3, 3)) * 10.0 ** rng.uniform(-6, 6, size=(n, 1, 1))
    return {k: np.ascontiguousarray(v.transpose(0, 2, 1).astype(np.float32)) for k, v in out.items()}


def oracle_reference(co, A, lo=0.975, hi=1.0045):
    """Per matrix: oracle polar R, oracle U clamp(sig) V^T (float64 recomposition of the oracle's float SVD),
    and whether the polar factor is well defined (numerical rank >= 2)."""
    n = A.shape[0]
    R = np.empty_like(A)
    G = np.empty_like(A)
    well = np.zeros(n, bool)
    for i in range(n):
        Uo, So, Vo = co.svd(A[i])
        Ro, _ = co.polar(A[i])
        u, s, v = Uo.T.astype(np.float64), np.diag(So.T).astype(np.float64), Vo.T.astype(np.float64)
        sc = np.clip(s, np.float32(lo), np.float32(hi))
        G[i] = ((u * sc) @ v.T).T
        R[i] = Ro
        smax = max(np.abs(s).max(), 1e-300)
        well[i] = (np.abs(s) > 1e-4 * smax).sum() >= 2
    return R, G, well


# absolute error bounds, relative to the conditioning of the case (measured: profiles/r01b_math.md)
TOL = {"snow_F": 4e-6, "snow_Fprime_q1": 1e-5, "snow_Fprime_phys": 2e-4, "jelly_rank2": 4e-6, "liquid_diag": 1e-6,
       "general": 5e-5, "step0": 1e-6, "scaled": 5e-5}

"""Generate the committed golden vectors from the REFERENCE ITSELF.

Runs only in the authoring container (needs /root/reference): oracle/_ref/libnclr_ref_strict.so is
the unmodified /root/reference/src/nclr.h built against oracle/eigen_standin with strict FP
(`make -C oracle ref`).  The reference ships no golden vectors of its own, so these
files are the pin for the oracle port (tests/test_oracle.py) and for the CUDA path (tests/test_parity_gpu.py).

    python tests/golden/make_golden.py

Outputs (np.savez_compressed, float32 unless noted):
  scene_{dim}d_{model}.npz   free-running trajectory: x0 + state after steps {1,2,3,10,100};
                             post-P2G (pre-grid_op) and post-grid_op grids of step 1 and step 101
                             stored sparsely (node ids + values)
  random_{dim}d_{model}.npz  one step from a random (v, F, C, Jp, mass, volume) state
  svd_{dim}d.npz             nclr_svd / nclr_polar inputs and outputs (incl. rank-deficient, Q1)
  cube.npz                   cube<dim>() point sets of the benchmark scenes (first/last points + checksums)
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import cpu_oracle as co  # noqa: E402

OUT = Path(__file__).resolve().parent
STEPS = (1, 2, 3, 10, 100)
KIND = "ref_strict"


def scene(dim: int) -> tuple[np.ndarray, int]:
    if dim == 2:  # config 5 shape: two 25x25 squares on the diagonal (Q15), res 64
        a = co.cube(2, 25, 0.40, 0.60, KIND)
        b = co.cube(2, 25, 0.10, 0.30, KIND)
        return np.concatenate([a, b]).astype(np.float32), 64
    return co.cube(3, 10, 0.42, 0.58, KIND), 32  # 1000 particles over ~5 cells/axis, res 32


def sparse_grid(gv: np.ndarray, gm: np.ndarray) -> dict:
    nz = np.flatnonzero((gm != 0) | (gv != 0).any(axis=1)).astype(np.int32)
    return dict(ids=nz, v=gv[nz], m=gm[nz])


def gen_scene(dim: int, model: int) -> None:
    x0, res = scene(dim)
    sim = co.CpuSim(x0, model, res, kind=KIND)
    out = dict(x0=x0, res=np.int32(res), model=np.int32(model), steps=np.array(STEPS, np.int32))
    step = 0
    for target in STEPS:
        while step < target:
            if step == 0:  # phase-by-phase on step 1 to capture both grid stages (step 101: below)
                sim.phase(0)
                g = sparse_grid(*sim.grid())
                for k, val in g.items():
                    out[f"p2g{step + 1}_{k}"] = val
                sim.phase(1)
                g = sparse_grid(*sim.grid())
                for k, val in g.items():
                    out[f"gop{step + 1}_{k}"] = val
                sim.phase(2)
            else:
                sim.advance(1)
            step += 1
        for k, val in sim.particles().items():
            out[f"s{target}_{k}"] = val
    # one more phase-split step from the step-100 state (teacher-forced target): state after 101
    sim.phase(0)
    for k, val in sparse_grid(*sim.grid()).items():
        out[f"p2g101_{k}"] = val
    sim.phase(1)
    for k, val in sparse_grid(*sim.grid()).items():
        out[f"gop101_{k}"] = val
    sim.phase(2)
    for k, val in sim.particles().items():
        out[f"s101_{k}"] = val
    np.savez_compressed(OUT / f"scene_{dim}d_{co.MODEL_NAMES[model]}.npz", **out)


def gen_random(dim: int, model: int) -> None:
    rng = np.random.default_rng(1000 + 10 * dim + model)
    n, res = 600, 32
    x = rng.uniform(0.25, 0.75, (n, dim)).astype(np.float32)
    v = rng.normal(0, 2.0, (n, dim)).astype(np.float32)
    eye = np.eye(dim, dtype=np.float32)
    F = (eye + rng.normal(0, 0.05, (n, dim, dim))).astype(np.float32)
    Cm = rng.normal(0, 20.0, (n, dim, dim)).astype(np.float32)
    Jp = rng.uniform(0.8, 1.2, n).astype(np.float32)
    mass = rng.uniform(0.5, 2.0, n).astype(np.float32)
    vol = rng.uniform(0.5, 2.0, n).astype(np.float32)
    E, nu, g = 3000.0, 0.3, -50.0
    sim = co.CpuSim(x, model, res, 1e-4, E, nu, g, v=v, F=F, Cm=Cm, Jp=Jp, mass=mass, volume=vol, kind=KIND)
    out = dict(x0=x, v0=v, F0=F, C0=Cm, Jp0=Jp, mass=mass, volume=vol, res=np.int32(res), model=np.int32(model),
               E=np.float32(E), nu=np.float32(nu), gravity=np.float32(g), dt=np.float32(1e-4))
    out["affine0"] = np.stack([sim.affine(p) for p in range(n)])
    sim.phase(0)
    for k, val in sparse_grid(*sim.grid()).items():
        out[f"p2g_{k}"] = val
    sim.phase(1)
    for k, val in sparse_grid(*sim.grid()).items():
        out[f"gop_{k}"] = val
    sim.phase(2)
    for k, val in sim.particles().items():
        out[f"s1_{k}"] = val
    np.savez_compressed(OUT / f"random_{dim}d_{co.MODEL_NAMES[model]}.npz", **out)


def gen_svd(dim: int) -> None:
    rng = np.random.default_rng(77 + dim)
    mats = []
    eye = np.eye(dim, dtype=np.float32)
    for _ in range(200):
        mats.append(eye + rng.normal(0, 0.02, (dim, dim)))  # near identity (the common case)
    for _ in range(100):
        mats.append(rng.normal(0, 1.0, (dim, dim)))  # general, both determinant signs
    if dim == 3:
        q1 = np.diag([1, 1, 0]).astype(np.float32)  # Q1: singular "identity"
        mats.append(q1)
        for _ in range(100):  # rank 2: third column exactly zero (3D jelly/liquid forever)
            m = (eye + rng.normal(0, 0.05, (3, 3))) @ q1
            mats.append(m)
        for _ in range(20):  # rank 1 (3D liquid after step 1)
            mats.append(np.diag([rng.uniform(0, 1e-3), 1.0, 0.0]))
        mats.append(np.zeros((3, 3)))
    A = np.ascontiguousarray(np.stack(mats), dtype=np.float32)  # stored column-major per matrix: A[k, j, i]
    U, S, V, R = (np.empty_like(A) for _ in range(4))
    for k in range(A.shape[0]):
        U[k], S[k], V[k] = co.svd(A[k], KIND)
        R[k], _ = co.polar(A[k], KIND)
    np.savez_compressed(OUT / f"svd_{dim}d.npz", A=A, U=U, S=S, V=V, R=R)


def gen_cube() -> None:
    out = {}
    for name, (dim, res, lo, hi) in {
        "cfg1": (2, 50, 0.4, 0.6), "cfg2": (3, 64, 0.375, 0.625), "cfg5": (2, 25, 0.1, 0.3),
        "flip": (2, 7, -0.9, 0.3), "one": (3, 1, 0.5, 0.7),
    }.items():
        pts = co.cube(dim, res, lo, hi, KIND)
        out[f"{name}_args"] = np.array([dim, res, lo, hi], np.float64)
        out[f"{name}_axis"] = pts[:: res ** (dim - 1), 0].copy()  # the LinSpaced values themselves
        out[f"{name}_head"] = pts[:5].copy()
        out[f"{name}_tail"] = pts[-5:].copy()
        out[f"{name}_sum"] = np.array([pts.astype(np.float64).sum(), (pts.view(np.uint32).astype(np.uint64)).sum()])
    np.savez_compressed(OUT / "cube.npz", **out)


if __name__ == "__main__":
    co.build(ref=True)
    for dim in (2, 3):
        gen_svd(dim)
        for model in (co.SNOW, co.JELLY, co.LIQUID):
            gen_scene(dim, model)
            gen_random(dim, model)
    gen_cube()
    total = sum(p.stat().st_size for p in OUT.glob("*.npz"))
    print(f"wrote {len(list(OUT.glob('*.npz')))} files, {total / 1e6:.2f} MB; Q3 oob events: {co.lib(KIND).oob_events()}")

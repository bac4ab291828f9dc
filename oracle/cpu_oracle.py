"""TEST INFRASTRUCTURE ONLY — ctypes loader for the CPU oracles.

Three interchangeable back ends behind one class (same C entry points, different prefix):

* ``port``       oracle/_build/libnclr_oracle.so    plain-C restatement (oracle/nclr_oracle.c)
* ``ref_strict`` oracle/_ref/libnclr_ref_strict.so  the UNMODIFIED reference header
                 (/root/reference/src/nclr.h) built against oracle/eigen_standin, strict FP
* ``ref_fast``   oracle/_ref/libnclr_ref_fast.so    same, with the reference's own flags
                 (-Ofast -DNDEBUG, CMakeLists.txt:8-10): the timed CPU baseline

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (nuclearmpm_b200) never does.

Array conventions (numpy, float32, C-contiguous): x,v (n,dim); F,C (n,dim,dim) stored per particle
COLUMN-major like Eigen, i.e. ``F[p, j, i]`` is the mathematical F(i,j); Jp,mass,volume (n,).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIBS = {
    "port": (HERE / "_build" / "libnclr_oracle.so", "nclr_oracle_"),
    "ref_strict": (HERE / "_ref" / "libnclr_ref_strict.so", "nclr_ref_"),
    "ref_fast": (HERE / "_ref" / "libnclr_ref_fast.so", "nclr_ref_"),
}
SNOW, JELLY, LIQUID = 0, 1, 2
MODEL_NAMES = {SNOW: "snow", JELLY: "jelly", LIQUID: "liquid"}

_fp = C.POINTER(C.c_float)
_cache: dict[str, "_Lib"] = {}


def build(ref: bool | None = None) -> None:
    """Compile the oracle port (always) and oracle/_ref (only where /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", str(HERE), "port"], check=True)
    if ref is None:
        ref = Path(os.environ.get("NMPM_REFERENCE", "/root/reference")).joinpath("src/nclr.h").exists()
    if ref:
        subprocess.run(["make", "-s", "-C", str(HERE), "ref"], check=True)


def available(kind: str) -> bool:
    return _LIBS[kind][0].exists()


class _Lib:
    def __init__(self, kind: str):
        path, pre = _LIBS[kind]
        if not path.exists():
            if kind == "port":
                build(ref=False)
            else:
                raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.dll = C.CDLL(str(path))
        self.pre = pre
        d = self.dll

        def fn(name, res, args):
            f = getattr(d, pre + name)
            f.restype = res
            f.argtypes = args
            return f

        self.create = fn("create", C.c_void_p, [C.c_int] * 3 + [C.c_float] * 4 + [C.c_long] + [_fp] * 7)
        self.destroy = fn("destroy", None, [C.c_void_p])
        self.advance = fn("advance", C.c_int, [C.c_void_p, C.c_int])
        self.phase = fn("phase", C.c_int, [C.c_void_p, C.c_int])
        self.time_advance = fn("time_advance", C.c_double, [C.c_void_p, C.c_int])
        self.num_particles = fn("num_particles", C.c_long, [C.c_void_p])
        self.get_particles = fn("get_particles", None, [C.c_void_p] + [_fp] * 5)
        self.get_grid = fn("get_grid", C.c_long, [C.c_void_p, _fp, _fp])
        self.lame = fn("lame", None, [C.c_void_p, _fp, _fp])
        self.svd = fn("svd", None, [C.c_int] + [_fp] * 4)
        self.polar = fn("polar", None, [C.c_int] + [_fp] * 3)
        self.affine = fn("affine", None, [C.c_void_p, C.c_long, _fp])
        self.cube = fn("cube", C.c_long, [C.c_int, C.c_int, C.c_float, C.c_float, _fp])
        self.oob_events = fn("oob_events", C.c_long, [])
        self.oob_reset = fn("oob_reset", None, [])
        if kind == "port":
            self.set_grid = fn("set_grid", C.c_long, [C.c_void_p, _fp, _fp])
            self.cell_keys = fn("cell_keys", C.c_long,
                                [C.c_int, C.c_int, C.c_long, _fp, C.c_int, C.c_int,
                                 C.POINTER(C.c_int32), C.POINTER(C.c_uint32)])
            self.stable_sort = fn("stable_sort", None, [C.c_long, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)])


def lib(kind: str = "port") -> _Lib:
    if kind not in _cache:
        _cache[kind] = _Lib(kind)
    return _cache[kind]


def _p(a):
    return None if a is None else a.ctypes.data_as(_fp)


def _f32(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


class CpuSim:
    """CPU MPMSimulation<dim> (src/nclr.h:63-384) — ctor args and defaults as src/nclr.h:74-75."""

    def __init__(self, x, model: int, res: int = 64, dt: float = 1e-4, E: float = 1e4, nu: float = 0.2,
                 gravity: float = -100.0, v=None, F=None, Cm=None, Jp=None, mass=None, volume=None,
                 kind: str = "port"):
        self.L = lib(kind)
        x = _f32(x)
        self.n, self.dim = x.shape
        self.res, self.model = res, model
        n, d = self.n, self.dim
        args = [x, _f32(v, (n, d)), _f32(F, (n, d, d)), _f32(Cm, (n, d, d)), _f32(Jp, (n,)), _f32(mass, (n,)),
                _f32(volume, (n,))]
        self.h = C.c_void_p(self.L.create(d, model, res, dt, E, nu, gravity, n, *[_p(a) for a in args]))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.destroy(self.h)
            self.h = None

    def advance(self, nsteps: int = 1) -> None:
        if self.L.advance(self.h, nsteps):
            raise IndexError("particle stencil left the grid (reference: std::out_of_range, src/nclr.h:163)")

    def phase(self, which: int) -> None:
        if self.L.phase(self.h, which):
            raise IndexError("particle stencil left the grid")

    def time_advance(self, nsteps: int) -> float:
        return float(self.L.time_advance(self.h, nsteps))

    def particles(self) -> dict:
        n, d = self.n, self.dim
        out = dict(x=np.empty((n, d), np.float32), v=np.empty((n, d), np.float32),
                   F=np.empty((n, d, d), np.float32), C=np.empty((n, d, d), np.float32),
                   Jp=np.empty((n,), np.float32))
        self.L.get_particles(self.h, *[_p(out[k]) for k in ("x", "v", "F", "C", "Jp")])
        return out

    def grid(self):
        """(velocity (cells,dim), mass (cells,)) — empty before the first p2g (src/solver.cpp:52-57)."""
        n1 = self.res + 1
        cells = n1 ** self.dim
        gv = np.empty((cells, self.dim), np.float32)
        gm = np.empty((cells,), np.float32)
        got = self.L.get_grid(self.h, _p(gv), _p(gm))
        if got == 0:
            return gv[:0], gm[:0]
        return gv, gm

    def set_grid(self, gv, gm) -> None:
        """port only: overwrite the grid between phases (slab protocol tests)."""
        gv, gm = _f32(gv), _f32(gm)
        assert self.L.set_grid(self.h, _p(gv), _p(gm)) == gm.size

    def lame(self):
        a, b = C.c_float(), C.c_float()
        self.L.lame(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def affine(self, p: int) -> np.ndarray:
        A = np.empty((self.dim, self.dim), np.float32)
        self.L.affine(self.h, p, _p(A))
        return A


def svd(a: np.ndarray, kind: str = "port"):
    """nclr_svd (src/nclr_math.h:50-74) on one column-major dim×dim matrix → (U, sig, V)."""
    a = _f32(a)
    d = a.shape[0]
    U, S, V = (np.empty((d, d), np.float32) for _ in range(3))
    lib(kind).svd(d, _p(a), _p(U), _p(S), _p(V))
    return U, S, V


def polar(m: np.ndarray, kind: str = "port"):
    m = _f32(m)
    d = m.shape[0]
    R, S = (np.empty((d, d), np.float32) for _ in range(2))
    lib(kind).polar(d, _p(m), _p(R), _p(S))
    return R, S


def cube(dim: int, res: int, lo: float, hi: float, kind: str = "port") -> np.ndarray:
    """cube<dim>(res, min, max) (src/nclr_math.h:100-129): res**dim points, x slowest."""
    n = res ** dim
    out = np.empty((n, dim), np.float32)
    got = lib(kind).cube(dim, res, lo, hi, _p(out))
    assert got == n
    return out


def cell_keys(x: np.ndarray, res: int, mode: int = 0, tb: int = 2):
    """(base (n,dim) int32, keys (n,) uint32, n_out_of_grid) — the binning oracle (port only)."""
    x = _f32(x)
    n, d = x.shape
    base = np.empty((n, d), np.int32)
    keys = np.empty((n,), np.uint32)
    bad = lib("port").cell_keys(d, res, n, _p(x), mode, tb, base.ctypes.data_as(C.POINTER(C.c_int32)),
                                keys.ctypes.data_as(C.POINTER(C.c_uint32)))
    return base, keys, int(bad)


def stable_sort(keys: np.ndarray) -> np.ndarray:
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    perm = np.empty(keys.shape, np.uint32)
    lib("port").stable_sort(keys.size, keys.ctypes.data_as(C.POINTER(C.c_uint32)),
                            perm.ctypes.data_as(C.POINTER(C.c_uint32)))
    return perm

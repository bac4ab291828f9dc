"""Python mirror of the reference's solver interface over the C-ABI (include/nmpm.h).

Reference surface mirrored (same names, argument meaning and error behaviour):
  nclr::MaterialModel                          src/nclr.h:57-61
  nclr::MPMSimulation<dim>(particles, model, res=64, dt=1e-4, E=1e4, nu=0.2, gravity=-100)
                                               src/nclr.h:74-78
  advance(), particles(), grid(), mu_0, lambda_0, kBoundary, k*Hardening
                                               src/nclr.h:66-72,80-87
  nclr::cube<dim>(res, min, max)               src/nclr_math.h:100-129
Out-of-grid particles raise OutOfGridError (an IndexError), the analogue of the std::out_of_range the
reference throws from vector::at (src/nclr.h:163).

numpy conventions: x,v (n,dim); F,C (n,dim,dim) stored per particle column-major like Eigen, i.e.
``F[p, j, i]`` is F(i,j); Jp,mass,volume (n,).  Grid: (cells,dim) velocity and (cells,) mass in the
reference's node order (x slowest).
"""
from __future__ import annotations

import ctypes as ct
import enum
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
_fp = ct.POINTER(ct.c_float)
_u32p = ct.POINTER(ct.c_uint32)
_i32p = ct.POINTER(ct.c_int32)
NMPM_T_COUNT = 8


class NmpmError(RuntimeError):
    pass


class OutOfGridError(IndexError):
    """A particle's stencil left the grid (reference: std::out_of_range, src/nclr.h:163; Q5)."""


class MaterialModel(enum.IntEnum):
    kSnow = 0
    kJelly = 1
    kLiquid = 2


class Options(ct.Structure):
    _fields_ = [("device", ct.c_int), ("sort_every", ct.c_int), ("p2g_variant", ct.c_int), ("use_graph", ct.c_int),
                ("slab_x0", ct.c_int), ("slab_x1", ct.c_int), ("capacity", ct.c_int), ("g2p_window", ct.c_int),
                ("fuse", ct.c_int), ("tiles", ct.c_int), ("reserved", ct.c_int * 6)]


def lib_path() -> Path:
    import os
    override = os.environ.get("NMPM_LIB")  # experiments only: an alternative build of the SAME C-ABI
    return Path(override) if override else _PKG / "lib" / "libnmpm.so"


_lib = None


def load_library():
    """Load libnmpm.so; fails loudly when it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not p.exists():
        raise NmpmError(f"{p} is missing: build it with `python -m nuclearmpm_b200.build` (nvcc, sm_100a). "
                        "There is no CPU fallback.")
    L = ct.CDLL(str(p))
    vp, sz, ci, cf = ct.c_void_p, ct.c_size_t, ct.c_int, ct.c_float

    def sig(name, res, args):
        f = getattr(L, name)
        f.restype, f.argtypes = res, args

    sig("nmpm_default_options", None, [ct.POINTER(Options)])
    sig("nmpm_create", ci, [ci, ci, ci, cf, cf, cf, cf, sz] + [_fp] * 7 + [ct.POINTER(Options), ct.POINTER(vp)])
    sig("nmpm_create_aos", ci, [ci, ci, ci, cf, cf, cf, cf, sz, vp, sz, ct.POINTER(Options), ct.POINTER(vp)])
    szp = ct.POINTER(ct.c_size_t)
    sig("nmpm_create_batch", ci, [ci, ci, cf, cf, ci, szp, _fp, _fp] + [_fp] * 7 + [ct.POINTER(Options), ct.POINTER(vp)])
    sig("nmpm_create_batch_aos", ci, [ci, ci, cf, cf, ci, szp, _fp, _fp, vp, sz, ct.POINTER(Options), ct.POINTER(vp)])
    sig("nmpm_num_scenes", ci, [vp])
    sig("nmpm_batch_lame", ci, [vp, _fp, _fp])
    sig("nmpm_destroy", None, [vp])
    sig("nmpm_advance", ci, [vp, ci])
    sig("nmpm_phase", ci, [vp, ci])
    sig("nmpm_synchronize", ci, [vp])
    sig("nmpm_download_particles", ci, [vp] + [_fp] * 5)
    sig("nmpm_download_particles_aos", ci, [vp, vp, sz])
    sig("nmpm_download_positions", ci, [vp, _fp])
    sig("nmpm_download_grid", ci, [vp, _fp, _fp, ct.POINTER(sz)])
    sig("nmpm_download_grid_aos", ci, [vp, vp, sz, ct.POINTER(sz)])
    sig("nmpm_upload_particles", ci, [vp] + [_fp] * 5)
    sig("nmpm_upload_particles_async", ci, [vp] + [_fp] * 5)
    sig("nmpm_download_particles_async", ci, [vp] + [_fp] * 5)
    sig("nmpm_num_particles", sz, [vp])
    sig("nmpm_grid_cells", sz, [vp])
    sig("nmpm_lame", ci, [vp, _fp, _fp])
    sig("nmpm_sort_debug", ci, [vp, _i32p, _u32p, _u32p, _u32p, _u32p])
    sig("nmpm_key_tile_bits", ci, [vp])
    sig("nmpm_svd_batch", ci, [ci, sz, _fp, _fp, _fp, _fp, ci])
    sig("nmpm_polar_batch", ci, [ci, sz, _fp, _fp, ci])
    sig("nmpm_snow_project_batch", ci, [ci, sz, _fp, cf, cf, _fp, ci])
    sig("nmpm_affine_debug", ci, [vp, _fp])
    sig("nmpm_timing_enable", ci, [vp, ci])
    sig("nmpm_timing_read", ci, [vp, _fp, ct.POINTER(ci), ci])
    sig("nmpm_launch_count", ct.c_longlong, [vp])
    sig("nmpm_fused", ci, [vp])
    sig("nmpm_tiles_active", ci, [vp])
    sig("nmpm_set_stream", ci, [vp, vp])
    sig("nmpm_get_stream", vp, [vp])
    sig("nmpm_grid_plane_ptr", vp, [vp, ci])
    sig("nmpm_grid_plane_bytes", sz, [vp])
    sig("nmpm_grid_add_planes", ci, [vp, ci, ci, vp])
    sig("nmpm_slab_p2g", ci, [vp])
    sig("nmpm_slab_grid_g2p", ci, [vp, vp, vp, sz, vp])
    sig("nmpm_slab_unpack", ci, [vp, vp, sz, vp, sz, sz])
    sig("nmpm_slab_set_range", ci, [vp, ci, ci])
    sig("nmpm_slab_histogram", ci, [vp, vp])
    sig("nmpm_nccl_unique_id", ci, [vp, ct.c_char_p])
    sig("nmpm_slab_comm_init", ci, [vp, vp, ci, ci, _i32p, sz, ct.c_char_p])
    sig("nmpm_slab_step", ci, [vp, ci])
    sig("nmpm_slab_set_global_count", ci, [vp, sz])
    sig("nmpm_slab_set_bounds", ci, [vp, _i32p])
    sig("nmpm_slab_migrated", ct.c_longlong, [vp])
    sig("nmpm_slab_counts", ci, [vp, ct.POINTER(ct.c_longlong), ct.POINTER(ct.c_longlong)])
    sig("nmpm_set_ids", ci, [vp, _u32p])
    sig("nmpm_download_particles_slots", ci, [vp] + [_fp] * 5 + [_u32p])
    sig("nmpm_num_slots", sz, [vp])
    sig("nmpm_migrate_record_bytes", sz, [vp])
    sig("nmpm_last_error", ct.c_char_p, [vp])
    sig("nmpm_build_info", ct.c_char_p, [])
    _lib = L
    return L


def _f32(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(shape) if shape is not None else a


def _p(a):
    return None if a is None else a.ctypes.data_as(_fp)


def cube(dim: int, res: int, lo: float, hi: float) -> np.ndarray:
    """nclr::cube<dim>(res, min, max) (src/nclr_math.h:100-129): res**dim points, x slowest, built
    from Eigen's LinSpaced rule in fp32 (step=(hi-lo)/(res-1); last value pinned to hi, or — when
    |hi| < |lo| — first value pinned to lo and the rest counted back from hi)."""
    lo32, hi32 = np.float32(lo), np.float32(hi)
    i = np.arange(res, dtype=np.float32)
    if res == 1:
        axis = np.array([lo32], np.float32)  # Eigen: size1 = 1, so i == 0 takes the `low + i*step` branch
    else:
        step = np.float32((hi32 - lo32) / np.float32(res - 1))
        if abs(hi32) < abs(lo32):
            axis = (hi32 - (np.float32(res - 1) - i) * step).astype(np.float32)
            axis[0] = lo32
        else:
            axis = (lo32 + i * step).astype(np.float32)
            axis[-1] = hi32
    g = np.meshgrid(*([axis] * dim), indexing="ij")
    return np.stack([a.ravel() for a in g], axis=1).astype(np.float32)


class MPMSimulation:
    """GPU MPMSimulation<dim>.  `particles` is an (n,dim) float32 position array (dim = 2 or 3)."""

    kBoundary = 3            # src/nclr.h:66
    kSnowHardening = 10.0    # src/nclr.h:67
    kJellyHardening = 0.3    # src/nclr.h:68
    kLiquidHardening = 1.0   # src/nclr.h:69

    def __init__(self, particles, model, res: int = 64, dt: float = 1e-4, E: float = 1e4, nu: float = 0.2,
                 gravity: float = -100.0, *, v=None, F=None, C=None, Jp=None, mass=None, volume=None,
                 device: int = 0, sort_every: int = 4, p2g_variant: int = 0, slab=None, capacity: int = 0,
                 ids=None, g2p_window: int = 0, fuse: int = 0, tiles: int = 0):
        self._L = load_library()
        x = _f32(particles)
        if x.ndim != 2 or x.shape[1] not in (2, 3):
            raise ValueError("particles must be an (n, 2) or (n, 3) array of positions")
        self.n, self.dim = x.shape
        self.res, self.model = int(res), MaterialModel(int(model))
        n, d = self.n, self.dim
        opt = Options()
        self._L.nmpm_default_options(C_byref(opt))
        opt.device, opt.sort_every, opt.p2g_variant, opt.g2p_window = device, sort_every, p2g_variant, g2p_window
        opt.fuse = int(fuse)   # 0 auto, 1 off, 2 on (not on the first step after an upload), 3 always
        opt.tiles = int(tiles)  # active node tiles: 0 adaptive, 1 never, 2 always
        if slab is not None:
            opt.slab_x0, opt.slab_x1 = slab
            opt.capacity = int(capacity)
        arrs = [x, _f32(v, (n, d)), _f32(F, (n, d, d)), _f32(C, (n, d, d)), _f32(Jp, (n,)), _f32(mass, (n,)),
                _f32(volume, (n,))]
        h = C_void_p()
        rc = self._L.nmpm_create(d, int(model), int(res), dt, E, nu, gravity, n, *[_p(a) for a in arrs],
                                 C_byref(opt), C_byref(h))
        if rc != 0:
            raise NmpmError(f"nmpm_create failed ({rc}): {self._L.nmpm_last_error(None).decode()}")
        self._h = h
        mu, lam = ct.c_float(), ct.c_float()
        self._L.nmpm_lame(self._h, C_byref(mu), C_byref(lam))
        self.mu_0, self.lambda_0 = mu.value, lam.value
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint32)
            assert ids.shape == (n,)
            self._check(self._L.nmpm_set_ids(self._h, ids.ctypes.data_as(_u32p)), "nmpm_set_ids")

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.nmpm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc == 0:
            return
        msg = self._L.nmpm_last_error(self._h).decode()
        if rc == 3:
            raise OutOfGridError(msg)
        raise NmpmError(f"{what} failed ({rc}): {msg}")

    # -- the reference's interface ----------------------------------------------------------------
    def advance(self, nsteps: int = 1, sync: bool = False) -> None:
        """advance() × nsteps (src/nclr.h:80-84).  Asynchronous; errors surface at the next sync."""
        self._check(self._L.nmpm_advance(self._h, int(nsteps)), "nmpm_advance")
        if sync:
            self.synchronize()

    def synchronize(self) -> None:
        self._check(self._L.nmpm_synchronize(self._h), "nmpm_synchronize")

    def particles(self) -> dict:
        """particles() (src/nclr.h:86): state in INPUT order."""
        n, d = self.n, self.dim
        out = dict(x=np.empty((n, d), np.float32), v=np.empty((n, d), np.float32),
                   F=np.empty((n, d, d), np.float32), C=np.empty((n, d, d), np.float32),
                   Jp=np.empty((n,), np.float32))
        self._check(self._L.nmpm_download_particles(self._h, *[_p(out[k]) for k in ("x", "v", "F", "C", "Jp")]),
                    "nmpm_download_particles")
        return out

    def num_particles(self) -> int:
        return int(self._L.nmpm_num_particles(self._h))

    def particles_slots(self, out: dict | None = None, compact: bool = True) -> dict:
        """Slab sims: the live particles in device slot order plus their global ids.

        `out` (optional): preallocated arrays x,v,F,C,Jp,ids with room for >= nmpm_num_slots entries (e.g. views
        of pinned host memory, which makes the copy run at PCIe speed); the result then holds views of them.
        compact=False keeps the slots of particles that have just migrated away (ids == 0xFFFFFFFF)."""
        n, d = int(self._L.nmpm_num_slots(self._h)), self.dim
        if out is None:
            out = dict(x=np.empty((n, d), np.float32), v=np.empty((n, d), np.float32),
                       F=np.empty((n, d, d), np.float32), C=np.empty((n, d, d), np.float32),
                       Jp=np.empty((n,), np.float32), ids=np.empty((n,), np.uint32))
        else:
            if any(len(out[k]) < n for k in ("x", "v", "F", "C", "Jp", "ids")):
                raise ValueError(f"particles_slots: preallocated arrays hold fewer than {n} slots")
            out = {k: out[k][:n] for k in ("x", "v", "F", "C", "Jp", "ids")}
        self._check(self._L.nmpm_download_particles_slots(self._h, *[_p(out[k]) for k in ("x", "v", "F", "C", "Jp")],
                                                          out["ids"].ctypes.data_as(_u32p)),
                    "nmpm_download_particles_slots")
        if not compact:
            return out
        live = out["ids"] != 0xFFFFFFFF  # slots of particles that have just migrated away
        return out if live.all() else {k: a[live] for k, a in out.items()}

    def positions(self) -> np.ndarray:
        x = np.empty((self.n, self.dim), np.float32)
        self._check(self._L.nmpm_download_positions(self._h, _p(x)), "nmpm_download_positions")
        return x

    def grid(self):
        """grid() (src/nclr.h:87): (velocity (cells,dim), mass (cells,)); empty before the first step."""
        cells = (self.res + 1) ** self.dim * getattr(self, "nscenes", 1)
        gv = np.empty((cells, self.dim), np.float32)
        gm = np.empty((cells,), np.float32)
        got = ct.c_size_t(0)
        self._check(self._L.nmpm_download_grid(self._h, _p(gv), _p(gm), C_byref(got)), "nmpm_download_grid")
        if got.value == 0:
            return gv[:0], gm[:0]
        return gv, gm

    # -- test / profiling hooks ---------------------------------------------------------------------
    def phase(self, which: int) -> None:
        self._check(self._L.nmpm_phase(self._h, int(which)), "nmpm_phase")

    def upload(self, x, v=None, F=None, C=None, Jp=None) -> None:
        n, d = self.n, self.dim
        arrs = [_f32(x, (n, d)), _f32(v, (n, d)), _f32(F, (n, d, d)), _f32(C, (n, d, d)), _f32(Jp, (n,))]
        self._check(self._L.nmpm_upload_particles(self._h, *[_p(a) for a in arrs]), "nmpm_upload_particles")

    def upload_async(self, x, v=None, F=None, C=None, Jp=None) -> None:
        """Enqueue-only upload (nmpm_upload_particles_async).  The arrays must be float32, C-contiguous, of the right
        shape (no conversion copies are made: they would be freed before the copy runs) and stay alive until
        synchronize(); pinned memory gives real overlap."""
        for a, shape in ((x, (self.n, self.dim)), (v, (self.n, self.dim)), (F, (self.n, self.dim, self.dim)),
                         (C, (self.n, self.dim, self.dim)), (Jp, (self.n,))):
            if a is not None and (a.dtype != np.float32 or not a.flags.c_contiguous or a.shape != shape):
                raise ValueError("upload_async needs C-contiguous float32 arrays of the particle shapes")
        self._check(self._L.nmpm_upload_particles_async(self._h, *[_p(a) for a in (x, v, F, C, Jp)]),
                    "nmpm_upload_particles_async")

    def download_async(self, out: dict) -> None:
        """Enqueue-only download into preallocated arrays out[x|v|F|C|Jp] (valid after synchronize())."""
        self._check(self._L.nmpm_download_particles_async(self._h, *[_p(out.get(k)) for k in ("x", "v", "F", "C", "Jp")]),
                    "nmpm_download_particles_async")

    def sort_debug(self) -> dict:
        n, d = self.n, self.dim
        out = dict(base=np.empty((n, d), np.int32), keys=np.empty(n, np.uint32), keys_sorted=np.empty(n, np.uint32),
                   perm=np.empty(n, np.uint32), ids=np.empty(n, np.uint32))
        self._check(self._L.nmpm_sort_debug(self._h, out["base"].ctypes.data_as(_i32p),
                                            out["keys"].ctypes.data_as(_u32p),
                                            out["keys_sorted"].ctypes.data_as(_u32p),
                                            out["perm"].ctypes.data_as(_u32p), out["ids"].ctypes.data_as(_u32p)),
                    "nmpm_sort_debug")
        out["tile_bits"] = int(self._L.nmpm_key_tile_bits(self._h))
        return out

    def affine(self) -> np.ndarray:
        A = np.empty((self.n, self.dim, self.dim), np.float32)
        self._check(self._L.nmpm_affine_debug(self._h, _p(A)), "nmpm_affine_debug")
        return A

    def timing_enable(self, on: bool = True) -> None:
        self._L.nmpm_timing_enable(self._h, int(on))

    def timing_read(self, reset: bool = True):
        ms = (ct.c_float * NMPM_T_COUNT)()
        steps = ct.c_int(0)
        self._L.nmpm_timing_read(self._h, ms, C_byref(steps), int(reset))
        return dict(sort=ms[0], p2g=ms[1], grid=ms[2], g2p=ms[3], steps=steps.value)

    def launch_count(self) -> int:
        return int(self._L.nmpm_launch_count(self._h))

    @property
    def fused(self) -> int:
        """1/2 when G2P also scatters the next step's P2G (include/nmpm.h: nmpm_options.fuse), else 0"""
        return int(self._L.nmpm_fused(self._h))

    @property
    def tiles_active(self) -> bool:
        """True while grid_op / the grid clear work on active node tiles (nmpm_options.tiles) instead of the node box"""
        return bool(self._L.nmpm_tiles_active(self._h))

    def set_stream(self, cuda_stream: int) -> None:
        self._check(self._L.nmpm_set_stream(self._h, ct.c_void_p(cuda_stream)), "nmpm_set_stream")

    @property
    def stream(self) -> int:
        return int(self._L.nmpm_get_stream(self._h) or 0)


class MPMBatch(MPMSimulation):
    """A batch of independent 2D scenes behind one handle (include/nmpm.h: nmpm_create_batch; BASELINE config 5).

    `scenes`: list of (n_s, 2) position arrays; `E`, `nu`: one value per scene.  particles() / upload() / grid() work on
    the scenes' arrays concatenated; scene_slices gives each scene's range, grids(s) its (res+1)^2 cells."""

    def __init__(self, scenes, model, res: int = 64, dt: float = 1e-4, E=None, nu=None, gravity: float = -100.0, *,
                 v=None, device: int = 0, sort_every: int = 4):
        self._L = load_library()
        xs = [_f32(x) for x in scenes]
        if not xs or any(x.ndim != 2 or x.shape[1] != 2 for x in xs):
            raise ValueError("scenes must be a non-empty list of (n, 2) position arrays")
        self.nscenes = len(xs)
        counts = np.array([len(x) for x in xs], dtype=np.uintp)
        x = np.ascontiguousarray(np.concatenate(xs))
        self.n, self.dim = x.shape
        self.res, self.model = int(res), MaterialModel(int(model))
        self.scene_slices = [slice(int(a), int(a + c)) for a, c in zip(np.cumsum(counts) - counts, counts)]
        Es = _f32(np.full(self.nscenes, 1e4) if E is None else E, (self.nscenes,))
        nus = _f32(np.full(self.nscenes, 0.2) if nu is None else nu, (self.nscenes,))
        n, d = self.n, 2
        opt = Options()
        self._L.nmpm_default_options(C_byref(opt))
        opt.device, opt.sort_every = device, sort_every
        vv = None if v is None else np.ascontiguousarray(np.concatenate([_f32(a) for a in v]))
        arrs = [x, _f32(vv, (n, d)), _f32(None, (n, d, d)), _f32(None, (n, d, d)), _f32(None, (n,)), _f32(None, (n,)),
                _f32(None, (n,))]
        h = C_void_p()
        rc = self._L.nmpm_create_batch(int(model), int(res), dt, gravity, self.nscenes,
                                       counts.ctypes.data_as(ct.POINTER(ct.c_size_t)), _p(Es), _p(nus),
                                       *[_p(a) for a in arrs], C_byref(opt), C_byref(h))
        if rc != 0:
            raise NmpmError(f"nmpm_create_batch failed ({rc}): {self._L.nmpm_last_error(None).decode()}")
        self._h = h
        mu, lam = np.empty(self.nscenes, np.float32), np.empty(self.nscenes, np.float32)
        self._L.nmpm_batch_lame(self._h, _p(mu), _p(lam))
        self.mu_0, self.lambda_0 = mu, lam

    def scene_particles(self, s: int, state: dict | None = None) -> dict:
        state = state or self.particles()
        return {k: a[self.scene_slices[s]] for k, a in state.items()}

    def grids(self, s: int, grid=None):
        gv, gm = grid or self.grid()
        cells = (self.res + 1) ** 2
        return gv[s * cells:(s + 1) * cells], gm[s * cells:(s + 1) * cells]


def C_byref(x):
    return ct.byref(x)


def C_void_p():
    return ct.c_void_p()


def svd_batch(A: np.ndarray, device: int = 0):
    """Device nclr_svd (src/nclr_math.h:50-74) on a batch of column-major matrices (k,dim,dim)."""
    L = load_library()
    A = _f32(A)
    k, d, _ = A.shape
    U, S, V = (np.empty_like(A) for _ in range(3))
    rc = L.nmpm_svd_batch(d, k, _p(A), _p(U), _p(S), _p(V), device)
    if rc:
        raise NmpmError(f"nmpm_svd_batch failed ({rc}): {L.nmpm_last_error(None).decode()}")
    return U, S, V


def polar_batch(A: np.ndarray, device: int = 0) -> np.ndarray:
    L = load_library()
    A = _f32(A)
    k, d, _ = A.shape
    R = np.empty_like(A)
    rc = L.nmpm_polar_batch(d, k, _p(A), _p(R), device)
    if rc:
        raise NmpmError(f"nmpm_polar_batch failed ({rc}): {L.nmpm_last_error(None).decode()}")
    return R


def snow_project_batch(A: np.ndarray, lo: float = 0.975, hi: float = 1.0045, device: int = 0) -> np.ndarray:
    """Device U clamp(sig, lo, hi) V^T of nclr_svd(A) (src/nclr.h:239-247), column-major (k,dim,dim)."""
    L = load_library()
    A = _f32(A)
    k, d, _ = A.shape
    G = np.empty_like(A)
    rc = L.nmpm_snow_project_batch(d, k, _p(A), lo, hi, _p(G), device)
    if rc:
        raise NmpmError(f"nmpm_snow_project_batch failed ({rc}): {L.nmpm_last_error(None).decode()}")
    return G

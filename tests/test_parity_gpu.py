"""GPU parity tests (run on the B200 with `-m gpu`): the CUDA path, called through the C-ABI
(include/nmpm.h via nuclearmpm_b200.sim), against the golden vectors generated from the reference
header itself (tests/golden) and against the oracle port live on the same seeded inputs.

Bar (BASELINE.json north_star): integer binning bit-exact; floating-point state
within the ONE-STEP tolerances below (teacher-forced from reference states, robust to chaos):
"""
import numpy as np
import pytest

from conftest import dense_grid
import nuclearmpm_b200 as nm
from oracle import cpu_oracle as co

pytestmark = pytest.mark.gpu

MODELS = [co.SNOW, co.JELLY, co.LIQUID]
FIELDS = ("x", "v", "F", "C", "Jp")

# one-step tolerances (absolute; v, C and grid velocity are relative to max(1, |.|max))
# Measured on B200 (profiles/r01_parity.md): x <= 1 ulp, v <= 4e-6, C <= 3.5e-5, F <= 1e-6, Jp <= 2e-6,
# grid v <= 1.6e-5 — the floor set by P2G summation order + FMA contraction (C is a difference of
# O(4/dx * |v|) terms, so it carries the largest relative error).  The reference's own -Ofast vs strict
# self-noise after ONE step is of the same order.
TOL_X = 2.4e-7      # 2 ulp at x < 1
TOL_V = 1e-5
TOL_C = 5e-5
TOL_F = 2e-5
TOL_JP = 1e-4
TOL_GRID_V = 3e-5
TOL_GRID_M = 1e-5   # relative to max node mass


def check_state(got, ref, what, scale=1.0):
    vmax = max(1.0, float(np.abs(ref["v"]).max()))
    cmax = max(1.0, float(np.abs(ref["C"]).max()))
    errs = dict(x=np.abs(got["x"] - ref["x"]).max(), v=np.abs(got["v"] - ref["v"]).max() / vmax,
                C=np.abs(got["C"] - ref["C"]).max() / cmax, F=np.abs(got["F"] - ref["F"]).max(),
                Jp=np.abs(got["Jp"] - ref["Jp"]).max())
    tol = dict(x=TOL_X, v=TOL_V, C=TOL_C, F=TOL_F, Jp=TOL_JP)
    bad = {k: (float(errs[k]), tol[k] * scale) for k in errs if not errs[k] <= tol[k] * scale}
    assert not bad, f"{what}: {bad} (all errors: { {k: float(v) for k, v in errs.items()} })"
    return errs


def check_grid(gv, gm, ref_gv, ref_gm, what, scale=1.0):
    vmax = max(1.0, float(np.abs(ref_gv).max()))
    mmax = max(1e-30, float(np.abs(ref_gm).max()))
    ev = np.abs(gv - ref_gv).max() / vmax
    em = np.abs(gm - ref_gm).max() / mmax
    assert ev <= TOL_GRID_V * scale and em <= TOL_GRID_M * scale, f"{what}: grid v err {ev}, mass err {em}"


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [2, 3])
def test_device_svd_polar_against_golden(golden, dim):
    """nclr_svd / nclr_polar on the device (src/nclr_math.h:50-98).  U and V are compared through the
    well-conditioned products (U sig V^T, R = U V^T), never column by column."""
    g = golden(f"svd_{dim}d")
    A = g["A"]
    U, S, V = nm.svd_batch(A)
    R = nm.polar_batch(A)
    eye = np.eye(dim)
    for k in range(A.shape[0]):
        a = A[k].T.astype(np.float64)
        u, s, v = U[k].T.astype(np.float64), S[k].T.astype(np.float64), V[k].T.astype(np.float64)
        scale = max(1.0, np.abs(a).max())
        assert np.abs(u @ s @ v.T - a).max() <= 3e-6 * scale
        assert np.abs(u @ u.T - eye).max() <= 3e-6 and np.abs(v @ v.T - eye).max() <= 3e-6
        sv, sv_ref = np.diag(s), np.diag(g["S"][k].T.astype(np.float64))
        assert np.abs(sv - sv_ref).max() <= 3e-6 * scale  # singular values (incl. the signed last one) agree
        if dim == 3:
            assert np.linalg.det(u) > 0 and np.linalg.det(v) > 0
        # polar factor: compared where it is well conditioned — 3D: rank >= 2 (the det=+1 completion makes
        # it unique); 2D closed form: |(m00+m11, m10-m01)| not tiny
        rank = int((sv_ref > 1e-6 * scale).sum())
        if dim == 3:
            well = rank >= 2
        else:
            well = np.hypot(a[0, 0] + a[1, 1], a[1, 0] - a[0, 1]) > 1e-2 * scale
        if well:
            assert np.abs(R[k] - g["R"][k]).max() <= 2e-5, (k, R[k], g["R"][k])


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("model", MODELS)
def test_binning_bit_exact(golden, dim, model):
    """K0: base coordinates, cell keys and the radix-sorted order are bit-exact against the oracle
    (std::stable_sort order by (key, slot))."""
    g = golden(f"scene_{dim}d_{co.MODEL_NAMES[model]}")
    res = int(g["res"])
    rng = np.random.default_rng(3)
    x = g["s100_x"].copy()
    rng.shuffle(x, axis=0)
    sim = nm.MPMSimulation(x, model, res)
    d = sim.sort_debug()
    base, keys, bad = co.cell_keys(x, res, mode=1, tb=d["tile_bits"])
    assert bad == 0
    assert (d["ids"] == np.arange(len(x))).all()
    assert (d["base"] == base).all()
    assert (d["keys"] == keys).all()
    perm = co.stable_sort(keys)
    assert (d["perm"] == perm).all()
    assert (d["keys_sorted"] == keys[perm]).all()


def test_radix_sort_large_random_keys():
    """Sort correctness at a size spanning many tiles and all 4 digit passes."""
    rng = np.random.default_rng(11)
    n = 300_000
    x = rng.uniform(0.02, 0.98, (n, 3)).astype(np.float32)
    sim = nm.MPMSimulation(x, co.JELLY, 512)
    d = sim.sort_debug()
    base, keys, bad = co.cell_keys(x, 512, mode=1, tb=d["tile_bits"])
    assert bad == 0 and (d["keys"] == keys).all()
    perm = np.argsort(keys, kind="stable").astype(np.uint32)
    assert (d["perm"] == perm).all() and (d["keys_sorted"] == keys[perm]).all()


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("variant", [1, 2, 3, 4, 42, 44])
def test_one_step_from_random_state(golden, dim, model, variant):
    """Whole step from a random (v,F,C,Jp,mass,volume) state, phase by phase: stress, post-P2G grid
    (+ conservation), post-grid_op grid, particle state."""
    g = golden(f"random_{dim}d_{co.MODEL_NAMES[model]}")
    res = int(g["res"])
    sim = nm.MPMSimulation(g["x0"], model, res, float(g["dt"]), float(g["E"]), float(g["nu"]), float(g["gravity"]),
                           v=g["v0"], F=g["F0"], C=g["C0"], Jp=g["Jp0"], mass=g["mass"], volume=g["volume"],
                           p2g_variant=variant)
    A = sim.affine()
    amax = np.abs(g["affine0"]).max()
    assert np.abs(A - g["affine0"]).max() <= 2e-5 * max(1.0, amax)
    cells = (res + 1) ** dim
    sim.phase(0)
    gv, gm = sim.grid()
    ref_gv, ref_gm = dense_grid(g, "p2g", cells, dim)
    check_grid(gv, gm, ref_gv, ref_gm, "post-P2G")
    # conservation (post-P2G, pre-grid_op: Q6)
    assert np.isclose(gm.astype(np.float64).sum(), g["mass"].astype(np.float64).sum(), rtol=1e-6)
    mom = (g["mass"][:, None].astype(np.float64) * g["v0"]).sum(0)
    assert np.allclose(gv.astype(np.float64).sum(0), mom, rtol=1e-4, atol=1e-2)
    sim.phase(1)
    gv, gm = sim.grid()
    ref_gv, ref_gm = dense_grid(g, "gop", cells, dim)
    check_grid(gv, gm, ref_gv, ref_gm, "post-grid_op")
    sim.phase(2)
    got = sim.particles()
    check_state(got, {k: g[f"s1_{k}"] for k in FIELDS}, "one step from random state")


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("model", MODELS)
def test_teacher_forced_steps(golden, dim, model):
    """Upload reference states (after 0,1,2,3(→ next), 100 steps), advance ONE step, compare."""
    g = golden(f"scene_{dim}d_{co.MODEL_NAMES[model]}")
    res = int(g["res"])
    sim = nm.MPMSimulation(g["x0"], model, res)
    # step 1 from the initial state, with both grid stages
    cells = (res + 1) ** dim
    assert sim.grid()[1].size == 0  # empty before the first advance (src/solver.cpp:52-57)
    sim.phase(0)
    check_grid(*sim.grid(), *dense_grid(g, "p2g1", cells, dim), "post-P2G step 1")
    sim.phase(1)
    check_grid(*sim.grid(), *dense_grid(g, "gop1", cells, dim), "post-grid_op step 1")
    sim.phase(2)
    check_state(sim.particles(), {k: g[f"s1_{k}"] for k in FIELDS}, "step 1")
    for a, b in [(1, 2), (2, 3), (100, 101)]:
        sim.upload(*[g[f"s{a}_{k}"] for k in FIELDS])
        sim.advance(1)
        check_state(sim.particles(), {k: g[f"s{b}_{k}"] for k in FIELDS}, f"teacher-forced {a}->{b}")
        if b == 101:
            check_grid(*sim.grid(), *dense_grid(g, "gop101", cells, dim), "grid step 101")


@pytest.mark.parametrize("dim,model,nsteps,tolx", [
    (2, co.JELLY, 100, 1e-5), (2, co.LIQUID, 100, 1e-5), (2, co.SNOW, 100, 2e-5),
    (3, co.JELLY, 100, 1e-5), (3, co.LIQUID, 100, 1e-5), (3, co.SNOW, 3, 1e-5),
])
def test_free_running(golden, dim, model, nsteps, tolx):
    """Free-running horizons over which the reference agrees with itself across FP modes:
    jelly/liquid and 2D snow 100 steps; 3D snow only 3 steps (it decorrelates at step 4, Q1)."""
    g = golden(f"scene_{dim}d_{co.MODEL_NAMES[model]}")
    sim = nm.MPMSimulation(g["x0"], model, int(g["res"]))
    sim.advance(nsteps)
    got = sim.particles()
    assert np.abs(got["x"] - g[f"s{nsteps}_x"]).max() <= tolx
    vmax = max(1.0, np.abs(g[f"s{nsteps}_v"]).max())
    assert np.abs(got["v"] - g[f"s{nsteps}_v"]).max() <= 2e-3 * vmax


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("model", MODELS)
def test_against_oracle_port_live(dim, model):
    """Same seeded inputs through the oracle port and the CUDA path, step by step (teacher-forced from
    the oracle's states), on a scene that is not in the fixtures."""
    rng = np.random.default_rng(2024 + dim * 3 + model)
    n = 3000 if dim == 2 else 4000
    x = rng.uniform(0.3, 0.7, (n, dim)).astype(np.float32)
    v = rng.normal(0, 1, (n, dim)).astype(np.float32)
    res = 64 if dim == 2 else 32
    cpu = co.CpuSim(x, model, res, v=v)
    gpu = nm.MPMSimulation(x, model, res, v=v)
    for step in range(12):
        cpu.advance(1)
        gpu.advance(1)
        ref = cpu.particles()
        check_state(gpu.particles(), ref, f"step {step + 1}")
        check_grid(*gpu.grid(), *cpu.grid(), f"grid step {step + 1}")
        gpu.upload(*[ref[k] for k in FIELDS])


@pytest.mark.parametrize("dim", [2, 3])
def test_variants_and_sort_cadence_agree(dim):
    """P2G variants (per-particle vs cell-segmented reductions) and sort cadences give the same physics
    up to summation order; particles() stays in input order whatever the device order."""
    rng = np.random.default_rng(5)
    n = 5000
    x = rng.uniform(0.35, 0.65, (n, dim)).astype(np.float32)
    ref = None
    for variant, sort_every in [(1, 0), (1, 1), (2, 1), (2, 3), (3, 1), (3, 3), (4, 1), (4, 4), (43, 3)]:
        sim = nm.MPMSimulation(x, co.JELLY, 64, p2g_variant=variant, sort_every=sort_every)
        sim.advance(20)
        st = sim.particles()
        if ref is None:
            ref = st
        else:
            check_state(st, ref, f"variant {variant} sort_every {sort_every}", scale=20.0)


@pytest.mark.parametrize("model", MODELS)
def test_p2g_group_kernel_ragged_counts(model):
    """Variants 3 and 4 (three 9-lane groups per warp walking thirds of the warp's slots / three streams over
    32*C slots) on particle counts that leave groups or streams empty or short, or fill exactly one warp:
    post-P2G grid against the per-particle scatter (variant 1) on the same cell-sorted state, mass conservation."""
    rng = np.random.default_rng(17)
    for n in (1, 2, 11, 12, 13, 21, 22, 23, 24, 25, 31, 32, 33, 45, 63, 64, 65, 127, 129, 300, 515):
        x = rng.uniform(0.40, 0.48, (n, 3)).astype(np.float32)   # a few cells: long and short segments
        v = rng.normal(size=(n, 3)).astype(np.float32)
        grids = {}
        for variant in (1, 3, 4, 42, 44):
            sim = nm.MPMSimulation(x, model, 32, v=v, p2g_variant=variant, sort_every=1)
            sim.phase(0)
            grids[variant] = sim.grid()
        gv1, gm1 = grids[1]
        assert np.isclose(gm1.astype(np.float64).sum(), n, rtol=1e-6)
        for variant in (3, 4, 42, 44):
            gv, gm = grids[variant]
            check_grid(gv, gm, gv1, gm1, f"n={n} variant {variant}")
            assert np.isclose(gm.astype(np.float64).sum(), n, rtol=1e-6)


def test_out_of_grid_is_reported():
    """Q5: std::out_of_range in the reference → OutOfGridError (an IndexError) here."""
    x = np.array([[0.5, 0.5], [0.999, 0.5]], np.float32)
    sim = nm.MPMSimulation(x, co.JELLY, 64)
    with pytest.raises(nm.OutOfGridError):
        sim.advance(1)
        sim.synchronize()
    # NaN positions are caught the same way instead of corrupting memory
    x = np.array([[0.5, 0.5], [np.nan, 0.5]], np.float32)
    sim = nm.MPMSimulation(x, co.SNOW, 64)
    with pytest.raises(nm.OutOfGridError):
        sim.advance(2, sync=True)


def test_empty_and_single_particle():
    sim = nm.MPMSimulation(np.zeros((0, 3), np.float32), co.SNOW, 32)
    sim.advance(3, sync=True)
    assert sim.particles()["x"].shape == (0, 3)
    gv, gm = sim.grid()
    assert gm.shape == (33 ** 3,) and not gm.any()
    x = np.array([[0.5, 0.5, 0.5]], np.float32)
    cpu = co.CpuSim(x, co.SNOW, 32)
    gpu = nm.MPMSimulation(x, co.SNOW, 32)
    cpu.advance(2)
    gpu.advance(2)
    check_state(gpu.particles(), cpu.particles(), "single particle", scale=2.0)


def test_aos_roundtrip_reference_layout():
    """nmpm_create_aos / nmpm_download_particles_aos use the reference's Particle<dim> record
    (src/nclr.h:20-48): x, v, F, C, Jp, mass, volume, c — 64 B in 2D, 112 B in 3D."""
    import ctypes as C
    L = nm.load_library()
    for dim, words in [(2, 16), (3, 28)]:
        rng = np.random.default_rng(dim)
        n = 777
        rec = np.zeros((n, words), np.float32)
        D = dim
        rec[:, :D] = rng.uniform(0.3, 0.7, (n, D))
        rec[:, D:2 * D] = rng.normal(0, 1, (n, D))
        eye = np.eye(D, dtype=np.float32)
        eye[2:, 2:] = 0
        rec[:, 2 * D:2 * D + D * D] = eye.T.reshape(-1)
        rec[:, 2 * D + 2 * D * D] = 1.0       # Jp
        rec[:, 2 * D + 2 * D * D + 1] = 1.5   # mass
        rec[:, 2 * D + 2 * D * D + 2] = 0.75  # volume
        rec.view(np.int32)[:, 2 * D + 2 * D * D + 3] = np.arange(n)  # colour
        h = C.c_void_p()
        rc = L.nmpm_create_aos(dim, co.JELLY, 32, 1e-4, 1e4, 0.2, -100.0, n, rec.ctypes.data, words * 4, None,
                               C.byref(h))
        assert rc == 0, L.nmpm_last_error(None)
        assert L.nmpm_advance(h, 5) == 0
        out = rec.copy()
        assert L.nmpm_download_particles_aos(h, out.ctypes.data, words * 4) == 0
        L.nmpm_destroy(h)
        cpu = co.CpuSim(rec[:, :D], co.JELLY, 32, v=rec[:, D:2 * D], mass=rec[:, 2 * D + 2 * D * D + 1],
                        volume=rec[:, 2 * D + 2 * D * D + 2])
        cpu.advance(5)
        ref = cpu.particles()
        got = dict(x=out[:, :D], v=out[:, D:2 * D], F=out[:, 2 * D:2 * D + D * D].reshape(n, D, D),
                   C=out[:, 2 * D + D * D:2 * D + 2 * D * D].reshape(n, D, D), Jp=out[:, 2 * D + 2 * D * D])
        check_state(got, ref, f"AoS {dim}D", scale=5.0)
        assert (out[:, 2 * D + 2 * D * D + 1] == 1.5).all() and (out[:, 2 * D + 2 * D * D + 2] == 0.75).all()
        assert (out.view(np.int32)[:, 2 * D + 2 * D * D + 3] == np.arange(n)).all()


@pytest.mark.parametrize("model", MODELS)
def test_full_size_properties_config2(model):
    """BASELINE config 2 size (64^3 particles, 128^3 grid): size-independent properties — mass and
    momentum conservation on the post-P2G grid, |grid v| <= vmax, displacement <= 0.9 dx per step,
    3D jelly/liquid keep F(:,2) == 0 (Q1), snow has Jp = 0.6 after step 1 (Q1)."""
    x = nm.cube(3, 64, 0.375, 0.625)
    sim = nm.MPMSimulation(x, model, 128)
    sim.advance(3)
    before = sim.particles()
    sim.phase(0)
    gv, gm = sim.grid()
    assert np.isclose(gm.astype(np.float64).sum(), float(len(x)), rtol=1e-6)
    mom = before["v"].astype(np.float64).sum(0)
    assert np.allclose(gv.astype(np.float64).sum(0), mom, rtol=1e-4, atol=1e-3 * len(x) * 1e-3)
    sim.phase(1)
    gv, gm = sim.grid()
    vmax = np.float32((1.0 / 128) * 0.9 / 1e-4)
    assert np.abs(gv).max() <= vmax
    sim.phase(2)
    after = sim.particles()
    assert np.abs(after["x"] - before["x"]).max() <= 0.9 / 128 * 1.0001
    if model != co.SNOW:
        assert not after["F"][:, 2, :].any()  # third COLUMN (stored [p, j, i]) stays exactly zero
    else:
        one = nm.MPMSimulation(x, model, 128)
        one.advance(1)
        assert (one.particles()["Jp"] == np.float32(0.6)).all()
    assert np.isfinite(after["x"]).all()


def test_full_size_properties_config4():
    """BASELINE config 4 size (256^3 = 16.8 M snow particles, 512^3 grid; the bench workload and the only size at
    which the three-stream P2G (variant 4) is the default): mass and momentum conservation on the post-P2G grid of an
    in-place (not re-binned) step, |grid v| <= vmax, displacement <= 0.9 dx, Jp = 0.6 after step 1 (Q1), and the
    default path against the plain column-lane kernel (variant 3) after 3 steps."""
    x = nm.cube(3, 256, 0.25, 0.5)
    n = len(x)
    sim = nm.MPMSimulation(x, co.SNOW, 512)          # auto: variant 4, sort cadence 4
    sim.advance(1)
    assert (sim.particles()["Jp"] == np.float32(0.6)).all()
    sim.advance(1)                                    # step 2 done; step 3 (index 2) is an in-place step
    before = sim.particles()
    sim.phase(0)
    gv, gm = sim.grid()
    assert np.isclose(gm.astype(np.float64).sum(), float(n), rtol=1e-6)
    mom = before["v"].astype(np.float64).sum(0)
    assert np.allclose(gv.astype(np.float64).sum(0), mom, rtol=1e-4, atol=1e-6 * n)
    del gv, gm
    sim.phase(1)
    gv, gm = sim.grid()
    vmax = np.float32((1.0 / 512) * 0.9 / 1e-4)
    assert np.abs(gv).max() <= vmax
    del gv, gm
    sim.phase(2)
    after = sim.particles()
    assert np.isfinite(after["x"]).all()
    assert np.abs(after["x"] - before["x"]).max() <= 0.9 / 512 * 1.0001
    del before, sim
    # three steps (one re-binned, two in place) of the default path against the plain column-lane kernel: same physics
    # up to the summation order of the reductions.  Three steps only: 3D snow decorrelates at step 4 (Q1, SURVEY.md
    # 4.3 - measured here: after 6 steps 0.1 % of the particles are more than a cell apart between the two variants)
    ref_sim = nm.MPMSimulation(x, co.SNOW, 512, p2g_variant=3)
    ref_sim.advance(3)
    ref = ref_sim.particles()
    assert np.abs(after["x"] - ref["x"]).max() <= 3e-5
    # velocities: at res 512 the stress term carries dt*vol*4/dx^2*2mu ~ 5e7, so the 1e-7 rounding of the polar factor
    # is a 1e-3 relative effect on v within one step (64x the golden scenes at res 64); measured between the two
    # variants after 3 steps: median 3.2e-4, 99.9 % 1.5e-3, max 2.9e-3 of |v|max (= the clamp speed, Q1)
    dv = np.abs(after["v"] - ref["v"]).max(axis=1)
    vmax = max(1.0, float(np.abs(ref["v"]).max()))
    stats = (float(np.median(dv)) / vmax, float(np.quantile(dv, 0.999)) / vmax, float(dv.max()) / vmax)
    assert stats[0] <= 1e-3 and stats[1] <= 5e-3 and stats[2] <= 1e-2, stats


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["snow_F", "snow_Fprime_q1", "snow_Fprime_phys", "jelly_rank2", "liquid_diag",
                                  "general", "step0", "scaled"])
def test_device_fast_polar_and_snow_projection(name):
    """The register fast path (one-sided Jacobi recompose, csrc/nmpm_math.cuh) against the oracle's
    nclr_polar / nclr_svd + clamp in every regime the solver meets (tests/mathcases.py)."""
    from mathcases import TOL, matrices, oracle_reference
    A = matrices()[name]
    R = nm.polar_batch(A)
    G = nm.snow_project_batch(A, 0.975, 1.0045)
    Ro, Go, well = oracle_reference(co, A)
    assert np.isfinite(R).all() and np.isfinite(G).all()
    if well.any():
        assert np.abs(R - Ro)[well].max() <= TOL[name]
        assert np.abs(G - Go)[well].max() <= TOL[name]


# ------------------------------------------------------------------------------------------
def test_positions_only_download_matches_particles():
    """N4: the positions-only hand-off a frame loop uses (src/example.cpp:56-82 reads just p.x every 10th step):
    nmpm_download_positions == the x of nmpm_download_particles, in input order, on re-binned and in-place steps."""
    rng = np.random.default_rng(41)
    for dim, model in [(2, co.SNOW), (3, co.JELLY)]:
        x = rng.uniform(0.3, 0.7, (3000, dim)).astype(np.float32)
        rng.shuffle(x, axis=0)
        gpu = nm.MPMSimulation(x, model, 64 if dim == 2 else 32)
        cpu = co.CpuSim(x, model, 64 if dim == 2 else 32)
        assert (gpu.positions() == x).all()          # before the first step: the input, in input order
        for step in range(6):                        # sort cadence 4: steps 0 and 4 re-bin, the others run in place
            gpu.advance(1)
            cpu.advance(1)
            pos = gpu.positions()
            assert pos.shape == (3000, dim) and (pos == gpu.particles()["x"]).all()
            assert np.abs(pos - cpu.particles()["x"]).max() <= TOL_X * (step + 1)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("after_phase", [0, 1])
def test_upload_in_the_middle_of_a_step_discards_the_partial_grid(dim, after_phase):
    """nmpm_upload_particles after nmpm_phase(P2G) / (GRID_OP): the aborted step's node sums must not leak into the
    following steps (the next step clears only the box of the previous COMPLETE P2G)."""
    rng = np.random.default_rng(7 + dim)
    n = 2000
    res = 64 if dim == 2 else 32
    xa = rng.uniform(0.55, 0.75, (n, dim)).astype(np.float32)   # the aborted step's particles: a different region
    xb = rng.uniform(0.25, 0.45, (n, dim)).astype(np.float32)
    v = rng.normal(0, 1, (n, dim)).astype(np.float32)
    gpu = nm.MPMSimulation(xa, co.JELLY, res, v=v)
    gpu.advance(1)
    for ph in range(after_phase + 1):
        gpu.phase(ph)
    gpu.upload(xb, v)
    cpu = co.CpuSim(xb, co.JELLY, res, v=v)
    for step in range(3):
        gpu.advance(1)
        cpu.advance(1)
        check_grid(*gpu.grid(), *cpu.grid(), f"grid after upload, step {step + 1}")
        check_state(gpu.particles(), cpu.particles(), f"state after upload, step {step + 1}", scale=step + 1.0)


@pytest.mark.parametrize("model", MODELS)
def test_g2p_tma_window_matches_global_gather(model):
    """G2P with the node window staged by the TMA (cp.async.bulk.tensor + mbarrier; nmpm_options.g2p_window = 2: one-shot
    CTAs, 3: the persistent software-pipelined kernel) reads exactly the nodes the plain gather reads: bit-identical particle state, step after step, on a compact
    block (every CTA in the window), on scattered particles (bounding boxes larger than the window: fall-back) and
    next to the upper grid boundary (boxes clipped by the tensor map: zero fill is never read)."""
    rng = np.random.default_rng(77)
    scenes = {
        "block": nm.cube(3, 40, 0.3, 0.6),
        "scattered": rng.uniform(0.1, 0.9, (20000, 3)).astype(np.float32),
        "corner": nm.cube(3, 24, 0.80, 0.95),
    }
    for name, x in scenes.items():
        v = rng.normal(0, 2, x.shape).astype(np.float32)
        a = nm.MPMSimulation(x, model, 64, v=v, g2p_window=1, fuse=1)   # the gather under test, not the fused scatter
        b = nm.MPMSimulation(x, model, 64, v=v, g2p_window=2)   # one-shot CTAs, TMA-staged window
        c = nm.MPMSimulation(x, model, 64, v=v, g2p_window=3)   # persistent CTAs, cp.async rows + TMA window, pipelined
        for step in range(6):
            s0 = a.particles()
            a.advance(1)
            # teacher-forced (3D snow is chaotic from step 4 on): the other two gathers step from a's state.  Not
            # bit-identical: every sim's P2G sums its node contributions with atomics in its own order (1 ulp)
            for other, what in ((b, "window"), (c, "pipelined window")):
                other.upload(*[s0[k] for k in FIELDS])
                other.advance(1)
                # (|v| ~ 2 per axis: snow's F sits at the edge of the one-step tolerance, and the atomics order varies
                # from run to run — twice the tolerance keeps the test about the gather, not about that noise)
                check_state(other.particles(), a.particles(), f"{name}: {what} vs global gather, step {step + 1}", scale=2.0)
        if name != "corner":
            cpu = co.CpuSim(x, model, 64, v=v)
            cpu.advance(1)
            one = nm.MPMSimulation(x, model, 64, v=v, g2p_window=2)
            one.advance(1)
            check_state(one.particles(), cpu.particles(), f"window {name}")


@pytest.mark.parametrize("dim", [2, 3])
def test_async_upload_download_pipeline(dim):
    """nmpm_upload_particles_async / nmpm_download_particles_async (copy-in of step k+1 and copy-out of step k-1 overlap
    step k): a teacher-forced loop over oracle states gives, after one synchronize, exactly what the blocking calls give."""
    rng = np.random.default_rng(90 + dim)
    n, res = 6000, (64 if dim == 2 else 32)
    x = rng.uniform(0.3, 0.7, (n, dim)).astype(np.float32)
    v = rng.normal(0, 1, (n, dim)).astype(np.float32)
    cpu = co.CpuSim(x, co.JELLY, res, v=v)
    states = []
    for _ in range(5):
        states.append({k: np.ascontiguousarray(a) for k, a in cpu.particles().items()})
        cpu.advance(1)
    states.append(cpu.particles())
    gpu = nm.MPMSimulation(x, co.JELLY, res, v=v)
    outs = [{k: np.empty_like(states[0][k]) for k in FIELDS} for _ in range(5)]
    for k in range(5):
        gpu.upload_async(*[states[k][f] for f in FIELDS])
        gpu.advance(1)
        gpu.download_async(outs[k])
    gpu.synchronize()
    for k in range(5):
        check_state(outs[k], states[k + 1], f"async pipeline step {k}")
    # and the blocking calls still agree after the async ones
    gpu.upload(*[states[2][f] for f in FIELDS])
    gpu.advance(1)
    check_state(gpu.particles(), states[3], "blocking after async")


@pytest.mark.parametrize("dim,sort_every", [(2, 4), (3, 4), (2, 1), (2, 3), (2, 0)])
def test_cycle_graph_equals_single_steps(dim, sort_every):
    """nmpm_advance(h, n) replays ONE CUDA graph per cycle of host states (8 steps at cadence 4) for runs of steps;
    the result must be what n single-step calls give (same kernels in the same order; only the atomics order differs)."""
    rng = np.random.default_rng(400 + dim)
    n, res = 5000, (64 if dim == 2 else 32)
    x = rng.uniform(0.3, 0.7, (n, dim)).astype(np.float32)
    v = rng.normal(0, 1, (n, dim)).astype(np.float32)
    a = nm.MPMSimulation(x, co.JELLY, res, v=v, sort_every=sort_every)
    b = nm.MPMSimulation(x, co.JELLY, res, v=v, sort_every=sort_every)
    a.advance(61)                      # singles up to the first aligned step, then whole cycles, then a tail of singles
    for _ in range(61):
        b.advance(1)
    check_state(a.particles(), b.particles(), f"cycle graphs, cadence {sort_every}", scale=61.0)
    cpu = co.CpuSim(x, co.JELLY, res, v=v)
    cpu.advance(61)
    check_state(a.particles(), cpu.particles(), "cycle graphs vs oracle", scale=61.0)

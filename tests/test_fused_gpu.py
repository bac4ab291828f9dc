"""Fused G2P+P2G (nuclearmpm_b200/csrc/nmpm_fused.cuh; nmpm_options.fuse): the G2P of step n scatters the P2G of step
n+1 into a second grid.  It must be invisible at the API: same states as the unfused kernels and the oracle, grid()
still the grid of the last advance(), uploads discard the sums scattered ahead, phases can still be called one by one."""
import numpy as np
import pytest

import nuclearmpm_b200 as nm
from oracle import cpu_oracle as co
from test_parity_gpu import FIELDS, MODELS, check_grid, check_state

pytestmark = pytest.mark.gpu

OFF, ON, ALWAYS = 1, 2, 3


def scene(seed, n=4000):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.3, 0.7, (n, 3)).astype(np.float32)
    v = rng.normal(0, 1, (n, 3)).astype(np.float32)
    return x, v


def test_fused_is_the_default_in_3d():
    x, v = scene(1, 100)
    assert nm.MPMSimulation(x, co.SNOW, 32, v=v).fused == 1
    assert nm.MPMSimulation(x, co.SNOW, 32, v=v, fuse=OFF).fused == 0
    assert nm.MPMSimulation(x, co.SNOW, 32, v=v, fuse=ALWAYS).fused == 2
    assert nm.MPMSimulation(x, co.SNOW, 32, v=v, sort_every=0).fused == 0          # needs the cell order
    assert nm.MPMSimulation(x[:, :2].copy(), co.SNOW, 64).fused == 0               # 3D kernel


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("sort_every", [1, 3, 4])
def test_teacher_forced_with_speculation_discarded_every_step(model, sort_every):
    """fuse=ALWAYS + an upload after every step: every fused G2P scatters ahead, every upload throws those sums away,
    the next step runs the stand-alone P2G on the uploaded state.  States and grids against the oracle."""
    x, v = scene(300 + model)
    cpu = co.CpuSim(x, model, 32, v=v)
    gpu = nm.MPMSimulation(x, model, 32, v=v, fuse=ALWAYS, sort_every=sort_every)
    for step in range(10):
        cpu.advance(1)
        gpu.advance(1)
        ref = cpu.particles()
        # random velocities of |v| ~ 1 per axis are harsher than the golden scenes: snow's F reaches ~1e-5 of the 2e-5
        # one-step tolerance and the atomics order varies from run to run
        slack = 1.5 if model == co.SNOW else 1.0
        check_state(gpu.particles(), ref, f"step {step + 1}", scale=slack)
        check_grid(*gpu.grid(), *cpu.grid(), f"grid step {step + 1}", scale=slack)
        gpu.upload(*[ref[k] for k in FIELDS])


@pytest.mark.parametrize("model", MODELS)
def test_scatter_ahead_against_the_oracle_one_step(model):
    """The P2G scattered ahead by the fused kernel (for snow: stress from the SVD factors of the plasticity projection
    instead of a polar decomposition of the stored F) against the oracle's p2g() of the same step, teacher-forced: upload
    the oracle's state k, advance one step (stand-alone P2G of step k+1, fused G2P scatters step k+2), then read the
    post-P2G grid of step k+2 on both sides."""
    x, v = scene(350 + model)
    cpu = co.CpuSim(x, model, 32, v=v)
    gpu = nm.MPMSimulation(x, model, 32, v=v, fuse=ALWAYS)
    for step in range(8):
        cpu.advance(1), gpu.advance(1)
        check_state(gpu.particles(), cpu.particles(), f"step {step + 1}", scale=1.5 if model == co.SNOW else 1.0)
        ref = cpu.particles()                        # the oracle's p2g() from the oracle's own state
        twin = co.CpuSim(ref["x"], model, 32, v=ref["v"], F=ref["F"], Cm=ref["C"], Jp=ref["Jp"])
        twin.phase(0)
        gpu.phase(0)
        gv, gm = gpu.grid()
        # the GPU scatters ITS state of step k+1 (one-step error: F to 2e-5, amplified by dt*vol*Dinv*2mu ~ 5 in the
        # stress): twice the one-step grid tolerance
        check_grid(gv, gm, *twin.grid(), f"P2G scattered ahead, step {step + 2}", scale=2.0)
        gpu.upload(*[ref[k] for k in FIELDS])        # discards nothing (the P2G phase consumed it), restarts the step


@pytest.mark.parametrize("model", [co.JELLY, co.LIQUID])
@pytest.mark.parametrize("sort_every", [1, 4])
def test_free_running_fused_against_oracle_and_unfused(model, sort_every):
    """Runs of fused steps of several lengths (single steps, partial and whole graph cycles): particle states against the
    oracle and the unfused kernels, and grid() = the grid of the LAST step (the scatter ahead went to the other buffer)."""
    x, v = scene(310 + model)
    cpu = co.CpuSim(x, model, 32, v=v)
    fused = nm.MPMSimulation(x, model, 32, v=v, fuse=ON, sort_every=sort_every)
    plain = nm.MPMSimulation(x, model, 32, v=v, fuse=OFF, sort_every=sort_every)
    done = 0
    for k in (1, 1, 2, 5, 8, 16, 3, 24):
        cpu.advance(k), fused.advance(k), plain.advance(k)
        done += k
        ref = cpu.particles()
        check_state(fused.particles(), ref, f"fused vs oracle after {done} steps", scale=float(done))
        check_state(fused.particles(), plain.particles(), f"fused vs unfused after {done} steps", scale=float(done))
        check_grid(*fused.grid(), *cpu.grid(), f"grid after {done} steps", scale=float(done))


def test_snow_three_fused_steps_against_oracle():
    """3D snow decorrelates at step 4 (Q1): three free-running steps, two of them fused."""
    x, v = scene(320)
    cpu = co.CpuSim(x, co.SNOW, 32, v=v)
    gpu = nm.MPMSimulation(x, co.SNOW, 32, v=v, fuse=ALWAYS)
    cpu.advance(3), gpu.advance(3)
    # one-step errors are amplified ~10x per step by the snow stress term (tests/test_parity_gpu.py::test_free_running
    # allows 2e-3 |v|max after 3 steps); the unfused kernels sit at the same distance from the oracle
    plain = nm.MPMSimulation(x, co.SNOW, 32, v=v, fuse=OFF)
    plain.advance(3)
    for sim, what in ((gpu, "fused"), (plain, "unfused")):
        check_state(sim.particles(), cpu.particles(), f"snow, 3 {what} steps", scale=20.0)
        check_grid(*sim.grid(), *cpu.grid(), f"snow grid after 3 {what} steps", scale=20.0)


@pytest.mark.parametrize("model", MODELS)
def test_phases_one_by_one_after_fused_steps(model):
    """nmpm_phase(P2G) after a fused step only swaps the buffers: the post-P2G grid must be what the stand-alone P2G
    gives on the same state (mass conserved), and the remaining phases must complete the step."""
    x, v = scene(330 + model)
    fused = nm.MPMSimulation(x, model, 32, v=v, fuse=ALWAYS)
    plain = nm.MPMSimulation(x, model, 32, v=v, fuse=OFF)
    fused.advance(2), plain.advance(2)
    state = plain.particles()
    fused.upload(*[state[k] for k in FIELDS])   # same state in both (snow: no drift to argue about)
    fused.advance(1), plain.advance(1)          # fused: stand-alone P2G (upload), fused G2P scatters step 4 ahead
    state = plain.particles()
    check_state(fused.particles(), state, "step 3")
    fused.phase(0), plain.phase(0)
    gv, gm = fused.grid()
    assert np.isclose(gm.astype(np.float64).sum(), len(x), rtol=1e-6)
    check_grid(gv, gm, *plain.grid(), "post-P2G grid of step 4", scale=4.0)
    fused.phase(1), plain.phase(1)
    check_grid(*fused.grid(), *plain.grid(), "post-grid_op grid of step 4", scale=4.0)
    fused.phase(2), plain.phase(2)
    check_state(fused.particles(), plain.particles(), "step 4", scale=4.0)
    # upload in the middle of the NEXT step (after its P2G phase consumed the sums scattered ahead)
    fused.phase(0)
    state = plain.particles()
    fused.upload(*[state[k] for k in FIELDS])
    fused.advance(2), plain.advance(2)
    if model != co.SNOW:
        check_state(fused.particles(), plain.particles(), "two steps after a mid-step upload", scale=4.0)
    gm = fused.grid()[1]
    assert np.isclose(gm.astype(np.float64).sum(), plain.grid()[1].astype(np.float64).sum(), rtol=1e-5)


@pytest.mark.parametrize("model", MODELS)
def test_fused_scatter_ragged_counts(model):
    """Particle counts that leave lane groups empty or short, fill exactly one warp, or spill into the next CTA:
    the P2G scattered ahead (read back through nmpm_phase(P2G) of the next step) against the per-particle scatter."""
    rng = np.random.default_rng(23)
    for n in (1, 2, 11, 12, 22, 23, 31, 32, 33, 45, 64, 65, 127, 128, 129, 300, 515):
        x = rng.uniform(0.40, 0.48, (n, 3)).astype(np.float32)
        v = rng.normal(size=(n, 3)).astype(np.float32)
        fused = nm.MPMSimulation(x, model, 32, v=v, fuse=ALWAYS, sort_every=2)
        plain = nm.MPMSimulation(x, model, 32, v=v, fuse=OFF, p2g_variant=1, sort_every=2)
        for step in range(3):
            fused.advance(1), plain.advance(1)
            state = plain.particles()
            check_state(fused.particles(), state, f"n={n} step {step + 1}", scale=2.0)
            fused.phase(0), plain.phase(0)
            gv, gm = fused.grid()
            assert np.isclose(gm.astype(np.float64).sum(), n, rtol=1e-6), f"n={n}: mass {gm.sum()}"
            check_grid(gv, gm, *plain.grid(), f"n={n} P2G of step {step + 2}", scale=2.0)
            fused.phase(1), fused.phase(2), plain.phase(1), plain.phase(2)
            state = plain.particles()
            fused.upload(*[state[k] for k in FIELDS])
            plain.upload(*[state[k] for k in FIELDS])


def test_out_of_grid_with_fused_steps():
    """Q5 under the fused kernel: a stencil outside the grid raises OutOfGridError (the sticky walls keep particles
    that start inside from ever leaving, so the only way in is a bad initial or uploaded position)."""
    x, v = scene(340, 500)
    x[7, 1] = 0.995
    gpu = nm.MPMSimulation(x, co.JELLY, 32, v=v, fuse=ALWAYS)
    with pytest.raises(nm.OutOfGridError):
        gpu.advance(2, sync=True)
    x[7, 1] = 0.5
    gpu = nm.MPMSimulation(x, co.JELLY, 32, v=v, fuse=ALWAYS)
    gpu.advance(3, sync=True)
    bad = gpu.particles()
    bad["x"][3, 0] = np.nan
    gpu.upload(*[bad[k] for k in FIELDS])
    with pytest.raises(nm.OutOfGridError):
        gpu.advance(1, sync=True)

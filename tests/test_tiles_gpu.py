"""Active node tiles (nmpm_options.tiles; nuclearmpm_b200/csrc/nmpm_kernels.cuh: tile_mark_warp, k_tiles3): grid_op and the
grid clear visit only the 4^3-node tiles that stencils cover.  Invisible in the results: the dense grid() must equal the
oracle's node for node (a node that was not cleared, or not updated, shows up at once), with the fused and the unfused
kernels, through uploads, and while the adaptive policy switches the mode on and off."""
import numpy as np
import pytest

import nuclearmpm_b200 as nm
from oracle import cpu_oracle as co
from test_parity_gpu import FIELDS, MODELS, check_grid, check_state

pytestmark = pytest.mark.gpu

AUTO, NEVER, ALWAYS, ADAPTIVE = 0, 1, 2, 3   # nmpm_options.tiles


def scene(seed, n=6000, lo=0.25, hi=0.75):
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, (n, 3)).astype(np.float32), rng.normal(0, 1.5, (n, 3)).astype(np.float32)


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("fuse", [1, 3])
def test_tiles_teacher_forced_against_the_oracle(model, fuse):
    x, v = scene(500 + model)
    cpu = co.CpuSim(x, model, 48, v=v)
    gpu = nm.MPMSimulation(x, model, 48, v=v, tiles=ALWAYS, fuse=fuse)
    for step in range(10):
        cpu.advance(1), gpu.advance(1)
        assert gpu.tiles_active          # decided on the device with every new set of positions
        ref = cpu.particles()
        check_state(gpu.particles(), ref, f"step {step + 1}", scale=2.0)   # |v| ~ 1.5 per axis at res 48: snow's F at the
        check_grid(*gpu.grid(), *cpu.grid(), f"grid step {step + 1}", scale=2.0)   # edge of the one-step tolerance
        if step % 3 != 2:                       # two teacher-forced steps, then one that carries on from the GPU state
            gpu.upload(*[ref[k] for k in FIELDS])
        else:
            cpu = co.CpuSim(gpu.particles()["x"], model, 48, **{("Cm" if k == "C" else k): a
                                                                 for k, a in gpu.particles().items() if k != "x"})


@pytest.mark.parametrize("model", [co.JELLY, co.LIQUID])
@pytest.mark.parametrize("fuse", [1, 2])
def test_tiles_free_running_and_grid_node_for_node(model, fuse):
    x, v = scene(510 + model)
    cpu = co.CpuSim(x, model, 48, v=v)
    tiled = nm.MPMSimulation(x, model, 48, v=v, tiles=ALWAYS, fuse=fuse)
    boxed = nm.MPMSimulation(x, model, 48, v=v, tiles=NEVER, fuse=fuse)
    done = 0
    for k in (1, 2, 5, 8, 16, 3, 24):
        cpu.advance(k), tiled.advance(k), boxed.advance(k)
        done += k
        check_state(tiled.particles(), cpu.particles(), f"tiles vs oracle after {done} steps", scale=float(done))
        check_state(tiled.particles(), boxed.particles(), f"tiles vs box after {done} steps", scale=float(done))
        gv, gm = tiled.grid()
        rv, rm = cpu.grid()
        check_grid(gv, gm, rv, rm, f"grid after {done} steps", scale=float(done))
        assert np.array_equal(gm != 0, rm != 0), "the set of non-zero nodes differs from the oracle's"


def test_adaptive_policy_switches_on_for_a_dispersed_scene_and_off_for_a_block():
    """Scattered particles: the node box holds ~6 nodes per particle -> tiles come on after the first read-back of the box;
    a compact 8-per-cell block (0.5 nodes per particle) stays in box mode.  Results agree with the box mode throughout."""
    rng = np.random.default_rng(77)
    x = rng.uniform(0.1, 0.9, (20000, 3)).astype(np.float32)
    v = rng.normal(0, 1, x.shape).astype(np.float32)
    auto = nm.MPMSimulation(x, co.JELLY, 64, v=v, tiles=ADAPTIVE)
    boxed = nm.MPMSimulation(x, co.JELLY, 64, v=v, tiles=NEVER)
    cpu = co.CpuSim(x, co.JELLY, 64, v=v)
    assert not auto.tiles_active
    for k in (1, 1, 1, 4, 9, 16, 8):
        auto.advance(k, sync=True), boxed.advance(k), cpu.advance(k)
    assert auto.tiles_active
    check_state(auto.particles(), boxed.particles(), "adaptive vs box", scale=40.0)
    check_state(auto.particles(), cpu.particles(), "adaptive vs oracle", scale=40.0)
    check_grid(*auto.grid(), *cpu.grid(), "adaptive grid", scale=40.0)
    block = nm.MPMSimulation(nm.cube(3, 40, 0.3, 0.6), co.SNOW, 64, tiles=ADAPTIVE)
    for k in (1, 1, 4, 8):
        block.advance(k, sync=True)
    assert not block.tiles_active
    small = nm.MPMSimulation(x, co.JELLY, 64, v=v)           # auto: a 65^3 grid never pays for the tile machinery
    small.advance(3, sync=True)
    assert not small.tiles_active


def test_mode_switch_in_both_directions_keeps_the_grid_clean():
    """NMPM-internal transitions (flags valid for some ring slots only): force them by toggling the policy through uploads
    of a compact and of a dispersed state into the SAME sim, and compare the dense grid with the oracle after each."""
    rng = np.random.default_rng(78)
    n = 8000
    compact = rng.uniform(0.45, 0.55, (n, 3)).astype(np.float32)       # 10^3 nodes: far below 1 node per particle
    spread = rng.uniform(0.1, 0.9, (n, 3)).astype(np.float32)         # 50^3 nodes
    sim = nm.MPMSimulation(compact, co.LIQUID, 64, tiles=ADAPTIVE)
    for phase, x in enumerate([spread, compact, spread, compact]):
        z = np.zeros((n, 3), np.float32)
        eye = np.tile(np.eye(3, dtype=np.float32), (n, 1, 1))
        sim.upload(x, z, eye, np.zeros((n, 3, 3), np.float32), np.ones(n, np.float32))
        cpu = co.CpuSim(x, co.LIQUID, 64)
        for k in (1, 1, 1, 1, 2, 3):
            sim.advance(k, sync=True), cpu.advance(k)
            check_grid(*sim.grid(), *cpu.grid(), f"phase {phase}", scale=10.0)
        assert sim.tiles_active == (phase % 2 == 0), f"phase {phase}: tiles_active = {sim.tiles_active}"
        check_state(sim.particles(), cpu.particles(), f"phase {phase}", scale=10.0)


def test_repaired_state_after_an_out_of_grid_error_starts_from_clean_grids():
    """An out-of-grid particle scatters to clamped nodes that no tile flag covers (its key is "out of grid").  The step
    fails; uploading a repaired state into the same handle must not inherit those sums."""
    x, v = scene(530, 3000)
    sim = nm.MPMSimulation(x, co.JELLY, 48, v=v, tiles=ALWAYS)
    sim.advance(2, sync=True)
    bad = sim.particles()
    bad["x"][11] = (0.999, 0.5, 0.5)
    sim.upload(*[bad[k] for k in FIELDS])
    with pytest.raises(nm.OutOfGridError):
        sim.advance(2, sync=True)
    cpu = co.CpuSim(x, co.JELLY, 48, v=v)
    init = cpu.particles()                        # (the reference's initial F is diag<3>(1) = diag(1, 1, 0), Q1)
    sim.upload(*[init[k] for k in FIELDS])
    for step in range(3):
        sim.advance(1, sync=True), cpu.advance(1)
        ref = cpu.particles()
        check_grid(*sim.grid(), *cpu.grid(), f"grid {step + 1} steps after the repaired upload", scale=2.0)
        check_state(sim.particles(), ref, f"state {step + 1} steps after the repaired upload", scale=2.0)
        assert np.array_equal(sim.grid()[1] != 0, cpu.grid()[1] != 0), "stale nodes survived the error"
        sim.upload(*[ref[k] for k in FIELDS])     # teacher-forced: node velocities of tiny masses amplify any drift

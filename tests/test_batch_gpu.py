"""Batches of independent 2D scenes behind one handle (include/nmpm.h: nmpm_create_batch — BASELINE.json config 5; the
reference runs one scene per process, src/solver.cpp:45-62): every scene must evolve exactly as it does alone — against
the oracle scene by scene, and against single-scene GPU runs."""
import numpy as np
import pytest

import nuclearmpm_b200 as nm
from oracle import cpu_oracle as co
from test_parity_gpu import FIELDS, MODELS, check_grid, check_state

pytestmark = pytest.mark.gpu


def two_cube_scene(s, n_side=25):
    """tools/make_scenes.py: two squares [a_k, a_k+0.2]^2, E ~ U(500,5000), nu ~ U(0.2,0.4)"""
    rng = np.random.default_rng(1234 + s)
    a = rng.uniform(0.1, 0.7, size=2)
    E, nu = rng.uniform(500, 5000), rng.uniform(0.2, 0.4)
    x = np.concatenate([nm.cube(2, n_side, float(a[k]), float(a[k]) + 0.2) for k in range(2)])
    return x, float(E), float(nu)


@pytest.mark.parametrize("model", MODELS)
def test_batch_against_the_oracle_scene_by_scene(model):
    scenes = [two_cube_scene(s) for s in range(7)]
    scenes[3] = (scenes[3][0][:700], scenes[3][1], scenes[3][2])       # ragged: scenes of different sizes
    batch = nm.MPMBatch([x for x, _, _ in scenes], model, 64, E=[e for _, e, _ in scenes], nu=[n for _, _, n in scenes])
    cpus = [co.CpuSim(x, model, 64, 1e-4, E, nu, -100.0) for x, E, nu in scenes]
    for s, c in enumerate(cpus):
        mu, lam = c.lame()
        assert batch.mu_0[s] == np.float32(mu) and batch.lambda_0[s] == np.float32(lam)
    for step in range(12):
        batch.advance(1)
        for c in cpus:
            c.advance(1)
        state, grid = batch.particles(), batch.grid()
        for s, c in enumerate(cpus):
            ref = c.particles()
            check_state(batch.scene_particles(s, state), ref, f"scene {s} step {step + 1}")
            check_grid(*batch.grids(s, grid), *c.grid(), f"scene {s} grid step {step + 1}")
        ref_all = {k: np.concatenate([c.particles()[k] for c in cpus]) for k in FIELDS}
        batch.upload(*[ref_all[k] for k in FIELDS])                    # teacher-forced


@pytest.mark.parametrize("model", [co.JELLY, co.LIQUID, co.SNOW])
def test_batch_free_running_equals_single_scene_runs(model):
    """200 free-running steps (whole graph cycles): the scenes of a batch against the same scenes run one by one."""
    scenes = [two_cube_scene(s) for s in range(10)]
    batch = nm.MPMBatch([x for x, _, _ in scenes], model, 64, E=[e for _, e, _ in scenes], nu=[n for _, _, n in scenes])
    batch.advance(200, sync=True)
    state = batch.particles()
    for s, (x, E, nu) in enumerate(scenes):
        one = nm.MPMSimulation(x, model, 64, E=E, nu=nu)
        one.advance(200, sync=True)
        check_state(batch.scene_particles(s, state), one.particles(), f"scene {s}", scale=200.0)


def test_batch_walls_are_per_scene():
    """A block falling onto the floor of its scene must stop there (sticky walls within 3 nodes of EVERY scene's border,
    src/nclr.h:298-306), not fall through into the scene stacked below it in the tall grid."""
    x = nm.cube(2, 20, 0.4, 0.6)
    x[:, 0] -= 0.3                                         # x is the stacking axis: push the block towards the scene border
    v = [np.tile(np.float32([-3.0, 0.0]), (len(x), 1))] * 3
    batch = nm.MPMBatch([x, x, x], co.JELLY, 64, v=v)
    one = nm.MPMSimulation(x, co.JELLY, 64, v=v[0])
    batch.advance(400, sync=True), one.advance(400, sync=True)
    ref = one.particles()
    assert ref["x"][:, 0].min() > 2.5 / 64                 # stopped by the wall
    state = batch.particles()
    for s in range(3):
        check_state(batch.scene_particles(s, state), ref, f"scene {s}", scale=400.0)


def test_batch_out_of_grid_in_one_scene_fails_the_batch():
    x, _, _ = two_cube_scene(0)
    bad = x.copy()
    bad[5, 0] = 0.999
    batch = nm.MPMBatch([x, bad, x], co.JELLY, 64)
    with pytest.raises(nm.OutOfGridError):
        batch.advance(1, sync=True)

"""Slab (multi-GPU) path on the device.

* one GPU, two logical slabs in one process: the nmpm_slab_* C-ABI calls (ghost-plane reduction, migration
  pack/unpack, sorted compaction) against the single-domain CUDA path and the oracle;
* >= 2 GPUs (skipped otherwise): slab.SlabSimulation over NCCL, one process per GPU.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))

import nuclearmpm_b200 as nm  # noqa: E402
from nuclearmpm_b200 import slab  # noqa: E402
from oracle import cpu_oracle as co  # noqa: E402

pytestmark = pytest.mark.gpu


def moving_block(dim, res):
    x = co.cube(2, 40, 0.3, 0.6) if dim == 2 else co.cube(3, 24, 0.3, 0.6)
    v = np.zeros_like(x)
    v[:, 0] = 100.0 if dim == 2 else 60.0
    return x, v


def check(got, ref, k, res, model=None):
    """k steps free-running.  At res 512 the rounding of the polar factor is amplified by the stress prefactor
    (tests/test_parity_fullres_gpu.py: amp = dt vol 4/dx^2 2 mu dx); below that the golden-scene tolerances hold."""
    vmax = max(1.0, float(np.abs(ref["v"]).max()))
    cmax = max(1.0, float(np.abs(ref["C"]).max()))
    noise = 0.0
    if res >= 256 and model is not None:
        e = {co.SNOW: float(np.exp(10.0 * (1.0 - float(ref["Jp"].min())))), co.JELLY: 0.3, co.LIQUID: 1.0}[model]
        noise = 8 * float(np.finfo(np.float32).eps) * 1e-4 * (4.0 * res * res) * 2.0 * (1e4 / 2.4) * e / res
    assert np.abs(got["x"] - ref["x"]).max() <= (2.4e-7 + 1e-4 * noise) * k
    assert np.abs(got["v"] - ref["v"]).max() <= (1e-5 * vmax + noise) * k
    assert np.abs(got["F"] - ref["F"]).max() <= (2e-5 + 1e-4 * 4 * res * noise) * k
    assert np.abs(got["C"] - ref["C"]).max() <= (5e-5 * cmax + 1e-6 * 4 * res * vmax + 4 * res * noise) * k
    assert np.abs(got["Jp"] - ref["Jp"]).max() <= (1e-4 + 1e-4 * 4 * res * noise) * k


@pytest.mark.parametrize("sort_every", [1, 3])
@pytest.mark.parametrize("dim,model", [(2, co.SNOW), (3, co.JELLY), (3, co.LIQUID), (3, co.SNOW)])
def test_two_logical_slabs_on_one_gpu(dim, model, sort_every):
    import torch
    res = 64 if dim == 2 else 32
    steps = 10 if dim == 2 else (3 if model == co.SNOW else 6)
    x, v = moving_block(dim, res)
    bx = slab.base_x(x, res)
    hist = np.bincount(bx, minlength=res + 1)
    b = slab.balanced_bounds(hist, 2)
    ranges = [(0, b[1]), (b[1], res + 1 + (1 << 20))]
    eng = []
    for (x0, x1) in ranges:
        mine = np.nonzero((bx >= x0) & (bx < x1))[0]
        eng.append(slab.GpuSlabEngine(x[mine], mine.astype(np.uint32), model, res, 1e-4, 1e4, 0.2, -100.0, (x0, x1),
                                      len(x) + 1024, 0, v=v[mine], sort_every=sort_every))
    W, cap = eng[0].rec_words, len(x)
    send = [[e.new_buffer(cap * W), e.new_buffer(cap * W)] for e in eng]
    counts = [e.new_buffer(4, "int32") for e in eng]
    migrated = 0
    for _ in range(steps):
        for e in eng:
            e.p2g()
        from_right = eng[1].plane_view(b[1], 2).clone()
        from_left = eng[0].plane_view(b[1], 2).clone()
        eng[0].add_planes(b[1], 2, from_right)
        eng[1].add_planes(b[1], 2, from_left)
        # the shared planes now hold identical sums on both sides
        torch.cuda.synchronize()
        assert torch.equal(eng[0].plane_view(b[1], 2), eng[1].plane_view(b[1], 2))
        for i, e in enumerate(eng):
            e.grid_g2p(send[i][0], send[i][1], cap, counts[i])
        c = [t.cpu().numpy() for t in counts]
        assert c[0][3] == 0 and c[1][3] == 0 and c[0][0] == 0 and c[1][1] == 0
        eng[0].unpack(send[1][0], 0, send[1][0], int(c[1][0]), int(c[0][0] + c[0][1]))  # left slab receives from the right
        eng[1].unpack(send[0][1], int(c[0][1]), send[0][1], 0, int(c[1][0] + c[1][1]))  # right slab receives from the left
        migrated += int(c[0][1] + c[1][0])
        assert eng[0].num_particles() + eng[1].num_particles() == len(x)
    assert migrated > 0
    got = {k: np.empty_like(a) for k, a in dict(x=x, v=v, F=np.empty((len(x), dim, dim), np.float32),
                                                C=np.empty((len(x), dim, dim), np.float32),
                                                Jp=np.empty(len(x), np.float32)).items()}
    seen = np.zeros(len(x), int)
    for e in eng:
        part = e.download_slots()
        ids = part["ids"].astype(np.int64)
        seen[ids] += 1
        for k in got:
            got[k][ids] = part[k]
    assert (seen == 1).all()
    ref = co.CpuSim(x, model, res, v=v)
    ref.advance(steps)
    check(got, ref.particles(), steps, res, model)
    # and the single-domain CUDA path agrees too
    one = nm.MPMSimulation(x, model, res, v=v)
    one.advance(steps)
    check(got, one.particles(), steps, res)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def thin_block_512():
    """cfg4-shaped: res 512, 8 particles per cell, a block only 24 cells thick along x — with 8 ranks the slabs hit
    the minimum width, some ranks start (almost) empty, and every particle is within a few cells of a slab cut."""
    x = nm.cube(3, 48, 0.25, 0.25 + 47 * (0.25 / 255))
    v = np.zeros_like(x)
    v[:, 0] = 12.0   # 0.6 cells per step at res 512: particles cross a cut every other step
    return x, v


def _nccl_worker(rank, world, port, dim, model, res, steps, rebalance, native, bounds, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        x, v = thin_block_512() if res == 512 else moving_block(dim, res)
        sim = slab.SlabSimulation(x, model, res, device=rank, rebalance_every=rebalance, native=native, v=v, bounds=bounds)
        sim.advance(steps)
        n_local = sim.num_local()          # device-driven step: synchronises and raises on a latched step error
        out = sim.particles(dst=0)
        if rank == 0:
            q.put(("ok", out, sim.migrated))
    except BaseException as e:  # noqa: BLE001 - a failing rank must not leave the parent waiting for the queue
        q.put(("error", f"rank {rank}: {e!r}", 0))
        os._exit(1)             # do not wait in destroy_process_group for ranks that are stuck behind this one
    dist.destroy_process_group()


@pytest.mark.parametrize("dim,model,rebalance,native,res", [
    (3, co.JELLY, 0, True, 32), (3, co.SNOW, 2, True, 32), (2, co.LIQUID, 3, True, 64), (3, co.JELLY, 2, False, 32),
    (2, co.SNOW, 0, False, 64), (3, co.JELLY, 3, True, 512), (3, co.SNOW, 0, True, 512), (3, co.LIQUID, 0, True, -32)])
def test_slab_simulation_over_nccl(dim, model, rebalance, native, res):
    """native=True: the device-driven protocol inside libnmpm (NCCL from C++); False: slab.py over torch.distributed.
    Every GPU of the box takes part (up to 8); res 512 = the cfg4-shaped thin-slab case."""
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    bounds = None
    if res < 0:   # res -32: some slabs start EMPTY (2 ranks: the block sits left of the cut and drifts into the empty slab;
        res = -res  # more ranks: evenly spaced cuts, the block covers planes 9..18 only)
        bounds = [0, 19, res + 1] if world == 2 else [round(k * (res + 1) / world) for k in range(world)] + [res + 1]
    # res 512 with the reference's default E / volume = 1 is far beyond the stable time step: perturbations of 1 ulp grow
    # tenfold per step (C at step 4 is off by its own magnitude; single GPU against the oracle just the same) — 3 steps, like 3D snow
    steps = 10 if dim == 2 else (3 if (model == co.SNOW or res == 512) else 6)
    if res == 512 and model == co.SNOW:
        steps = 2  # amp ~ 1e5 (tests/test_parity_fullres_gpu.py): a 1e-3 difference in F after step 2 is a clamp-speed
                   # difference in v at step 3 — free-running 3D snow decorrelates one step earlier than at res 32
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, dim, model, res, steps, rebalance, native, bounds, q))
             for r in range(world)]
    for p in procs:
        p.start()
    try:
        status, got, migrated = q.get(timeout=180)
    except Exception:
        for p in procs:
            p.kill()
        raise
    if status != "ok":
        for p in procs:
            p.kill()
        pytest.fail(got)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert migrated > 0
    x, v = thin_block_512() if res == 512 else moving_block(dim, res)
    ref = co.CpuSim(x, model, res, v=v)
    ref.advance(steps)
    check(got, ref.particles(), steps, res, model)

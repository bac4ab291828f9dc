#!/usr/bin/env python
"""Debug driver (torchrun): thin res-512 block on N slabs vs the single-GPU path, step by step."""
import os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, torch.distributed as dist
import nuclearmpm_b200 as nm
from nuclearmpm_b200 import slab

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
reb = int(sys.argv[1]) if len(sys.argv) > 1 else 3
model = int(sys.argv[2]) if len(sys.argv) > 2 else 1
x = nm.cube(3, 48, 0.25, 0.25 + 47 * (0.25 / 255))
v = np.zeros_like(x); v[:, 0] = 12.0
sim = slab.SlabSimulation(x, model, 512, device=local, rebalance_every=reb, v=v)
one = nm.MPMSimulation(x, model, 512, v=v, device=local) if rank == 0 else None
for step in range(1, 13):
    sim.advance(1)
    got = sim.particles(dst=0)
    import ctypes as ct
    np_, ns_ = ct.c_longlong(), ct.c_longlong()
    sim.engine._L.nmpm_slab_counts(sim.engine._h, ct.byref(np_), ct.byref(ns_))
    cnt = torch.tensor([float(np_.value), float(ns_.value)], device="cuda", dtype=torch.float64)
    allc = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt)
    if rank == 0:
        one.advance(1)
        ref = one.particles()
        dx = np.abs(got["x"] - ref["x"]).max(axis=1)
        bad = np.nonzero(dx > 1e-5)[0]
        print(f"step {step}: bounds {sim.bounds} counts {[c.tolist() for c in allc]} sum {sum(c[0].item() for c in allc)} "
              f"max dx {dx.max():.3e} n_bad {len(bad)} migrated {sim.migrated}", flush=True)
        if len(bad):
            bx = slab.base_x(ref["x"][bad], 512)
            print("   bad base.x histogram:", dict(zip(*np.unique(bx, return_counts=True))), flush=True)
dist.destroy_process_group()

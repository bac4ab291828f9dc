#!/usr/bin/env python
"""Where does a slab step spend its time?  torchrun --nproc-per-node N tools/slab_profile.py [--workload cfg4]
Serialises the segments of slab.SlabSimulation.step with device syncs (so: GPU + host cost per segment),
then reports the free-running step time for comparison."""
import argparse
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist

import bench
from nuclearmpm_b200 import slab

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg4")
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x, model, res, desc = bench.scene(a.workload)
sim = slab.SlabSimulation(x, model, res, device=local, native=False)
sim.advance(5)
torch.cuda.synchronize()
seg = {}


def timed(name, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    h, g = seg.get(name, (0.0, 0.0))
    seg[name] = (h + t1 - t0, g + t2 - t0)


e = sim.engine
for _ in range(a.steps):
    timed("p2g (sort+clear+P2G)", e.p2g)
    timed("exchange A (planes)", sim._exchange_planes)
    timed("grid_op+G2P+pack", lambda: e.grid_g2p(sim.send_left, sim.send_right, sim.cap_records, sim.counts))
    timed("exchange B (counts+migrants+unpack)", sim._exchange_migrants)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
sim.advance(a.steps)
torch.cuda.synchronize()
dist.barrier()
free = (time.perf_counter() - t0) / a.steps
# the same steps through the native (in-library NCCL) driver
del sim
simn = slab.SlabSimulation(x, model, res, device=local, native=True)
simn.advance(5)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
simn.advance(a.steps)
torch.cuda.synchronize()
dist.barrier()
free_native = (time.perf_counter() - t0) / a.steps
sim = simn
if rank == 0:
    print(f"{desc}; world {dist.get_world_size()}; local particles {sim.num_local()}")
    for k, (h, g) in seg.items():
        print(f"  {k:40s} host-issue {1e3 * h / a.steps:7.3f} ms   issue+device {1e3 * g / a.steps:7.3f} ms")
    print(f"  free-running step: python driver {1e3 * free:.3f} ms, native driver {1e3 * free_native:.3f} ms")
dist.destroy_process_group()

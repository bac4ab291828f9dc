#!/usr/bin/env python
"""Short driver for ncu: a few advance() steps of one bench workload (no timing, no CPU baseline).

    ncu ... python tools/profile_step.py --workload snow128 --warmup 3 --steps 3 [--advect 200]
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import nuclearmpm_b200 as nm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="snow128")
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--sort-every", type=int, default=1)
ap.add_argument("--p2g-variant", type=int, default=0)
a = ap.parse_args()
x, model, res, desc = bench.scene(a.workload)
sim = nm.MPMSimulation(x, model, res, sort_every=a.sort_every, p2g_variant=a.p2g_variant)
sim.advance(a.warmup, sync=True)
sim.advance(a.steps, sync=True)
print(desc, "launches", sim.launch_count())

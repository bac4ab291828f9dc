#!/usr/bin/env python
"""How sparse is the grid of a bench workload at a given step?  Non-zero nodes, their bounding box, and the share of
4^3 / 8^3 node tiles that hold at least one non-zero node (what an active-tile clear / grid_op would touch).

    python tools/grid_occupancy.py --workload cfg4 --steps 40 300
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import nuclearmpm_b200 as nm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg4")
ap.add_argument("--steps", type=int, nargs="+", default=[40, 300])
a = ap.parse_args()
x, model, res, desc = bench.scene(a.workload)
sim = nm.MPMSimulation(x, model, res)
n1 = res + 1
done = 0
for target in a.steps:
    sim.advance(target - done, sync=True)
    done = target
    gm = sim.grid()[1].reshape(n1, n1, n1) != 0          # mass > 0 <=> the node was written by P2G and kept by grid_op
    nz = int(gm.sum())
    idx = np.nonzero(gm.any(axis=(1, 2)))[0], np.nonzero(gm.any(axis=(0, 2)))[0], np.nonzero(gm.any(axis=(0, 1)))[0]
    lo = [int(i.min()) for i in idx]
    hi = [int(i.max()) for i in idx]
    box = int(np.prod([h - l + 1 for l, h in zip(lo, hi)]))
    out = {"workload": desc, "step": target, "nodes": n1 ** 3, "nonzero_nodes": nz, "box": [lo, hi], "box_nodes": box,
           "nonzero_share_of_box": nz / box}
    for t in (4, 8):
        pad = (-n1) % t
        g = np.pad(gm, ((0, pad),) * 3)
        m = g.shape[0] // t
        tiles = g.reshape(m, t, m, t, m, t).any(axis=(1, 3, 5))
        bt = np.prod([(h // t) - (l // t) + 1 for l, h in zip(lo, hi)])
        out[f"active_tiles_{t}"] = int(tiles.sum())
        out[f"active_tile_nodes_share_of_box_{t}"] = float(tiles.sum() * t ** 3 / box)
        out[f"box_tiles_{t}"] = int(bt)
    print(json.dumps(out))

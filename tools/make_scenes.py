#!/usr/bin/env python
"""Scene list for BASELINE.json config 5: 64 randomized two-cube 2D scenes.

Per scene s: rng = numpy.random.default_rng(1234+s); two squares [a_k, a_k+0.2]^2 with a_k ~ U(0.1,0.7)
(CLI: --cube<k>-x a_k --cube<k>-y a_k+0.2, Q15), material cycled jelly/snow/liquid, E ~ U(500,5000),
nu ~ U(0.2,0.4), --cube-res 25 (1 250 particles), --steps N [--dump].
    python tools/make_scenes.py [--scenes 64] [--steps 4000] [--dump] > scenes.txt
"""
import argparse

import numpy as np

MATERIALS = ("jelly", "snow", "liquid")


def scene_flags(s: int, steps: int, dump: bool) -> str:
    rng = np.random.default_rng(1234 + s)
    a = rng.uniform(0.1, 0.7, size=2)
    E = rng.uniform(500, 5000)
    nu = rng.uniform(0.2, 0.4)
    f = [f"--steps {steps}", "--cubes 2", "--cube-res 25", f"--cube0-x {a[0]:.6f}", f"--cube0-y {a[0] + 0.2:.6f}",
         f"--cube1-x {a[1]:.6f}", f"--cube1-y {a[1] + 0.2:.6f}", f"--material-model {MATERIALS[s % 3]}", f"--E {E:.3f}",
         f"--nu {nu:.5f}"]
    if dump:
        f.append("--dump")
    return " ".join(f)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=64)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--dump", action="store_true")
    a = ap.parse_args()
    for s in range(a.scenes):
        print(scene_flags(s, a.steps, a.dump))

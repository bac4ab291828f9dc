"""Reader for the `tmp/{step}_{key}.txt` dumps of nuclear_mpm_solver --dump.

Restates the on-disk contract the reference's post-processor relies on (python/ioutils.py:32-102):
natural sort of the file names (:76-79), `"{n}_{key}.txt"` split (:89-95), and the per-key reshapes
x,v -> (N,2); F,C -> (N,2,2); Jp, timestep, lame as loaded; mass -> (65,65); velocity -> (65,65,2) (:35-73).
`load_tmp` returns {"{n}_": {key: array}} — the same keys as the reference's `results.pickle`, with plain
dicts instead of SimResult objects.  `load_step_bin` reads the --dump-bin side channel.
"""
from __future__ import annotations

import os
import re
from collections import defaultdict

import numpy as np

GRID_N = 65  # hard-coded in the reference (python/ioutils.py:62,68)


def _natural_key(s: str):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s)]


def _shape(key: str, a: np.ndarray) -> np.ndarray:
    if key in ("x", "v"):
        return a.reshape(len(a) // 2, 2)
    if key in ("F", "C"):
        return a.reshape(a.shape[0] // 2, 2, 2)
    if key == "mass":
        return a.reshape(GRID_N, GRID_N)
    if key == "velocity":
        return a.reshape(GRID_N, GRID_N, 2)
    if key in ("Jp", "timestep", "lame"):
        return a
    raise KeyError(f"ValueKey {key} is invalid")


def load_tmp(tmp: str) -> dict:
    names = sorted((f for f in os.listdir(tmp) if f.endswith(".txt") and "_" in f and f.split("_")[0].isdigit()),
                   key=_natural_key)
    results: dict = defaultdict(dict)
    for fname in names:
        n, end = fname.split("_")
        key, _ = end.split(".")
        results[f"{n}_"][key] = _shape(key, np.loadtxt(os.path.join(tmp, fname)))
    return dict(results)


def load_step_bin(tmp: str, step: int) -> dict:
    """`{step}_particles.bin`: raw Particle<2> records (src/nclr.h:20-48), 16 words each."""
    rec = np.fromfile(os.path.join(tmp, f"{step}_particles.bin"), np.float32).reshape(-1, 16)
    return dict(x=rec[:, 0:2], v=rec[:, 2:4], F=rec[:, 4:8].reshape(-1, 2, 2), C=rec[:, 8:12].reshape(-1, 2, 2),
                Jp=rec[:, 12], mass=rec[:, 13], volume=rec[:, 14], c=rec[:, 15].view(np.int32))

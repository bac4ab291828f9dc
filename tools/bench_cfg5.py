#!/usr/bin/env python
"""BASELINE.json config 5: headless `nuclear_mpm_solver` batch of 64 randomized two-cube 2D scenes (dataset generation).

    python tools/bench_cfg5.py [--scenes 64] [--steps 1000] [--dump-steps 50]

Metric: scenes * particles * steps / wall second of the whole CLI process (scene set-up, stepping,
and with --dump the reference-format text snapshots of every step), once without and once with --dump.  The CPU figure
to hold against it is bench.py's `cpu_baseline` on cfg1 (the same kind of scene: 2D, 64^2 grid, reference header, 1 core).
"""
import argparse
import json
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tools import build_cli, make_scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=64)
ap.add_argument("--steps", type=int, default=4000)  # BASELINE.json config 5: --steps 4000
ap.add_argument("--dump-steps", type=int, default=50)
a = ap.parse_args()
exe = build_cli.build()
N = 2 * 25 * 25  # --cubes 2 --cube-res 25
out = {}
# batched (the default: scenes of one model in ONE simulation) and, for comparison, one simulation per scene
for tag, steps, dump, batch in (("no_dump", a.steps, False, "1"), ("dump", a.dump_steps, True, "1"),
                                ("no_dump_per_scene_sims", a.steps, False, "0")):
    tmp = Path(tempfile.mkdtemp(prefix="nmpm_cfg5_"))
    (tmp / "scenes.txt").write_text("\n".join(make_scenes.scene_flags(s, steps, dump) for s in range(a.scenes)) + "\n")
    t0 = time.perf_counter()
    import os
    r = subprocess.run([str(exe), "--scenes", str(tmp / "scenes.txt"), "--out-dir", str(tmp / "out")], capture_output=True,
                       text=True, cwd=tmp, env=dict(os.environ, NMPM_CLI_TIMING="1", NMPM_CLI_BATCH=batch))
    dt = time.perf_counter() - t0
    inner = {}
    for ln in r.stderr.splitlines():
        if ln.startswith("{") and "run_s" in ln:
            inner = json.loads(ln)
    if r.returncode != 0:
        raise SystemExit(f"{tag}: CLI failed: {r.stderr[-500:]}")
    files = sum(1 for _ in (tmp / "out").rglob("*.txt")) if dump else 0
    out[tag] = {"batched": batch == "1", "scenes": a.scenes, "particles_per_scene": N, "steps": steps, "wall_s": dt,
                "value": a.scenes * N * steps / dt, "unit": "particle-steps/s", "snapshot_files": files,
                "setup_s": inner.get("setup_s"), "run_s": inner.get("run_s"),
                "value_stepping_only": (a.scenes * N * steps / inner["run_s"]) if inner.get("run_s") else None,
                "note": "value = whole CLI process (CUDA start-up, 64 sims, graph capture, stepping[, text snapshots]); "
                        "value_stepping_only = the advance loop alone"}
    shutil.rmtree(tmp, ignore_errors=True)

print(json.dumps({"workload": "cfg5: nuclear_mpm_solver --scenes (64 two-cube 2D scenes, 1250 p each, 64^2 grid)", **out}))

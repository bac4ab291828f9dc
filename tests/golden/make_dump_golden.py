#!/usr/bin/env python
"""Generate tests/golden/dump_ref/ : the files the REFERENCE's own headless solver writes with --dump.

Builds /root/reference/src/solver.cpp UNMODIFIED (copied to a temp dir only so that `#include "nclr.h"` resolves
through -I) against the reference's src/nclr.h, its flags submodule and oracle/eigen_standin, strict FP, and runs

    nuclear_mpm_solver --steps 3 --cubes 2 --cube-res 4 --cube0-x 0.4 --cube0-y 0.6 --cube1-x 0.15 --cube1-y 0.3
                       --material-model snow --E 2500 --nu 0.25 --dump

Only runs where /root/reference exists (this container).  The fixture pins the on-disk contract of SURVEY.md
§8(f) N1/N2: file names, one value per line, Eigen's default matrix formatting, the `timestep`/`lame` rules
and the 64-stride grid dump (Q12, Q13).
"""
import shutil
import subprocess
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
OUT = ROOT / "tests" / "golden" / "dump_ref"
ARGS = ["--steps", "3", "--cubes", "2", "--cube-res", "4", "--cube0-x", "0.4", "--cube0-y", "0.6", "--cube1-x", "0.15",
        "--cube1-y", "0.3", "--material-model", "snow", "--E", "2500", "--nu", "0.25", "--dump"]


def build_reference_solver(workdir: Path) -> Path:
    shutil.copyfile(REF / "src" / "solver.cpp", workdir / "solver.cpp")
    exe = workdir / "ref_solver"
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", f"-I{ROOT / 'oracle' / 'eigen_standin'}",
                    f"-I{REF / 'src'}", f"-I{REF / 'flags' / 'include'}", str(workdir / "solver.cpp"), "-o", str(exe)],
                   check=True)
    return exe


def main():
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        exe = build_reference_solver(td)
        out = subprocess.run([str(exe), *ARGS], cwd=td, check=True, capture_output=True, text=True).stdout
        if OUT.exists():
            shutil.rmtree(OUT)
        shutil.copytree(td / "tmp", OUT)
        (OUT / "STDOUT.txt").write_text(out)
        (OUT / "ARGS.txt").write_text(" ".join(ARGS) + "\n")
        help_out = subprocess.run([str(exe), "--help", "--cube0-x", "0.4", "--cube0-y", "0.6", "--steps", "0"], cwd=td,
                                  check=True, capture_output=True, text=True).stdout
        (OUT / "HELP.txt").write_text(help_out)
    print("wrote", OUT, sum(1 for _ in OUT.iterdir()), "files")


if __name__ == "__main__":
    main()

"""Build cli/bin/nuclear_mpm_solver (host C++17 on include/nclr.h, linked against libnmpm.so)."""
from __future__ import annotations

import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "cli" / "nuclear_mpm_solver.cpp"
EXE = ROOT / "cli" / "bin" / "nuclear_mpm_solver"
LIBDIR = ROOT / "nuclearmpm_b200" / "lib"


def stale() -> bool:
    if not EXE.exists():
        return True
    t = EXE.stat().st_mtime
    return any(p.stat().st_mtime > t for p in (SRC, ROOT / "include" / "nclr.h", ROOT / "include" / "nmpm.h"))


def build(force: bool = False) -> Path:
    if not force and not stale():
        return EXE
    EXE.parent.mkdir(parents=True, exist_ok=True)
    # $ORIGIN-relative rpath: the binary finds libnmpm.so wherever the repo snapshot lands (GPU box)
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", f"-I{ROOT / 'include'}", str(SRC), "-o", str(EXE),
                    f"-L{LIBDIR}", "-lnmpm", "-pthread", "-Wl,-rpath,$ORIGIN/../../nuclearmpm_b200/lib"], check=True)
    return EXE


if __name__ == "__main__":
    print(build(force=True))

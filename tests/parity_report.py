"""Print the measured one-step parity errors of the CUDA path against the golden vectors
(teacher-forced from reference states) next to the stated tolerances.  GPU box only; test infrastructure (it uses
the oracle for the cell keys), hence under tests/.

    python tests/parity_report.py > profiles/rNN_parity.md
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))  # conftest helpers / test_parity_gpu tolerances
import nuclearmpm_b200 as nm  # noqa: E402
from oracle import cpu_oracle as co  # noqa: E402
from test_parity_gpu import TOL_C, TOL_F, TOL_GRID_V, TOL_JP, TOL_V, TOL_X  # noqa: E402

FIELDS = ("x", "v", "F", "C", "Jp")


def errs(got, ref):
    vmax = max(1.0, float(np.abs(ref["v"]).max()))
    cmax = max(1.0, float(np.abs(ref["C"]).max()))
    return dict(x=np.abs(got["x"] - ref["x"]).max(), v=np.abs(got["v"] - ref["v"]).max() / vmax,
                C=np.abs(got["C"] - ref["C"]).max() / cmax, F=np.abs(got["F"] - ref["F"]).max(),
                Jp=np.abs(got["Jp"] - ref["Jp"]).max(), vmax=vmax, cmax=cmax)


def main():
    print("# One-step parity of the CUDA path vs the reference (golden vectors from the reference header)\n")
    print(f"tolerances: x {TOL_X:g}, v {TOL_V:g}·max(1,|v|max), C {TOL_C:g}·max(1,|C|max), F {TOL_F:g}, Jp {TOL_JP:g}, "
          f"grid v {TOL_GRID_V:g}·max(1,|v|max)\n")
    print("| dim | model | from→to | Δx | Δv/vmax | ΔC/cmax | ΔF | ΔJp | |v|max | |C|max | cell-key mismatches |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for dim in (2, 3):
        for model in (co.SNOW, co.JELLY, co.LIQUID):
            g = dict(np.load(ROOT / "tests" / "golden" / f"scene_{dim}d_{co.MODEL_NAMES[model]}.npz"))
            res = int(g["res"])
            sim = nm.MPMSimulation(g["x0"], model, res)
            for a, b in [(0, 1), (1, 2), (2, 3), (100, 101)]:
                if a:
                    sim.upload(*[g[f"s{a}_{k}"] for k in FIELDS])
                d = sim.sort_debug()
                xa = g["x0"] if a == 0 else g[f"s{a}_x"]
                _, keys, _ = co.cell_keys(xa, res, 1, d["tile_bits"])
                mism = int((d["keys"] != keys).sum())
                sim.advance(1)
                e = errs(sim.particles(), {k: g[f"s{b}_{k}"] for k in FIELDS})
                print(f"| {dim} | {co.MODEL_NAMES[model]} | {a}→{b} | {e['x']:.2e} | {e['v']:.2e} | {e['C']:.2e} | "
                      f"{e['F']:.2e} | {e['Jp']:.2e} | {e['vmax']:.3g} | {e['cmax']:.3g} | {mism} |")


if __name__ == "__main__":
    main()

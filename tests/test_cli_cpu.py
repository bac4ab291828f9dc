"""nuclear_mpm_solver (cli/) without a GPU: it builds, parses flags with the reference's rules
(flags/include/flags.h), prints the reference's help text, refuses to compute on the host;
and the dump reader reproduces the reference post-processor's view of the reference's own dump."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden" / "dump_ref"
sys.path.insert(0, str(ROOT))

from nuclearmpm_b200 import dumpio  # noqa: E402
from tools import build_cli  # noqa: E402


@pytest.fixture(scope="module")
def exe():
    import nuclearmpm_b200 as nm
    nm.load_library()
    return build_cli.build()


def parse(exe, *args):
    r = subprocess.run([str(exe), "--parse-only", *args], capture_output=True, text=True)
    return r.returncode, r.stdout.strip()


def test_flag_rules_match_the_reference_parser(exe):
    rc, out = parse(exe, "--cube0-x", "0.4", "--cube0-y", "0.6")
    assert rc == 0 and "steps=1000 cubes=1 cube_res=25 dim=2 E=1000 nu=0.3 gravity=-100 material=jelly dump=0" in out  # Q14
    assert "any=0" in out  # only cube coordinates given: the reference prints the help text
    # a token starting with '-' is an option: `--gravity -100` leaves gravity WITHOUT a value -> default
    rc, out = parse(exe, "--gravity", "-50", "--cube0-x", "0.4", "--cube0-y", "0.6")
    assert "gravity=-100" in out
    rc, out = parse(exe, "--gravity=-50", "--cube0-x", "0.4", "--cube0-y", "0.6")
    assert "gravity=-50" in out
    # the first occurrence of a key wins
    rc, out = parse(exe, "--steps", "5", "--steps", "9", "--cube0-x", "0.4", "--cube0-y", "0.6")
    assert "steps=5 " in out
    # booleans: present = true unless the value is one of 0/n/no/f/false
    for val, want in (("", 1), ("no", 0), ("false", 0), ("0", 0), ("yes", 1), ("f", 0)):
        rc, out = parse(exe, "--cube0-x", "0.4", "--cube0-y", "0.6", "--dump", *([val] if val else []))
        assert f"dump={want}" in out, (val, out)
    # unparsable numbers fall back to the default
    rc, out = parse(exe, "--steps", "abc", "--E", "x1", "--cube0-x", "0.4", "--cube0-y", "0.6")
    assert "steps=1000" in out and "E=1000" in out
    # every cube needs both coordinates (src/solver.cpp:140-143)
    rc, out = parse(exe, "--cubes", "2", "--cube0-x", "0.4", "--cube0-y", "0.6")
    assert rc == 1 and "Cube: 1 is missing coordinates" in out
    rc, out = parse(exe, "--material-model", "snow", "--cubes", "2", "--cube0-x", "0.4", "--cube0-y", "0.6", "--cube1-x=0.1",
                    "--cube1-y=0.3")
    assert rc == 0 and "material=snow" in out and "cube=(0.4,0.6) cube=(0.1,0.3)" in out  # Q15: (min,max) pairs


def test_invalid_material_and_help_text(exe):
    r = subprocess.run([str(exe), "--material-model", "sand"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("Invalid Option: sand")
    want = (GOLD / "HELP.txt").read_text()
    help_text = want[:want.index("Running simulation")]
    assert r.stdout == help_text  # byte for byte the reference's help (src/solver.cpp:25-43)


def test_no_cpu_fallback(exe, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([str(exe), "--steps", "2", "--cube0-x", "0.4", "--cube0-y", "0.6", "--dump"], capture_output=True,
                       text=True, cwd=tmp_path)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    assert not (tmp_path / "tmp").exists()


def test_dump_reader_on_the_reference_dump():
    res = dumpio.load_tmp(str(GOLD))
    assert sorted(res) == ["0_", "1_", "2_"]
    n = 32  # two 4x4 squares
    for k, step in res.items():
        assert step["x"].shape == (n, 2) and step["v"].shape == (n, 2)
        assert step["F"].shape == (n, 2, 2) and step["C"].shape == (n, 2, 2)
        assert step["Jp"].shape == (n,) and step["timestep"].shape == (n,) and step["lame"].shape == (n, 2)
        assert step["mass"].shape == (65, 65) and step["velocity"].shape == (65, 65, 2)
        assert np.allclose(step["lame"], 10000.0)  # Q13: E=2500, nu=0.25 -> mu0=lambda0=1000, x kSnowHardening
    assert np.allclose(res["0_"]["timestep"], 1e-4) and np.allclose(res["1_"]["timestep"], 1e-4)  # src/solver.cpp:92
    assert np.allclose(res["2_"]["timestep"], 2e-4)
    assert res["0_"]["mass"].sum() == 0 and res["1_"]["mass"].sum() > 0  # zero grid at step 0 (src/solver.cpp:54-57)
    assert np.allclose(res["0_"]["F"], np.eye(2))


def test_dump_reader_agrees_with_the_reference_post_processor(tmp_path, monkeypatch):
    ref = Path("/root/reference/python")
    if not (ref / "ioutils.py").exists():
        pytest.skip("reference tree not present (GPU box)")
    pytest.importorskip("loguru")
    pytest.importorskip("tqdm")
    import importlib.util
    import pickle
    spec = importlib.util.spec_from_file_location("ref_ioutils", ref / "ioutils.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["ref_ioutils"] = mod  # so that the pickled SimResult objects resolve
    only_txt = tmp_path / "tmp"
    only_txt.mkdir()
    for f in GOLD.glob("[0-9]*_*.txt"):
        (only_txt / f.name).write_bytes(f.read_bytes())
    monkeypatch.chdir(tmp_path)
    mod.process_tmp(str(only_txt))  # writes results.pickle in the cwd (python/ioutils.py:100-101)
    with open(tmp_path / "results.pickle", "rb") as fh:
        ref_res = pickle.load(fh)
    mine = dumpio.load_tmp(str(only_txt))
    assert sorted(ref_res) == sorted(mine)
    for k in mine:
        for key, arr in mine[k].items():
            assert np.array_equal(np.asarray(ref_res[k].__dict__[key]), arr), (k, key)

#!/bin/bash
# Iteration visit: parity tests, bench lines, ncu --set full of P2G/G2P in the bench regime (dispersed scene, sort cadence 4).
# bash tools/gpu_iter.sh <tag> [ncu_skip_steps]
tag=${1:-it}; skip=${2:-28}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
for w in cfg4 snow128; do
  timeout 600 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > $out/bench_$w.json 2> $out/bench_$w.err; echo "bench $w rc=$?"
  python - <<PY
import json
d=json.load(open("$out/bench_$w.json")); r=d["roofline"]
print("$w", "%.3e p-steps/s"%d["value"], "ms/step %.3f"%d["ms_per_step"], {k:round(v,4) for k,v in r["phase_ms"].items()}, "frac", round(r["frac"],3), r["kernel"], "e2e %.3e"%d["e2e"]["value"])
PY
done
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_p2g_cell|k_g2p_gather' -s $((2*skip)) -c 8 \
  -f -o $out/prof_cfg4_late python tools/profile_step.py --workload cfg4 --warmup $skip --steps 6 --sort-every 4 > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log

#!/bin/bash
# Quick GPU visit: parity tests + bench lines (+ optional sort cadence sweep).  bash tools/gpu_quick.sh <tag> [workloads...]
tag=${1:-q}; shift
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest_gpu.log
for w in ${@:-snow128 cfg4}; do
  timeout 600 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > $out/bench_$w.json 2> $out/bench_$w.err; echo "bench $w rc=$?"
  python - <<PY
import json
d=json.load(open("$out/bench_$w.json")); r=d["roofline"]
print("$w", "%.3e p-steps/s"%d["value"], "ms/step %.3f"%d["ms_per_step"], {k:round(v,4) for k,v in r["phase_ms"].items()}, "frac", round(r["frac"],3), r["kernel"], "e2e %.3e"%d["e2e"]["value"])
PY
done

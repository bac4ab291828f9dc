#!/bin/bash
# Round-end visit: parity tests, the default bench line + the other workloads, ncu launch list and full captures.
# bash tools/gpu_final.sh <tag> [light]   (light: no small-kernel capture, no cfg5)
tag=${1:-final}; light=${2:-}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi > $out/nvidia-smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
timeout 600 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench default rc=$?"; tail -c 700 $out/bench_default.json
for w in snow128 cfg2 cfg3 cfg1; do
  timeout 600 python bench.py --workload $w --steps 50 --warmup 5 > $out/bench_$w.json 2> $out/bench_$w.err; echo "bench $w rc=$?"
  python - $out/bench_$w.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(d["config"]["workload"][:40], "%.3e p-steps/s"%d["value"], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in r["phase_ms"].items()}, "frac", round(r["frac"],3), r["kernel"], "e2e %.3e"%d["e2e"]["value"], "cpu %.3e"%(d["cpu_baseline"] or {}).get("value",0))
except Exception as e:
    print("FAILED", e)
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_cfg4.csv \
  python tools/profile_step.py --workload cfg4 --warmup 28 --steps 8 --sort-every 4 > $out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_p2g|k_g2p_gather' -s 56 -c 8 \
  -f -o $out/prof_cfg4_late python tools/profile_step.py --workload cfg4 --warmup 28 --steps 6 --sort-every 4 > $out/ncu_full.log 2>&1
if [ -z "$light" ]; then timeout 900 ncu --set full --clock-control none -k regex:'k_grid_op|k_clear_box|sort_scatter|sort_hist' -s 70 -c 6 \
  -f -o $out/prof_cfg4_small python tools/profile_step.py --workload cfg4 --warmup 28 --steps 4 --sort-every 4 >> $out/ncu_full.log 2>&1
fi
tail -2 $out/ncu_full.log
ls -la $out | head -30
if [ -z "$light" ]; then timeout 600 python tools/bench_cfg5.py > $out/cfg5.json 2> $out/cfg5.err; echo "cfg5 rc=$?"; cat $out/cfg5.json
fi

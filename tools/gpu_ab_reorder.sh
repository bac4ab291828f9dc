out=gpurun_out/r02d; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
for w in cfg4 snow128 cfg3; do for r in 1; do
  NMPM_LOCAL_REORDER=$r timeout 600 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > $out/bench_${w}_lr$r.json 2> $out/bench_${w}_lr$r.err
  python - $out/bench_${w}_lr$r.json "$w reorder=$r" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(sys.argv[2], "%.3e p-steps/s"%d["value"], "ms/step %.3f"%d["ms_per_step"], {k:round(v,4) for k,v in r["phase_ms"].items()})
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
done; done

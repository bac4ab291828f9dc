#!/bin/bash
# A/B visit: parity tests, then bench lines per (P2G variant : sort cadence [: alternative build]) combination,
# then one ncu capture.  Alternative builds are nuclearmpm_b200/lib/exp/libnmpm_<name>.so (same C-ABI, NMPM_LIB).
# bash tools/gpu_ab.sh <tag> "<variant:cadence[:build] ...>" [ncu_variant] [ncu_build] [ncu_skip_steps] [workloads]
tag=${1:-ab}; combos=${2:-"0:4"}; nv=${3:-0}; nb=${4:-}; skip=${5:-28}; wl=${6:-"cfg4 snow128"}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu.log
summ() {
  python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(sys.argv[2], "%.3e p-steps/s"%d["value"], "ms/step %.3f"%d["ms_per_step"], {k:round(v,4) for k,v in r["phase_ms"].items()}, "frac", round(r["frac"],3), r["kernel"], "e2e %.3e"%d["e2e"]["value"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for w in $wl; do
  for vc in $combos; do
    IFS=: read v c b <<< "$vc"
    f=$out/bench_${w}_v${v}_s${c}${b:+_$b}.json
    lib=; [ -n "$b" ] && lib=$PWD/nuclearmpm_b200/lib/exp/libnmpm_$b.so
    NMPM_LIB=$lib timeout 600 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu --p2g-variant $v --sort-every $c > $f 2> ${f%.json}.err
    summ $f "$w v$v s$c $b"
  done
done
if [ "$nv" != "none" ]; then
  lib=; [ -n "$nb" ] && lib=$PWD/nuclearmpm_b200/lib/exp/libnmpm_$nb.so
  NMPM_LIB=$lib timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_p2g_cell|k_g2p_gather' -s $((2*skip)) -c 8 \
    -f -o $out/prof_cfg4_late python tools/profile_step.py --workload cfg4 --warmup $skip --steps 6 --sort-every 4 --p2g-variant $nv > $out/ncu_full.log 2>&1
  tail -2 $out/ncu_full.log
fi

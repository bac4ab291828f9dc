#!/bin/bash
# Scaling visit on an N-GPU box: bash tools/gpu_scale.sh <tag> "<list of N>"
tag=${1:-s}; ns=${2:-"1 2 4 8"}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L | wc -l
for n in $ns; do
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu > $out/scale_$n.json 2> $out/scale_$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 30 --warmup 5 > $out/scale_$n.json 2> $out/scale_$n.err
  fi
  echo "N=$n rc=$?"; tail -2 $out/scale_$n.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("$out/scale_$n.json").read().strip().splitlines()[-1]); r=d.get("roofline") or {}
    print("N=$n %.3e p-steps/s  %.3f ms/step"%(d["value"], d["ms_per_step"]), (d.get("run") or {}).get("bounds"), (d.get("run") or {}).get("particles_per_rank_min_max"), "checks", (d.get("checks") or {}).get("ok"), "verify", d.get("verification"), "roofline", r.get("kernel"), round(r.get("frac",0),3), "e2e %.3e"%d["e2e"]["value"])
    for row in (r.get("phase_ms_per_rank") or []): print("   ", {k:round(v,3) for k,v in row.items()})
except Exception as e: print("parse failed", e)
PY
done

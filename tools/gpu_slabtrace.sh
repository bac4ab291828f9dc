#!/bin/bash
# Multi-GPU visit: device-time trace of the native slab step (NMPM_SLAB_TRACE) + the plain bench line.
# bash tools/gpu_slabtrace.sh <tag> <N>
tag=${1:-st}; n=${2:-2}
out=gpurun_out/$tag
mkdir -p $out
NMPM_SLAB_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $n --steps 30 --warmup 5 > $out/trace_$n.json 2> $out/trace_$n.err
grep "slab trace" $out/trace_$n.err
bash tools/gpu_scale.sh $tag "$n"

#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL slab tests, then the N-GPU bench lines.  bash tools/gpu_multi.sh <tag> "<Ns>"
tag=${1:-m}; ns=${2:-"2"}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L > $out/gpus.txt; nvidia-smi topo -m >> $out/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -x -q > $out/pytest_slab.log 2>&1; echo "pytest slab rc=$?"; tail -5 $out/pytest_slab.log
bash tools/gpu_scale.sh $tag "$ns"

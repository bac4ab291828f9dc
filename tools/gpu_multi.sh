#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL slab tests, N-GPU bench, pipe microbenchmark.  bash tools/gpu_multi.sh <tag> "<Ns>"
tag=${1:-m}; ns=${2:-"2"}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L > $out/gpus.txt; nvidia-smi topo -m >> $out/gpus.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o $out/ubench_pipes tools/ubench_pipes.cu && $out/ubench_pipes > $out/ubench_pipes.txt 2>&1; cat $out/ubench_pipes.txt
timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -x -q > $out/pytest_slab.log 2>&1; echo "pytest slab rc=$?"; tail -5 $out/pytest_slab.log
bash tools/gpu_scale.sh $tag "$ns"

#!/bin/bash
# bash tools/gpu_bench3.sh <tag>: the bench lines quoted in README / DESIGN (default flags, the driver's 20 steps, snow128)
tag=${1:-b3}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "default rc=$?"; python tools/bench_summary.py $out/bench_default.json
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > $out/bench_s20.json 2> $out/bench_s20.err; echo "s20 rc=$?"; python tools/bench_summary.py $out/bench_s20.json
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --fuse 1 > $out/bench_s20_unfused.json 2> $out/bench_s20_unfused.err; echo "s20 unfused rc=$?"; python tools/bench_summary.py $out/bench_s20_unfused.json
timeout 900 python bench.py --workload snow128 --steps 50 --warmup 5 --no-cpu > $out/bench_snow128.json 2> $out/bench_snow128.err; echo "snow128 rc=$?"; python tools/bench_summary.py $out/bench_snow128.json
timeout 600 python -m pytest tests/test_bench_gpu.py -m gpu -q 2>&1 | tail -2

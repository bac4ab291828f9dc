#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list + full capture of the two hot kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi > $out/nvidia-smi.txt 2>&1
python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
for w in snow128 cfg4 cfg2 cfg3 cfg1; do
  timeout 600 python bench.py --workload $w --steps 50 --warmup 5 > $out/bench_$w.json 2> $out/bench_$w.err; echo "bench $w rc=$?"
  tail -c 600 $out/bench_$w.json
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_snow128.csv \
  python tools/profile_step.py --workload snow128 --warmup 3 --steps 3 > $out/ncu_launches.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_cfg4.csv \
  python tools/profile_step.py --workload cfg4 --warmup 3 --steps 3 >> $out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_p2g_cell|k_g2p_gather|k_grid_op' -s 9 -c 3 \
  -f -o $out/prof_snow128 python tools/profile_step.py --workload snow128 --warmup 3 --steps 1 > $out/ncu_full.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_p2g_cell|k_g2p_gather|k_grid_op' -s 9 -c 3 \
  -f -o $out/prof_cfg4 python tools/profile_step.py --workload cfg4 --warmup 3 --steps 1 >> $out/ncu_full.log 2>&1
ls -la $out

#!/bin/bash
# ncu visit: launch list + full capture of the hot kernels.  bash tools/gpu_prof.sh <tag> <workload>
tag=${1:-p}; w=${2:-snow128}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_$w.csv \
  python tools/profile_step.py --workload $w --warmup 3 --steps 3 > $out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_p2g_cell|k_g2p_gather' -s 6 -c 2 \
  -f -o $out/prof_$w python tools/profile_step.py --workload $w --warmup 3 --steps 1 > $out/ncu_full.log 2>&1
tail -3 $out/ncu_full.log

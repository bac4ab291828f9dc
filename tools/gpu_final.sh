#!/bin/bash
# Round-end visit (1 GPU): parity tests, smoke, the default bench line + reference arm + the other workloads, launch list,
# ncu captures, cfg5.   bash tools/gpu_final.sh <tag>
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi > $out/nvidia-smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_parity_fullres_gpu.py > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
rm -f gpurun_out/parity_fullres.md; timeout 1500 python -m pytest tests/test_parity_fullres_gpu.py -m gpu -q > $out/pytest_fullres.log 2>&1; echo "fullres rc=$?"; tail -3 $out/pytest_fullres.log; cp gpurun_out/parity_fullres.md $out/ 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench default rc=$?"; python tools/bench_summary.py $out/bench_default.json
timeout 900 python bench.py --fuse 1 --no-cpu > $out/bench_default_unfused.json 2> $out/bench_default_unfused.err; echo "bench unfused rc=$?"; python tools/bench_summary.py $out/bench_default_unfused.json
timeout 1500 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 $out/bench_ref.json
for w in snow128 cfg2 cfg3 cfg1; do
  timeout 600 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > $out/bench_$w.json 2> $out/bench_$w.err; echo "bench $w rc=$?"
  python tools/bench_summary.py $out/bench_$w.json
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_cfg4.csv \
  python tools/profile_step.py --workload cfg4 --warmup 28 --steps 8 --sort-every 4 > $out/ncu_launches.log 2>&1; echo "launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_g2p_p2g' -s 28 -c 4 \
  -f -o $out/prof_cfg4_fused python tools/profile_step.py --workload cfg4 --warmup 28 --steps 6 --sort-every 4 > $out/ncu_full.log 2>&1
NMPM_FUSE=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_p2g_streams|k_g2p_gather' -s 56 -c 8 \
  -f -o $out/prof_cfg4_unfused python tools/profile_step.py --workload cfg4 --warmup 28 --steps 6 --sort-every 4 >> $out/ncu_full.log 2>&1
NMPM_TILES=2 timeout 900 ncu --set full --clock-control none -k regex:'k_tiles3|k_mark_tiles|sort_scatter|sort_hist' -s 60 -c 8 \
  -f -o $out/prof_cfg4_small python tools/profile_step.py --workload cfg4 --warmup 28 --steps 4 --sort-every 4 >> $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log
timeout 600 python tools/bench_cfg5.py > $out/cfg5.json 2> $out/cfg5.err; echo "cfg5 rc=$?"; cut -c1-600 $out/cfg5.json
ls -la $out | head -40

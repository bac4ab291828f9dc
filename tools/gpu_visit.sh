#!/bin/bash
# GPU visit: bash tools/gpu_visit.sh <tag> [steps...]   steps: test fullres fusedtest tiletest batchtest smoke bench quick
#   quickoff (unfused) notiles alwaystiles minb6 abminb ref snow128|cfg1|cfg2|cfg3 launches ncu ncufused ncug2p occupancy cfg5
tag=${1:-visit}; shift
what=${@:-test bench}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/nvidia-smi.txt 2>&1
for w in $what; do
case $w in
test)    timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_parity_fullres_gpu.py > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu.log;;
fullres) rm -f gpurun_out/parity_fullres.md; timeout 1500 python -m pytest tests/test_parity_fullres_gpu.py -m gpu -q > $out/pytest_fullres.log 2>&1; echo "fullres rc=$?"; tail -15 $out/pytest_fullres.log; cp gpurun_out/parity_fullres.md $out/ 2>/dev/null;;
fusedtest) timeout 600 python -m pytest tests/test_fused_gpu.py -m gpu -q > $out/pytest_fused.log 2>&1; echo "fused pytest rc=$?"; tail -25 $out/pytest_fused.log;;
quickoff) timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --fuse 1 > $out/bench_quick_unfused.json 2> $out/bench_quick_unfused.err; echo "bench unfused rc=$?"; tail -3 $out/bench_quick_unfused.err; python tools/bench_summary.py $out/bench_quick_unfused.json;;
abminb)  for mb in 5 8; do NMPM_FUSED_MINB=$mb timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --late-step 0 > $out/bench_minb$mb.json 2> $out/bench_minb$mb.err; echo "minb $mb rc=$?"; python tools/bench_summary.py $out/bench_minb$mb.json; done;;
ncufused) timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_g2p_p2g' -s 28 -c 4 \
            -f -o $out/prof_cfg4_fused python tools/profile_step.py --workload cfg4 --warmup 28 --steps 6 --sort-every 4 > $out/ncu_fused.log 2>&1; echo "ncu rc=$?"; tail -2 $out/ncu_fused.log;;
occupancy) timeout 900 python tools/grid_occupancy.py --workload cfg4 --steps 40 150 300 > $out/grid_occupancy.jsonl 2> $out/grid_occupancy.err; echo "occupancy rc=$?"; cat $out/grid_occupancy.jsonl;;
notiles) NMPM_TILES=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $out/bench_quick_notiles.json 2> $out/bench_quick_notiles.err; echo "bench notiles rc=$?"; python tools/bench_summary.py $out/bench_quick_notiles.json;;
minb6)   NMPM_FUSED_MINB=6 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $out/bench_quick_minb6.json 2> $out/bench_quick_minb6.err; echo "bench minb6 rc=$?"; python tools/bench_summary.py $out/bench_quick_minb6.json;;
ncug2p)  NMPM_FUSE=0 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_g2p_gather' -s 30 -c 2 \
            -f -o $out/prof_cfg4_g2p python tools/profile_step.py --workload cfg4 --warmup 28 --steps 6 --sort-every 4 > $out/ncu_g2p.log 2>&1; echo "ncu rc=$?"; tail -2 $out/ncu_g2p.log;;
batchtest) timeout 600 python -m pytest tests/test_batch_gpu.py -m gpu -q -x > $out/pytest_batch.log 2>&1; echo "batch pytest rc=$?"; tail -25 $out/pytest_batch.log;;
tiletest) timeout 600 python -m pytest tests/test_tiles_gpu.py -m gpu -q -x > $out/pytest_tiles.log 2>&1; echo "tiles pytest rc=$?"; tail -25 $out/pytest_tiles.log;;
alwaystiles) NMPM_TILES=2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $out/bench_quick_alwaystiles.json 2> $out/bench_quick_alwaystiles.err; echo "bench always tiles rc=$?"; python tools/bench_summary.py $out/bench_quick_alwaystiles.json;;
smoke)   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log;;
bench)   timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?"; tail -3 $out/bench_default.err; python tools/bench_summary.py $out/bench_default.json;;
quick)   timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $out/bench_quick.json 2> $out/bench_quick.err; echo "bench rc=$?"; tail -3 $out/bench_quick.err; python tools/bench_summary.py $out/bench_quick.json;;
ref)     timeout 1500 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"; cut -c1-1500 $out/bench_ref.json;;
snow128|cfg1|cfg2|cfg3) timeout 900 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > $out/bench_$w.json 2> $out/bench_$w.err; echo "bench $w rc=$?"; python tools/bench_summary.py $out/bench_$w.json;;
launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_cfg4.csv \
            python tools/profile_step.py --workload cfg4 --warmup 28 --steps 8 --sort-every 4 > $out/ncu_launches.log 2>&1; echo "launches rc=$?";;
ncu)     timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_p2g|k_g2p' -s 56 -c 8 \
            -f -o $out/prof_cfg4 python tools/profile_step.py --workload cfg4 --warmup 28 --steps 6 --sort-every 4 > $out/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -2 $out/ncu_full.log;;
cfg5)    timeout 600 python tools/bench_cfg5.py > $out/cfg5.json 2> $out/cfg5.err; echo "cfg5 rc=$?"; cat $out/cfg5.json;;
*) echo "unknown step $w";;
esac
done
ls $out

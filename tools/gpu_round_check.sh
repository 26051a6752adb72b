#!/bin/bash
# One GPU call per kernel change: GPU tests, the default bench line (both arms), the ncu launch list
# and one `ncu --set full` capture of the distance kernel (each after the plain run has finished), raw
# page exported as CSV.  Run on the GPU box:  gpurun --timeout 2400 -- 'bash tools/gpu_round_check.sh TAG'
# Outputs land in gpurun_out/ (TAG in the names); copy what is to be kept into profiles/.
TAG=${1:-r02_final}
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pipeline --distinct 2 > gpurun_out/${TAG}_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:distance_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pipeline --distinct 2 > gpurun_out/${TAG}_ncu_b.log 2>&1
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_distance_kernel_ncu_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_distance_kernel_ncu_raw.csv "ncu --set full --clock-control none, bench.py --steps 2 --warmup 3, 4th launch" > gpurun_out/${TAG}_distance_kernel_ncu_summary.json 2>/dev/null
ls -la gpurun_out/${TAG}_prof.ncu-rep

#!/bin/bash
# One GPU call per kernel change: GPU tests, the default bench line, the ncu launch list and one
# `ncu --set full` capture of the distance kernel (each after the plain run has finished), raw page
# exported as CSV.  Run on the GPU box:  gpurun --timeout 900 -- 'bash tools/gpu_round_check.sh'
# Outputs land in gpurun_out/ (suffix below); copy what is to be kept into profiles/.
python -m pytest tests -m gpu -x -q > gpurun_out/r01_pytest_gpu_s18.log 2>&1; tail -2 gpurun_out/r01_pytest_gpu_s18.log
python bench.py > gpurun_out/r01_bench_s18.json 2> gpurun_out/bench_s18.err; tail -c 1500 gpurun_out/r01_bench_s18.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_ncu_launches_s18.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --distinct 2 > gpurun_out/ncu_s18a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:distance_kernel -s 3 -c 1 -o gpurun_out/r01_s18_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --distinct 2 > gpurun_out/ncu_s18b.log 2>&1
ncu -i gpurun_out/r01_s18_prof.ncu-rep --page raw --csv > gpurun_out/r01_distance_kernel_s18_ncu_raw.csv 2>/dev/null
ls -la gpurun_out/r01_s18_prof.ncu-rep

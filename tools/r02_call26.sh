#!/bin/bash
mkdir -p gpurun_out
FANDOM_SEARCH_TILE_GROUP=103 ncu --set full --clock-control none --import-source on -k regex:distance_kernel_n128 -s 3 -c 1 -o gpurun_out/r02_c26_prof_n128 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pipeline --distinct 2 > gpurun_out/r02_c26_ncu.log 2>&1
ncu -i gpurun_out/r02_c26_prof_n128.ncu-rep --page raw --csv > gpurun_out/r02_c26_ncu_raw_n128.csv 2>/dev/null
ls -la gpurun_out/r02_c26_prof_n128.ncu-rep
FS_NVCC_EXTRA="-DFS_TIMELINE" python -m fandom_search_b200.build --force > gpurun_out/r02_c26_build.log 2>&1

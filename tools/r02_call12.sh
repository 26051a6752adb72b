#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_device_records.py -m gpu -q -x ) > gpurun_out/r02_c12_pytest.log 2>&1
tail -3 gpurun_out/r02_c12_pytest.log
for G in 7 23; do
  FANDOM_SEARCH_TILE_GROUP=$G ncu --set full --clock-control none --import-source on -k regex:distance_kernel -s 3 -c 1 -o gpurun_out/r02_c12_prof_g$G -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pipeline --distinct 2 > gpurun_out/r02_c12_ncu_g$G.log 2>&1
  ncu -i gpurun_out/r02_c12_prof_g$G.ncu-rep --page raw --csv > gpurun_out/r02_c12_ncu_raw_g$G.csv 2>/dev/null
done
ls -la gpurun_out/r02_c12_prof_g*.ncu-rep

#!/bin/bash
# floor probes of the distance kernel: what the MMAs + barriers cost without the epilogue's work
mkdir -p gpurun_out
FS_NVCC_EXTRA=-DFS_FLOOR_PROBE python -m fandom_search_b200.build --force > gpurun_out/r02_c13_build.log 2>&1
for G in 7 23 55 119 31 63; do
  for rep in 1 2; do
  timeout 300 python tools/sweep.py --one 6 2500000 25000 300 --pair 2 --group $G --reps 20 >> gpurun_out/r02_c13_floor.jsonl 2>> gpurun_out/r02_c13.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_c13_floor.jsonl'):
    d=json.loads(l); print(d['group'], d['dim_pad'], round(d['kernel_ms'],2), round(d['windows_per_s']/1e6,1), d['clocks'])
PY
tail -3 gpurun_out/r02_c13.err

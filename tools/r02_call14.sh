#!/bin/bash
mkdir -p gpurun_out
FS_NVCC_EXTRA=-DFS_FLOOR_PROBE python -m fandom_search_b200.build --force > gpurun_out/r02_c14_build.log 2>&1
tail -3 gpurun_out/r02_c14_build.log
for W in 0 1 2 3 4 6 7; do
  for G in 23 119; do
  FS_DEBUG_WAIT=$W timeout 300 python tools/sweep.py --one 6 2500000 25000 300 --pair 2 --group $G --reps 20 2>> gpurun_out/r02_c14.err | sed "s/^{/{\"wait\": $W, /" >> gpurun_out/r02_c14_floor.jsonl
  done
done
FS_DEBUG_WAIT=4 timeout 300 python tools/sweep.py --one 6 2500000 25000 300 --pair 2 --group 7 --reps 20 2>> gpurun_out/r02_c14.err | sed "s/^{/{\"wait\": 4, /" >> gpurun_out/r02_c14_floor.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r02_c14_floor.jsonl'):
    d=json.loads(l); print('wait', d['wait'], 'group', d['group'], round(d['kernel_ms'],2), round(d['windows_per_s']/1e6,1), d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))
PY
tail -3 gpurun_out/r02_c14.err

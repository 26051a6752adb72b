#!/bin/bash
mkdir -p gpurun_out
FS_NVCC_EXTRA=-DFS_FLOOR_PROBE python -m fandom_search_b200.build --force > gpurun_out/r02_c16_build.log 2>&1
tail -1 gpurun_out/r02_c16_build.log
for W in 0 2 8 9 10 11 40 41 43 24 27; do
  for G in 7 23 119; do
  FS_DEBUG_WAIT=$W timeout 300 python tools/sweep.py --one 6 2500000 25000 300 --pair 2 --group $G --reps 20 2>> gpurun_out/r02_c16.err | sed "s/^{/{\"wait\": $W, /" >> gpurun_out/r02_c16_wait.jsonl
  done
done
python - <<'PY'
import json, collections
res=collections.defaultdict(dict)
for l in open('gpurun_out/r02_c16_wait.jsonl'):
    d=json.loads(l); res[d['wait']][d['group']]=(round(d['kernel_ms'],2), d['clocks']['sm_mhz'])
for w in res: print('wait', w, res[w])
PY
tail -3 gpurun_out/r02_c16.err

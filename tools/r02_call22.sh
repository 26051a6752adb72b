#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_c22_pytest.log 2>&1
tail -3 gpurun_out/r02_c22_pytest.log
( timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/r02_c22_bench.json 2> gpurun_out/r02_c22_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_c22_bench.json'))
print('value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1))
p=d['pipeline']; print('pipeline', round(p['value']/1e6,1), 'steady', round(p['steady_state']['value']/1e6,1), p['works'], round(p['seconds'],3), {k[:12]:round(v,3) for k,v in p['rank0_phases_s'].items()}, p['host_threads_per_rank'])
PY
for K in gather_kernel hash_probe_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -o gpurun_out/r02_c22_prof_$K -f python tools/stage_bench.py --tokens 10000000 > gpurun_out/r02_c22_ncu_$K.log 2>&1
  ncu -i gpurun_out/r02_c22_prof_$K.ncu-rep --page raw --csv > gpurun_out/r02_c22_ncu_raw_$K.csv 2>/dev/null
done
ls -la gpurun_out/r02_c22_prof_*.ncu-rep

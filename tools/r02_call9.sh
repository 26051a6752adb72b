#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_c9_pytest.log 2>&1
tail -12 gpurun_out/r02_c9_pytest.log
( time timeout 1200 python bench.py ) > gpurun_out/r02_c9_bench.json 2> gpurun_out/r02_c9_bench.err
tail -5 gpurun_out/r02_c9_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_c9_bench.json'))
print('value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'frac', round(d['roofline']['frac'],3), 'issued', round(d['roofline']['issued_frac'],3))
print('stages', {k:(round(v['ms'],3), round(v['achieved']), round(v['frac'],3)) for k,v in d['roofline']['stages'].items()})
print('pipeline', d['pipeline'])
print('check', d['check'])
print('cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']), 'peak_clocks', d['roofline']['peak_clocks'])
PY
( timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_c9_bench_ref.json 2>> gpurun_out/r02_c9_bench.err
python - <<'PY'
import json
a=json.load(open('gpurun_out/r02_c9_bench.json')); b=json.load(open('gpurun_out/r02_c9_bench_ref.json'))
print('same config:', a['config']==b['config'], 'ref value', round(b['value']))
PY

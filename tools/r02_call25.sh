#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -x -k "grouped_stages" ) > gpurun_out/r02_c27_pytest_a.log 2>&1
tail -15 gpurun_out/r02_c27_pytest_a.log
( FANDOM_SEARCH_TILE_GROUP=103 timeout 900 python -m pytest tests -m gpu -q -x -k "config_size or full_size or fp8_search or heterogeneous or device_records or ragged or multiple_scripts or golden" ) > gpurun_out/r02_c27_pytest_b.log 2>&1
tail -5 gpurun_out/r02_c27_pytest_b.log
for G in 103 39 103 39; do
  FANDOM_SEARCH_TILE_GROUP=$G timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-pipeline 2>> gpurun_out/r02_c27.err | sed "s/^{/{\"group\": $G, /" >> gpurun_out/r02_c27_bench.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_c27_bench.jsonl'):
    d=json.loads(l)
    print('group', d['group'], round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), round(d['roofline']['kernel_ms_per_launch'],2), d['details']['candidates_per_step'], d['clocks'])
PY
tail -3 gpurun_out/r02_c27.err

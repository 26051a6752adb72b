#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "grouped_stages or defaults or config_size or full_size or prefilter or fp8_search or heterogeneous or device_records" ) > gpurun_out/r02_c19_pytest.log 2>&1
tail -3 gpurun_out/r02_c19_pytest.log
for G in 39 7 39 7; do
  FANDOM_SEARCH_TILE_GROUP=$G timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-pipeline 2>> gpurun_out/r02_c19.err | sed "s/^{/{\"group\": $G, /" >> gpurun_out/r02_c19_bench.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_c19_bench.jsonl'):
    d=json.loads(l)
    print('group', d['group'], round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), round(d['roofline']['kernel_ms_per_launch'],2), d['details']['candidates_per_step'], d['clocks'])
PY

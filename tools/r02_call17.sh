#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_c17_pytest.log 2>&1
tail -3 gpurun_out/r02_c17_pytest.log
for W in 2 0 2 0; do
  FS_DEBUG_WAIT=$W timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-pipeline 2>> gpurun_out/r02_c17.err | sed "s/^{/{\"wait\": $W, /" >> gpurun_out/r02_c17_bench.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_c17_bench.jsonl'):
    d=json.loads(l)
    print('wait', d['wait'], round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), round(d['roofline']['kernel_ms_per_launch'],2), d['clocks'])
PY

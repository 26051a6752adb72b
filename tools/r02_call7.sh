#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_device_records.py -m gpu -q -x ) > gpurun_out/r02_c7_pytest_records.log 2>&1
tail -30 gpurun_out/r02_c7_pytest_records.log
( timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_device_records.py ) > gpurun_out/r02_c7_pytest.log 2>&1
tail -5 gpurun_out/r02_c7_pytest.log
timeout 600 python tools/pipeline_bench.py --works 20000 --script-tokens 25000 --repeat 4 2> gpurun_out/r02_c7_pipeline.err | grep '^{' > gpurun_out/r02_c7_pipeline.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r02_c7_pipeline.jsonl'):
    d=json.loads(l); print(round(d['total_s'],3), round(d['pipeline_windows_per_s']/1e6,1), {k[:14]:round(v,3) for k,v in d['stage_s'].items()})
PY
tail -3 gpurun_out/r02_c7_pipeline.err

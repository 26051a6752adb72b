#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "two_gpus or two_devices or device_records or reuse_histogram" ) > gpurun_out/r02_c11_pytest_2gpu.log 2>&1
tail -6 gpurun_out/r02_c11_pytest_2gpu.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r02_c11_bench_2gpu.json 2> gpurun_out/r02_c11_bench_2gpu.err
tail -4 gpurun_out/r02_c11_bench_2gpu.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_c11_bench_2gpu.json'):
    if not l.startswith('{'): continue
    d=json.loads(l)
    print('value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'n', d['n_gpus'])
    p=d['pipeline']; print('pipeline', round(p['value']/1e6,1), 'steady', round(p['steady_state']['value']/1e6,1), p['seconds'], p['rank0_phases_s'], p['host_threads_per_rank'])
    print('check', d['check'])
PY

#!/bin/bash
mkdir -p gpurun_out
nproc > gpurun_out/r02_c23_nproc.txt; df -h /tmp | tail -1 >> gpurun_out/r02_c23_nproc.txt; cat gpurun_out/r02_c23_nproc.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/r02_c23_bench_8gpu.json 2> gpurun_out/r02_c23_bench_8gpu.err
tail -4 gpurun_out/r02_c23_bench_8gpu.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_c23_bench_8gpu.json'):
    if not l.startswith('{'): continue
    d=json.loads(l)
    print('value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'n', d['n_gpus'], d['clocks'])
    p=d['pipeline']; print('pipeline', round(p['value']/1e6,1), 'steady', round(p['steady_state']['value']/1e6,1), p['works'], p['seconds'], p['rank0_phases_s'], p['host_threads_per_rank'], p['corpus_generation_s'])
    print('check', d['check'])
PY

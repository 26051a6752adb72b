#!/bin/bash
# C3: 1 M synthetic fanworks sharded over 8 B200, files -> CSV through search.analyze
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 8 --steps 5 --warmup 3 --pipeline-works 1000000 ) > gpurun_out/r02_c24_bench_8gpu_1m.json 2> gpurun_out/r02_c24_bench_8gpu_1m.err
tail -4 gpurun_out/r02_c24_bench_8gpu_1m.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_c24_bench_8gpu_1m.json'):
    if not l.startswith('{'): continue
    d=json.loads(l)
    print('value', round(d['value']/1e6,1), 'n', d['n_gpus'])
    p=d['pipeline']; print('pipeline', round(p['value']/1e6,1), 'steady', round(p['steady_state']['value']/1e6,1), p['works'], p['windows'], p['seconds'], p['csv_rows'], p['rank0_phases_s'], p['host_threads_per_rank'], p['corpus_generation_s'])
PY

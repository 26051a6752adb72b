#!/bin/bash
# round 2, GPU call 2: full GPU suite with the half2-packed bounds, bench at the defaults, the
# pre-filter column sweep (C2 bench workload; d = 768), files -> CSV pipeline with two clusters in flight
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r02_c2_pytest.log 2>&1
tail -22 gpurun_out/r02_c2_pytest.log
timeout 600 python bench.py > gpurun_out/r02_c2_bench_default.json 2> gpurun_out/r02_c2_bench_default.err
for K in 288 256 224 192; do
  FANDOM_SEARCH_PREFILTER_DIMS=$K timeout 300 python bench.py --steps 10 --no-cpu-baseline >> gpurun_out/r02_c2_bench_k.jsonl 2>> gpurun_out/r02_c2_bench_k.err
done
python - <<'PY'
import json
for f in ('gpurun_out/r02_c2_bench_default.json','gpurun_out/r02_c2_bench_k.jsonl'):
  for l in open(f):
    d=json.loads(l)
    print(d['config'].get('kept_dims'), round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), round(d['roofline']['kernel_ms_per_launch'],2), d['config']['candidates_per_step'], d['clocks'])
PY
for K in 0 640 512; do
  FANDOM_SEARCH_PREFILTER_DIMS=$K timeout 300 python tools/sweep.py --one 6 2500000 25000 768 --pair 2 --reps 5 >> gpurun_out/r02_c2_sweep_d768.jsonl 2>> gpurun_out/r02_c2_sweep.err
done
cut -c 1-330 gpurun_out/r02_c2_sweep_d768.jsonl
timeout 600 python tools/pipeline_bench.py --works 20000 --script-tokens 25000 --repeat 3 > gpurun_out/r02_c2_pipeline.jsonl 2> gpurun_out/r02_c2_pipeline.err
FANDOM_SEARCH_PREFILTER_DIMS=256 timeout 600 python tools/pipeline_bench.py --works 20000 --script-tokens 25000 --repeat 3 >> gpurun_out/r02_c2_pipeline.jsonl 2>> gpurun_out/r02_c2_pipeline.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_c2_pipeline.jsonl'):
    d=json.loads(l); print(round(d['total_s'],3), round(d['pipeline_windows_per_s']/1e6,1), {k[:14]:round(v,3) for k,v in d['stage_s'].items()})
PY
tail -3 gpurun_out/r02_c2_pipeline.err

#!/bin/bash
# round 2, GPU call 1: full GPU test suite (incl. config-size parity), bench at the defaults,
# pre-filter column sweep on the C2 bench workload and at d = 768
mkdir -p gpurun_out
nproc > gpurun_out/r02_c1_nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r02_c1_pytest.log 2>&1
tail -25 gpurun_out/r02_c1_pytest.log
timeout 600 python bench.py > gpurun_out/r02_c1_bench_default.json 2> gpurun_out/r02_c1_bench_default.err
tail -c 600 gpurun_out/r02_c1_bench_default.json
for K in 288 256 224 192; do
  FANDOM_SEARCH_PREFILTER_DIMS=$K timeout 300 python bench.py --steps 10 --no-cpu-baseline >> gpurun_out/r02_c1_bench_k.jsonl 2>> gpurun_out/r02_c1_bench_k.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_c1_bench_k.jsonl'):
    d=json.loads(l)
    print(d['config'].get('kept_dims'), round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), d['roofline']['kernel_ms_per_launch'], d['config']['candidates_per_step'], d['clocks'])
PY
for K in 0 640 512; do
  FANDOM_SEARCH_PREFILTER_DIMS=$K timeout 300 python tools/sweep.py --one 6 2500000 25000 768 --pair 2 --reps 5 >> gpurun_out/r02_c1_sweep_d768.jsonl 2>> gpurun_out/r02_c1_sweep.err
done
cat gpurun_out/r02_c1_sweep_d768.jsonl | cut -c 1-400

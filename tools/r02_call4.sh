#!/bin/bash
# round 2, GPU call 4: alternating epilogue warp sets (FS_OPT_TILE_GROUP bit 3) vs the 16-warp lockstep epilogue
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "grouped_stages or defaults or config_size or full_size or prefilter or fp8_search or resident" ) > gpurun_out/r02_c4_pytest.log 2>&1
tail -4 gpurun_out/r02_c4_pytest.log
for G in 7 15 7 15; do
  FANDOM_SEARCH_TILE_GROUP=$G timeout 300 python bench.py --steps 20 --no-cpu-baseline >> gpurun_out/r02_c4_bench.jsonl 2>> gpurun_out/r02_c4_bench.err
done
for K in 0 192; do
  FANDOM_SEARCH_PREFILTER_DIMS=$K timeout 300 python bench.py --steps 20 --no-cpu-baseline >> gpurun_out/r02_c4_bench.jsonl 2>> gpurun_out/r02_c4_bench.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_c4_bench.jsonl'):
    d=json.loads(l)
    print(d['config'].get('kept_dims'), round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), round(d['roofline']['kernel_ms_per_launch'],2), d['config']['candidates_per_step'], d['clocks'])
PY
for K in -1 0; do
  FANDOM_SEARCH_PREFILTER_DIMS=$K timeout 300 python tools/sweep.py --one 6 2500000 25000 768 --pair 2 --reps 5 >> gpurun_out/r02_c4_sweep_d768.jsonl 2>> gpurun_out/r02_c4_sweep.err
done
cut -c 1-330 gpurun_out/r02_c4_sweep_d768.jsonl

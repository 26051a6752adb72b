#!/bin/bash
# round 2, GPU call 6: epilogue with prefetched bounds (bit 4) +- alternating warp sets (bit 3)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "grouped_stages or defaults or config_size or full_size or prefilter or fp8_search or resident or ragged or multiple_scripts" ) > gpurun_out/r02_c6_pytest.log 2>&1
tail -4 gpurun_out/r02_c6_pytest.log
for G in 7 23 31 7 23 31; do
  FANDOM_SEARCH_TILE_GROUP=$G timeout 300 python bench.py --steps 20 --no-cpu-baseline >> gpurun_out/r02_c6_bench.jsonl 2>> gpurun_out/r02_c6_bench.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_c6_bench.jsonl'):
    d=json.loads(l)
    print(d['config'].get('kept_dims'), round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), round(d['roofline']['kernel_ms_per_launch'],2), d['config']['candidates_per_step'], d['clocks'])
PY
FS_NVCC_EXTRA=-DFS_TIMELINE python -m fandom_search_b200.build --force > gpurun_out/r02_c6_build.log 2>&1
python tools/timeline.py FS_OPT_TILE_GROUP=23 > gpurun_out/r02_c6_timeline_g23.txt 2> gpurun_out/r02_c6_tl.err
python tools/timeline.py FS_OPT_TILE_GROUP=31 > gpurun_out/r02_c6_timeline_g31.txt 2>> gpurun_out/r02_c6_tl.err
head -1 gpurun_out/r02_c6_timeline_g23.txt gpurun_out/r02_c6_timeline_g31.txt

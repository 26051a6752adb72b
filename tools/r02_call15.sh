#!/bin/bash
mkdir -p gpurun_out
FS_NVCC_EXTRA="-DFS_FLOOR_PROBE -DFS_TIMELINE" python -m fandom_search_b200.build --force > gpurun_out/r02_c15_build.log 2>&1
for G in 7 23 55 119; do
  python tools/timeline.py FS_OPT_TILE_GROUP=$G > gpurun_out/r02_c15_timeline_g$G.txt 2>> gpurun_out/r02_c15.err
  head -1 gpurun_out/r02_c15_timeline_g$G.txt
done
tail -2 gpurun_out/r02_c15.err

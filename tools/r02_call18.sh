#!/bin/bash
mkdir -p gpurun_out
FS_NVCC_EXTRA="-DFS_TIMELINE" python -m fandom_search_b200.build --force > gpurun_out/r02_c18_build.log 2>&1
python tools/timeline.py --c2 --hot > gpurun_out/r02_c18_timeline_c2.txt 2>> gpurun_out/r02_c18.err
python tools/timeline.py --hot > gpurun_out/r02_c18_timeline_rand.txt 2>> gpurun_out/r02_c18.err
head -1 gpurun_out/r02_c18_timeline_c2.txt gpurun_out/r02_c18_timeline_rand.txt
tail -2 gpurun_out/r02_c18.err

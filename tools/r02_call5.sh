#!/bin/bash
mkdir -p gpurun_out
python tools/timeline.py FS_OPT_TILE_GROUP=7 > gpurun_out/r02_c5_timeline_g7.txt 2> gpurun_out/r02_c5_tl.err
python tools/timeline.py FS_OPT_TILE_GROUP=15 > gpurun_out/r02_c5_timeline_g15.txt 2>> gpurun_out/r02_c5_tl.err
python tools/timeline.py FS_OPT_TILE_GROUP=7 FS_OPT_PREFILTER_DIMS=0 > gpurun_out/r02_c5_timeline_g7_k300.txt 2>> gpurun_out/r02_c5_tl.err
head -1 gpurun_out/r02_c5_timeline_g7.txt gpurun_out/r02_c5_timeline_g15.txt gpurun_out/r02_c5_timeline_g7_k300.txt
tail -3 gpurun_out/r02_c5_tl.err

#!/bin/bash
# round 2, GPU call 20: the other BASELINE configs through bench.py, C5 sweep with the default kernel
# (grid + 1e8/1e9 points), stage bench, ncu captures of the HBM-bound and post-processing kernels
mkdir -p gpurun_out
for C in C1 C4 D768; do
  timeout 600 python bench.py --config $C --steps 5 2>> gpurun_out/r02_c20.err >> gpurun_out/r02_c20_bench_configs.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_c20_bench_configs.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:14], d['config']['script_windows'], round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), round(d['roofline']['frac'],3), round(d['roofline']['issued_frac'],3), d['details']['kept_dims'])
PY
timeout 600 python tools/stage_bench.py > gpurun_out/r02_c20_stage_bench.json 2>> gpurun_out/r02_c20.err
cat gpurun_out/r02_c20_stage_bench.json | cut -c 1-600
timeout 900 python tools/sweep.py --defaults > gpurun_out/r02_c20_sweep_defaults.jsonl 2>> gpurun_out/r02_c20.err
timeout 1200 python tools/sweep.py --big > gpurun_out/r02_c20_sweep_big.jsonl 2>> gpurun_out/r02_c20.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_c20_sweep_defaults.jsonl','gpurun_out/r02_c20_sweep_big.jsonl'):
    for l in open(f):
        d=json.loads(l); print(d['dim'], d['kept_dims'], d['script_windows'], d['total_fan_windows'], round(d['kernel_ms'],2), round(d['windows_per_s']/1e6,1))
PY
for K in gather_kernel hash_probe_kernel rescore_kernel post_rank_lev_kernel window_norm_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/r02_c20_prof_$K -f python tools/pipeline_bench.py --works 1500 --script-tokens 25000 > gpurun_out/r02_c20_ncu_$K.log 2>&1
  ncu -i gpurun_out/r02_c20_prof_$K.ncu-rep --page raw --csv > gpurun_out/r02_c20_ncu_raw_$K.csv 2>/dev/null
done
ls -la gpurun_out/r02_c20_prof_*.ncu-rep
tail -3 gpurun_out/r02_c20.err

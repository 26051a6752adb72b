#!/bin/bash
# round 2, GPU call 3: defaults with automatic pre-filter columns; ncu capture of the distance kernel
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_c3_pytest.log 2>&1
tail -4 gpurun_out/r02_c3_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_c3_bench.json 2> gpurun_out/r02_c3_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_c3_bench.json'))
print(d['config'].get('kept_dims'), round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), round(d['roofline']['kernel_ms_per_launch'],2), d['config']['candidates_per_step'], d['clocks'])
PY
python - <<'PY' > gpurun_out/r02_c3_index_create.txt 2>&1
import time, numpy as np, sys
sys.path.insert(0, '.')
from fandom_search_b200 import synth
from fandom_search_b200.engine import DeviceIndex
lex = synth.SynthLexicon(vocab=50000, dim=300, oov_frac=0.0, seed=1001)
script = synth.make_script_tokens(lex, 25000).astype(np.int32)
for k in range(4):
    t=time.perf_counter(); idx = DeviceIndex(lex.table_all, script); dt=time.perf_counter()-t
    print("index create %.1f ms kept %d" % (dt*1e3, idx.kept_dims)); idx.close()
PY
cat gpurun_out/r02_c3_index_create.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c3_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --distinct 2 > gpurun_out/r02_c3_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:distance_kernel -s 3 -c 1 -o gpurun_out/r02_c3_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --distinct 2 > gpurun_out/r02_c3_ncu_b.log 2>&1
ncu -i gpurun_out/r02_c3_prof.ncu-rep --page raw --csv > gpurun_out/r02_c3_distance_kernel_ncu_raw.csv 2>/dev/null
ls -la gpurun_out/r02_c3_prof.ncu-rep
timeout 600 python tools/pipeline_bench.py --works 20000 --script-tokens 25000 --repeat 4 2> gpurun_out/r02_c3_pipeline.err | grep '^{' > gpurun_out/r02_c3_pipeline.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r02_c3_pipeline.jsonl'):
    d=json.loads(l); print(round(d['total_s'],3), round(d['pipeline_windows_per_s']/1e6,1), {k[:14]:round(v,3) for k,v in d['stage_s'].items()})
PY

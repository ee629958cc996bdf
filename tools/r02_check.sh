#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_large.py tests/test_gpu_strips.py tests/test_gpu_cluster.py -x -q -m gpu > gpurun_out/chk_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/chk_tests.log
tail -4 gpurun_out/chk_tests.log
timeout 900 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/chk_bench.json 2> gpurun_out/chk_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/chk_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d.get('stage_ms_per_step'), d['roofline']['frac'])
for k in ('config1','config3','config3_shallow','config4'):
    if k in d: print(k, d[k].get('ms_per_step'), d[k].get('ms_per_step_median'), d[k].get('stage_ms_per_step'), d[k].get('roofline',{}).get('frac'))
PY
FEROX_SOLVE_TRACE=1 timeout 300 python tools/solve_trace.py config2 400 2>&1 | tail -9 | head -8
FEROX_SOLVE_TRACE=1 timeout 400 python tools/solve_trace.py config3 300 2>&1 | tail -19 | head -8

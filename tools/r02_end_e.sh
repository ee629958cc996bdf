#!/bin/bash
# dev: last call of the round = full GPU test suite + the default bench line
mkdir -p gpurun_out
(timeout 230 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log)
tail -3 gpurun_out/r02g_pytest.log | cut -c1-300
timeout 100 python bench.py --skip-sub --no-cpu-baseline > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
timeout 100 python bench.py --workload config3 --steps 100 --warmup 10 --skip-sub --no-cpu-baseline > gpurun_out/r02g_c3.json 2>/dev/null
python - <<'PY'
import json
for f in ('r02g_bench','r02g_c3'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["stage_ms_per_step"])
    except Exception as e: print(f, 'failed', e)
PY

#!/bin/bash
# dev: quick check after restoring the round-1 pair emission: candidate-sequence / golden / 262k parity tests + the 1 M-circle bed
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -q -p no:cacheprovider -k "sequence or golden or 262k" 2>&1 | tail -2
timeout 200 python bench.py --workload config3 --steps 100 --warmup 10 --skip-sub --no-cpu-baseline 2>/dev/null > gpurun_out/r02e_c3.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_c3.json').read().strip().splitlines()[-1]); print(d["ms_per_step"], d["stage_ms_per_step"], d["roofline"]["frac"])
PY

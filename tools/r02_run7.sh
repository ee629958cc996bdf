#!/bin/bash
# dev: one GPU call = full GPU test suite, cluster-solver size sweep, default bench, ncu capture of config3 (shallow bed)
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest7.log)
tail -4 gpurun_out/r02_pytest7.log | cut -c1-300
for n in 6000 12000 24000 40000; do for c in 0 16; do
FEROX_CLUSTER=$c python bench.py --bodies $n --steps 200 --warmup 5 --settle 400 --skip-sub --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N', sys.argv[1], 'CLUSTER', sys.argv[2], 'ms', round(d['ms_per_step'],4), 'solve', round(d['stage_ms_per_step']['solve'],4), d['config']['counters']['manifolds'], d['config']['counters']['colors'])
" $n $c
done; done
timeout 900 python bench.py > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench7.err
python bench.py --workload config3 --steps 2 --warmup 3 --skip-sub --no-cpu-baseline --settle 300 > gpurun_out/plain3.log 2>&1 && \
FEROX_NCU_RANGE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_solve_colored|k_narrowphase_tiled|k_sort_pass|k_emit_pairs|k_body_prepass|k_emit_entries" -c 16 -o gpurun_out/r02_config3 python bench.py --workload config3 --steps 2 --warmup 3 --skip-sub --no-cpu-baseline --settle 300 > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log | cut -c1-200; ls -la gpurun_out/r02_config3.ncu-rep

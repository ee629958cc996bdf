#!/bin/bash
# dev: end-of-round GPU test suite + ncu capture of the 1 M-circle bed (config3_shallow): broadphase kernels, narrow phase, cache, solver
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log)
tail -4 gpurun_out/r02e_pytest.log | cut -c1-300
CMD="python bench.py --workload config3 --steps 2 --warmup 3 --skip-sub --no-cpu-baseline --settle 300"
$CMD > gpurun_out/plain3.log 2>&1 && \
FEROX_NCU_RANGE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_solve_colored|k_narrowphase_tiled|k_sort_pass|k_emit_pairs|k_body_prepass|k_emit_entries|k_carry_over|k_hash_build" -c 22 -o gpurun_out/r02e_config3 $CMD > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log | cut -c1-200; ls -la gpurun_out/r02e_config3.ncu-rep

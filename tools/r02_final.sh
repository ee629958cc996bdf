#!/bin/bash
# dev: final single-GPU validation = full GPU test suite, smoke, default bench line, launch list + full capture of config 2
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_final.log)
tail -4 gpurun_out/r02_pytest_final.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_final.err
python bench.py --steps 3 --warmup 3 --skip-sub --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
FEROX_NCU_RANGE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -c 39 -o gpurun_out/r02_config2_full python bench.py --steps 3 --warmup 3 --skip-sub --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu2.log | cut -c1-200; ls -la gpurun_out/r02_config2_full.ncu-rep

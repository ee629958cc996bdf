#!/bin/bash
# dev: end-of-round single-GPU evidence = smoke, default bench line, full ncu capture of three config-2 steps
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02e_bench_n1.err
python bench.py --steps 3 --warmup 3 --skip-sub --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
FEROX_NCU_RANGE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -c 39 -o gpurun_out/r02e_config2_full python bench.py --steps 3 --warmup 3 --skip-sub --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu2.log | cut -c1-200; ls -la gpurun_out/r02e_config2_full.ncu-rep

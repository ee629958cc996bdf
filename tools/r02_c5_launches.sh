#!/bin/bash
# launch list of ONE strip step at 1 M bodies on one GPU: the solver's stages are separate launches there
# (set-up + warm start, ten sweeps, write-back), which splits the 1.7 ms the fused kernel spends
mkdir -p gpurun_out
CMD="python bench.py --workload config5 --steps 2 --warmup 3 --skip-sub --no-cpu-baseline --settle 300"
timeout 600 $CMD > gpurun_out/c5_plain.json 2> gpurun_out/c5_plain.err || exit 1
FEROX_NCU_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02_launches_config5_n1.csv $CMD > gpurun_out/c5_ncu.log 2>&1
tail -2 gpurun_out/c5_ncu.log

#!/bin/bash
# dev: end-of-round 2-GPU check = the NCCL / peer-direct strip tests + a short 2-rank bench line (config 2 replicas, config 4 sharded, config 5 strips)
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_strips.py -m gpu -q -p no:cacheprovider -k "nccl" > gpurun_out/r02e_pytest_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest_n2.log)
tail -3 gpurun_out/r02e_pytest_n2.log | cut -c1-300
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --sub-steps 200 > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02e_bench_n2.err

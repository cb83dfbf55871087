#!/bin/bash
# round 2, visit c: whole GPU suite, new bench (CUPTI per-class timing), chain on/off, configs 3/4/5
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r02c_gpu_tests.log 2>&1; echo "tests exit=$?"; tail -8 gpurun_out/r02c_gpu_tests.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench.log 2>&1; echo "bench exit=$?"; tail -c 3500 gpurun_out/r02c_bench.log
ORVB_FF_CHAIN=1 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_chain.log 2>&1; echo "bench chain exit=$?"; tail -c 3500 gpurun_out/r02c_bench_chain.log
for C in 3 4 5; do timeout 600 python bench.py --config $C --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_cfg$C.log 2>&1; echo "config $C exit=$?"; tail -c 1200 gpurun_out/r02c_bench_cfg$C.log | cut -c1-1200; done

#!/bin/bash
# round 2, visit b: full GPU suite with the v101 ABI (tight tests, new pipeline goldens), chain on/off in the forward
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/r02b_gpu_tests.log 2>&1; echo "tests exit=$?"; tail -25 gpurun_out/r02b_gpu_tests.log
ORVB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zz_gpu_experimental.py -m gpu -q -x > gpurun_out/r02b_experimental.log 2>&1; echo "experimental exit=$?"; tail -5 gpurun_out/r02b_experimental.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench.log 2>&1; echo "bench exit=$?"; tail -c 2500 gpurun_out/r02b_bench.log
ORVB_FF_CHAIN=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_chain.log 2>&1; echo "bench chain exit=$?"; tail -c 2500 gpurun_out/r02b_bench_chain.log

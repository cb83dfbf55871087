#!/bin/bash
# round 2, visit a: first hardware run of the chained FF kernel, GEMM timeline, clips-per-call sweep
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
ORVB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zz_gpu_experimental.py -m gpu -q -x > gpurun_out/r02a_experimental.log 2>&1; echo "experimental exit=$?"; tail -15 gpurun_out/r02a_experimental.log
ORVB_LIB_PATH=orv_b200/liborv_b200_tl.so timeout 200 python tools/profile_gemm_timeline.py > gpurun_out/r02a_timeline.log 2>&1; echo "timeline exit=$?"; cat gpurun_out/r02a_timeline.log
for B in 2 4; do timeout 200 python tools/bench_batch.py $B 3 > gpurun_out/r02a_batch$B.log 2>&1; echo "batch $B exit=$?"; tail -c 500 gpurun_out/r02a_batch$B.log; done
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench.log 2>&1; echo "bench exit=$?"; tail -c 3000 gpurun_out/r02a_bench.log

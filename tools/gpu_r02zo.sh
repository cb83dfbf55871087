#!/bin/bash
# fast gated-residual epilogue: parity, then timeline A/B (fast vs generic, in place)
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_fast_resid.py tests/test_gpu_ops.py tests/test_gpu_tight.py -q -x --timeout 300 > gpurun_out/r02zo_tests.log 2>&1; echo "tests exit=$?"; tail -15 gpurun_out/r02zo_tests.log
ORVB_LIB_PATH=orv_b200/liborv_b200_tl.so timeout 300 python tools/profile_gemm_timeline.py > gpurun_out/r02zo_timeline_fast.log 2>&1; echo "timeline fast exit=$?"
ORVB_GEMM_FAST_RESID=0 ORVB_LIB_PATH=orv_b200/liborv_b200_tl.so timeout 300 python tools/profile_gemm_timeline.py > gpurun_out/r02zo_timeline_generic.log 2>&1; echo "timeline generic exit=$?"
grep -A4 "IN PLACE" gpurun_out/r02zo_timeline_fast.log gpurun_out/r02zo_timeline_generic.log | cut -c1-420

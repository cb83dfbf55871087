#!/bin/bash
# compute-sanitizer memcheck over the GEMM paths of the third session of round 2: prefetching gated-residual epilogue,
# mixed tile list, k_wrap (small K keeps the ~50x slowdown bearable)
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gemm_fast_resid.py tests/test_gpu_gemm_mixed_tiles.py tests/test_gpu_tight.py -q -x --timeout 900 \
  -k "(bit_identical_to_generic and not 1920-1920) or optional_operands or single_cta or remainder_tile or (qkv_mixed and 256) or k_wrap" > gpurun_out/r02zzm_memcheck_gemm.log 2>&1
echo "gemm memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02zzm_memcheck_gemm.log | tail -3

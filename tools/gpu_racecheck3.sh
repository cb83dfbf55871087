#!/bin/bash
# compute-sanitizer racecheck / synccheck over the prefetching gated-residual epilogue (gate / bias rows staged in shared
# memory by the warp, residual tile read-modify-written in place) and the mixed tile list
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
for tool in racecheck synccheck; do
timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_gemm_fast_resid.py tests/test_gpu_gemm_mixed_tiles.py -q --timeout 600 \
  -k "(bit_identical_to_generic and not 1920-1920) or optional_operands or remainder_tile" > gpurun_out/r02zzm_${tool}_gemm.log 2>&1
echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r02zzm_${tool}_gemm.log | sort | uniq -c | tail -6
done

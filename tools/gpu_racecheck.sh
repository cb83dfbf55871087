#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) and synccheck over the kernels of round 2 that exchange data through
# shared memory: GroupNorm statistics, the convolution's per-CTA statistics fold, the look-back scans, the label vote
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
for tool in racecheck synccheck; do
timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_vae.py tests/test_zz_gpu_voxelize.py -q -x --timeout 900 \
  -k "gn_stats or fused_groupnorm or spatial_norm or untiled_small or golden or label or small" > gpurun_out/r02zl_$tool.log 2>&1
echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r02zl_$tool.log | tail -4
done

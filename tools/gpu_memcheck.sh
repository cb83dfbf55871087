#!/bin/bash
# compute-sanitizer memcheck over the kernels written in round 2 (small cases only: the tool slows kernels ~50x)
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_vae.py -q -x --timeout 900 \
  -k "conv_cl_f32 and not 60 or gn_stats or spatial_norm or upsample or cl_to_planar or untiled_small or tiled_small" > gpurun_out/r02zc_memcheck_vae.log 2>&1
echo "vae memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02zc_memcheck_vae.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tight.py -q -x --timeout 600 -k "attention" > gpurun_out/r02zc_memcheck_attn.log 2>&1
echo "attention memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02zc_memcheck_attn.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_voxelize.py -q -x --timeout 600 -k "not 5m and not large" > gpurun_out/r02zc_memcheck_voxel.log 2>&1
echo "voxel memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02zc_memcheck_voxel.log | tail -3

#!/bin/bash
# round 2, VAE decode bring-up: operator + end-to-end parity, then the full-size timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vae.py -x -q --timeout 180 > gpurun_out/r02l_vae_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02l_vae_tests.log
tail -25 gpurun_out/r02l_vae_tests.log
true
echo "bench rc=$?"
tail -5 gpurun_out/r02l_vae_bench.err
cat gpurun_out/r02l_vae_bench.json

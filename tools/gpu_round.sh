#!/bin/bash
# One GPU-box visit: tests, bench, ncu launch list + full captures.  Usage: tools/gpu_round.sh <tag>
TAG=${1:-r01}
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/${TAG}_gpu_tests.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench.log 2>&1; echo "bench exit=$?"; tail -c 2500 gpurun_out/${TAG}_bench.log
KREG='regex:gemm_bf16|attention_kernel|ln_modulate|skinny|patchify|ab_combine|sampler_step|build_emb|timestep_sin|add_hidden|actions_to'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 230 -c 231 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_forward.py 2 > gpurun_out/${TAG}_ncu1.log 2>&1; echo "ncu launches exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 31 -c 1 -f -o gpurun_out/${TAG}_attn python tools/profile_forward.py 2 > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu attn exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 126 -c 4 -f -o gpurun_out/${TAG}_gemm python tools/profile_forward.py 2 > gpurun_out/${TAG}_ncu3.log 2>&1; echo "ncu gemm exit=$?"
ls -la gpurun_out | tail -12

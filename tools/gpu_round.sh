#!/bin/bash
# One GPU-box visit: tests, bench, ncu launch list + per-kernel captures.  Usage: tools/gpu_round.sh <tag>
# Condense afterwards (here, on the CPU box) with: python tools/summarize_profiles.py <tag>
TAG=${1:-r02}
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/${TAG}_gpu_tests.log
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1; echo "bench exit=$?"; tail -c 1500 gpurun_out/${TAG}_bench.log
KREG='regex:gemm|attention_kernel|ln_|skinny|patchify|ab_combine|sampler_step|build_emb|timestep_sin|add_hidden|actions_to|fill_tables'
# every launch of one forward with its device time (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 231 -c 232 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_forward.py 2 > gpurun_out/${TAG}_ncu1.log 2>&1; echo "ncu launches exit=$?"
# Full-set captures patch the kernel for the SourceCounters section, which fails to launch for kernels that already
# use the whole 227 KB of shared memory (attention, CTA-pair GEMM): capture every section except that one.
SECS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section InstructionStats"
timeout 600 ncu $SECS --clock-control none -k regex:attention_kernel -s 31 -c 1 -f -o gpurun_out/${TAG}_attn python tools/profile_forward.py 2 > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu attn exit=$?"
timeout 600 ncu $SECS --clock-control none -k regex:gemm2_bf16 -s 126 -c 4 -f -o gpurun_out/${TAG}_gemm python tools/profile_forward.py 2 > gpurun_out/${TAG}_ncu3.log 2>&1; echo "ncu gemm exit=$?"
timeout 600 ncu $SECS --clock-control none -k regex:ln_ab_kernel\|skinny_linear -s 61 -c 3 -f -o gpurun_out/${TAG}_pointwise python tools/profile_forward.py 2 > gpurun_out/${TAG}_ncu4.log 2>&1; echo "ncu pointwise exit=$?"
ls -la gpurun_out | tail -14
# VAE decode: timing + per-kernel device times + one capture of the 128-channel convolution
timeout 600 python tools/bench_vae.py --steps 5 --warmup 2 --profile > gpurun_out/${TAG}_vae_bench.json 2> gpurun_out/${TAG}_vae_bench.err; echo "vae bench exit=$?"; cut -c1-400 gpurun_out/${TAG}_vae_bench.json
timeout 600 ncu $SECS --clock-control none -k regex:conv2_bf16 -s 330 -c 2 -f -o gpurun_out/${TAG}_conv python tools/bench_vae.py --steps 1 --warmup 0 --skip-eager > gpurun_out/${TAG}_ncu6.log 2>&1; echo "ncu conv exit=$?"
# attention kernel alone on the shapes of every config, next to torch SDPA
timeout 300 python tools/bench_attention.py > gpurun_out/${TAG}_bench_attention.log 2>&1; echo "bench attention exit=$?"; cat gpurun_out/${TAG}_bench_attention.log
# voxelization row: bench line + launch list with DRAM bytes (condense with tools/summarize_voxel_profile.py <tag>)
cp gpurun_out/${TAG}_gpu_tests.log gpurun_out/${TAG}_voxel_gpu_tests.log
timeout 120 python tools/bench_voxelize.py 2000000 > gpurun_out/${TAG}_voxel_bench.json 2> gpurun_out/${TAG}_voxel_bench.err; echo "voxel bench exit=$?"
timeout 200 python tools/bench_voxelize.py 20000000 > gpurun_out/${TAG}_voxel_bench_20m.json 2> gpurun_out/${TAG}_voxel_bench_20m.err; echo "voxel bench 20M exit=$?"
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${TAG}_voxel_launches.csv -k "regex:voxel|radix|scan|segment|gather|label|sort_key" python tools/profile_voxelize.py > gpurun_out/${TAG}_ncu5.log 2>&1; echo "ncu voxel exit=$?"
# clips per pipeline call (the reference's evaluation batch is 4-16, config/base_eval.yaml:119)
for B in 2 4; do timeout 300 python bench.py --clips-per-gpu $B --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_b$B.log 2>&1; echo "B=$B exit=$?"; tail -c 300 gpurun_out/${TAG}_bench_b$B.log; done
# the other BASELINE.json configs through the pipeline
for C in 3 4 5; do timeout 600 python bench.py --config $C --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg$C.log 2>&1; echo "config $C exit=$?"; tail -c 300 gpurun_out/${TAG}_bench_cfg$C.log; done

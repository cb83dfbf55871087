#!/bin/bash
# round 2, visit d: iteration-boundary timelines (config 2, config 5), configs 4 / 5 through the pipeline
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 300 python tools/profile_step_timeline.py 2 > gpurun_out/r02d_timeline_cfg2.log 2>&1; echo "timeline2 exit=$?"; grep -v Warning gpurun_out/r02d_timeline_cfg2.log | tail -40
timeout 400 python tools/profile_step_timeline.py 5 > gpurun_out/r02d_timeline_cfg5.log 2>&1; echo "timeline5 exit=$?"; grep -v Warning gpurun_out/r02d_timeline_cfg5.log | tail -40
for C in 4 5; do timeout 600 python bench.py --config $C --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_cfg$C.log 2>&1; echo "config $C exit=$?"; tail -c 1500 gpurun_out/r02d_bench_cfg$C.log | cut -c1-1500; done

#!/bin/bash
# round 2, visit i: Gaussian rasteriser vs the reference extension; per-clip overhead breakdown
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_gs_render.py -m gpu -q -x -s > gpurun_out/r02i_gs_tests.log 2>&1; echo "gs tests exit=$?"; tail -30 gpurun_out/r02i_gs_tests.log
timeout 120 python tools/make_gs_golden.py; echo "golden exit=$?"
timeout 300 python tools/bench_gs_render.py > gpurun_out/r02i_gs_bench.json 2> gpurun_out/r02i_gs_bench.err; echo "gs bench exit=$?"; cat gpurun_out/r02i_gs_bench.json; tail -3 gpurun_out/r02i_gs_bench.err
timeout 300 python tools/profile_step_timeline.py 2 > gpurun_out/r02i_timeline_cfg2.log 2>&1; echo "timeline exit=$?"; grep -v Warning gpurun_out/r02i_timeline_cfg2.log | head -40

#!/bin/bash
# noise staging ring: 8 slots (default) vs 2 (as before), per-clip times of config 5; pipeline goldens first
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py -q --timeout 300 2>&1 | tail -2
echo "== 8 slots"; timeout 300 python tools/time_clips.py 5 10 2>&1 | tail -2
echo "== 2 slots"; ORVB_NOISE_SLOTS=2 timeout 300 python tools/time_clips.py 5 10 2>&1 | tail -2
echo "== 8 slots, config 2"; timeout 300 python tools/time_clips.py 2 12 2>&1 | tail -2

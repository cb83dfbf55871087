#!/bin/bash
# tensor-core AdaLN table build: full GPU suite, then clip-boundary timing A/B (ORVB_MOD_TABLES_TC=0 vs 1)
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02zq_gpu_tests.log 2>&1; echo "tests exit=$?"; tail -6 gpurun_out/r02zq_gpu_tests.log
ORVB_MOD_TABLES_TC=0 timeout 300 python tools/profile_step_timeline.py 2 2>&1 | grep -E "^clip|clip prologue" > gpurun_out/r02zq_prologue_cuda_cores.log; cat gpurun_out/r02zq_prologue_cuda_cores.log
timeout 300 python tools/profile_step_timeline.py 2 2>&1 | grep -E "^clip|clip prologue" > gpurun_out/r02zq_prologue_tensor_cores.log; cat gpurun_out/r02zq_prologue_tensor_cores.log
ORVB_MOD_TABLES_TC=0 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cuda-core tables', round(d['value'],3), round(d['ms_per_step'],1), d['tensor_frac_of_peak'], d['clocks']['sm_mhz'])"
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tensor-core tables', round(d['value'],3), round(d['ms_per_step'],1), d['tensor_frac_of_peak'], d['clocks']['sm_mhz'])"

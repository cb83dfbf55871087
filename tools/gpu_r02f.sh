#!/bin/bash
# round 2, visit f: schedule read in place + noise upload on a side stream: tests, timeline, bench (configs 2 and 4)
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02f_gpu_tests.log 2>&1; echo "tests exit=$?"; tail -4 gpurun_out/r02f_gpu_tests.log
timeout 300 python tools/profile_step_timeline.py 2 > gpurun_out/r02f_timeline_cfg2.log 2>&1; echo "timeline2 exit=$?"; grep -v Warning gpurun_out/r02f_timeline_cfg2.log | tail -34
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench.log 2>&1; echo "bench exit=$?"; tail -c 600 gpurun_out/r02f_bench.log; python - <<'PY'
import json
for ln in open('gpurun_out/r02f_bench.log'):
    if ln.startswith('{'):
        d=json.loads(ln); print('value %.2f e2e %.2f ms/step %.1f frac %.4f fwd_sum %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['tensor_frac_of_peak'],d['kernel_timing']['forward_ms_sum_of_kernels']))
PY
timeout 600 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench_cfg4.log 2>&1; echo "config 4 exit=$?"; python - <<'PY'
import json
for ln in open('gpurun_out/r02f_bench_cfg4.log'):
    if ln.startswith('{'):
        d=json.loads(ln); print('cfg4 value %.2f e2e %.2f ms/step %.1f frac %.4f fwd_sum %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['tensor_frac_of_peak'],d['kernel_timing']['forward_ms_sum_of_kernels']))
PY

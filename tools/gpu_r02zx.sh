#!/bin/bash
# mixed tile list (QKV 22 x 256 + 128): parity, then config-2 bench A/B on one box
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_mixed_tiles.py tests/test_gpu_gemm_fast_resid.py tests/test_gpu_ops.py tests/test_gpu_tight.py tests/test_gpu_forward.py tests/test_gpu_golden.py -q -x --timeout 300 > gpurun_out/r02zx_tests.log 2>&1; echo "tests exit=$?"; tail -5 gpurun_out/r02zx_tests.log
for w in uniform mixed; do
  if [ $w = uniform ]; then export ORVB_GEMM_MIXED_TILES=0; else unset ORVB_GEMM_MIXED_TILES; fi
  timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02zx_bench_$w.json 2> gpurun_out/r02zx_bench_$w.err; echo "bench $w exit=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02zx_bench_$w.json").read().strip().splitlines()[-1])
k=d.get("kernels") or {}
print("$w", round(d["value"],3), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],3), d["tensor_frac_of_peak"], d["clocks"]["sm_mhz"], {n:(round(v.get("us_per_launch",0),1)) for n,v in k.items()})
PY
done

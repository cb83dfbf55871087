#!/bin/bash
# full GPU suite + config-2 bench A/B of the fast gated-residual epilogue (same box)
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02zp_gpu_tests.log 2>&1; echo "tests exit=$?"; tail -4 gpurun_out/r02zp_gpu_tests.log
ORVB_GEMM_FAST_RESID=0 timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02zp_bench_generic.json 2> gpurun_out/r02zp_bench_generic.err; echo "bench generic exit=$?"
timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02zp_bench_fast.json 2> gpurun_out/r02zp_bench_fast.err; echo "bench fast exit=$?"
for w in generic fast; do python - <<PY
import json
d=json.loads(open("gpurun_out/r02zp_bench_$w.json").read().strip().splitlines()[-1])
k=d.get("kernels") or {}
print("$w", round(d["value"],3), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],3), d["tensor_frac_of_peak"], d["clocks"]["sm_mhz"], {n:(round(v.get("us_per_launch",0),1)) for n,v in k.items()})
PY
done

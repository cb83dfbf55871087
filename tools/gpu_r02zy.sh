#!/bin/bash
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
for B in 2 4; do
  timeout 400 python bench.py --clips-per-gpu $B --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02zy_bench_b$B.json 2> gpurun_out/r02zy_bench_b$B.err; echo "B=$B exit=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02zy_bench_b$B.json").read().strip().splitlines()[-1])
k=d.get("kernels") or {}
print("B=$B", round(d["value"],3), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],3), d["tensor_frac_of_peak"], d["clocks"]["sm_mhz"], {n:(round(v.get("us_per_launch",0),1), round(v.get("frac_of_peak") or 0,3)) for n,v in k.items() if n.startswith("gemm") or n=="attention"})
PY
done

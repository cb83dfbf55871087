#!/bin/bash
# in-step A/B of the attention issue order (natural-order variant library vs the product library), same box, alternating
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
for rep in 1 2; do
for w in nat prod; do
  if [ $w = nat ]; then export ORVB_LIB_PATH=orv_b200/liborv_b200_nat.so; else unset ORVB_LIB_PATH; fi
  timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r03c_bench_${w}_$rep.json 2> gpurun_out/r03c_bench_${w}_$rep.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r03c_bench_${w}_$rep.json").read().strip().splitlines()[-1])
k=d.get("kernels") or {}
print("$w $rep", round(d["value"],3), round(d["ms_per_step"],1), d["tensor_frac_of_peak"], d["clocks"]["sm_mhz"], "attention", round(k["attention"]["us_per_launch"],1))
PY
done
done

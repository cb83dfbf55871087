#!/bin/bash
# config 4 (D = 3072): attn-out / FF2 at 240 columns with the generic epilogue vs 192 with the prefetching one
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
for w in 1.03 1.20; do
  ORVB_GEMM_FAST_PREF=$w timeout 500 python bench.py --config 4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02zz_cfg4_pref$w.json 2> gpurun_out/r02zz_cfg4_pref$w.err; echo "pref $w exit=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02zz_cfg4_pref$w.json").read().strip().splitlines()[-1])
k=d.get("kernels") or {}
print("pref $w", round(d["value"],3), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],3), d["tensor_frac_of_peak"], d["clocks"]["sm_mhz"], {n:(round(v.get("us_per_launch",0),1), round(v.get("frac_of_peak") or 0,3)) for n,v in k.items() if n.startswith("gemm") or n=="attention"})
PY
done

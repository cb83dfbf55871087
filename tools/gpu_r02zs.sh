#!/bin/bash
# config 5: why did the device-resident leg read 3999 ms (e2e leg 3375 ms) in r02zr?  A/B of this session's host-side changes
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 500 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02zs_cfg5_$tag.json 2> gpurun_out/r02zs_cfg5_$tag.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r02zs_cfg5_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"],3), "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],3), round(d["e2e"]["ms_per_step"],1), "frac", d["tensor_frac_of_peak"], "clk", d["clocks"]["sm_mhz"])
PY
}
run default X=1
run cuda_core_tables ORVB_MOD_TABLES_TC=0
run pageable ORVB_PINNED_UPLOADS=0

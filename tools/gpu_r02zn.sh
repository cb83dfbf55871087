#!/bin/bash
# A/B on one box: host prologue of clip i+1 overlapping the tail of clip i (pinned uploads, asynchronous read-back)
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
for C in 2 4; do
  ORVB_PINNED_UPLOADS=0 ORVB_BENCH_BLOCKING_READBACK=1 timeout 400 python bench.py --config $C --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02zn_cfg${C}_old.json 2> gpurun_out/r02zn_cfg${C}_old.err; echo "old cfg $C exit=$?"
  timeout 400 python bench.py --config $C --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02zn_cfg${C}_new.json 2> gpurun_out/r02zn_cfg${C}_new.err; echo "new cfg $C exit=$?"
  for w in old new; do python - <<PY
import json
d=json.loads(open("gpurun_out/r02zn_cfg${C}_$w.json").read().strip().splitlines()[-1])
print("cfg${C} $w", round(d["value"],3), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],3), round(d["e2e"]["ms_per_step"],1), d["tensor_frac_of_peak"], d["clocks"]["sm_mhz"])
PY
  done
done

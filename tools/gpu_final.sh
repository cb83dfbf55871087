#!/bin/bash
# what the driver runs at round end: GPU tests, smoke(), the default bench line
TAG=${1:-final}
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/${TAG}_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
k=d.get("kernels") or {}
print(round(d["value"],3), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],3), d["tensor_frac_of_peak"], d["clocks"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "cpu", round(d["cpu_baseline"]["value"],4), "eager", round(d["torch_eager_gpu"]["value"],2))
print({n:(round(v.get("us_per_launch",0),1), round(v.get("frac_of_peak") or 0,3)) for n,v in k.items()})
PY

#!/bin/bash
# Short GPU-box visit: the whole -m gpu suite and the default bench line.  Usage: tools/gpu_quick.sh <tag>
TAG=${1:-r02}
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "tests exit=$?"; tail -5 gpurun_out/${TAG}_gpu_tests.log
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err; echo "bench exit=$?"; tail -c 600 gpurun_out/${TAG}_bench.err; python - <<PY
import json
l = [x for x in open("gpurun_out/${TAG}_bench.log") if x.startswith("{")]
d = json.loads(l[-1])
print({k: d[k] for k in ("value", "ms_per_step", "tensor_frac_of_peak", "gpu_launches")})
print("e2e", d["e2e"]); print("decode", d.get("decode")); print("roofline", d["roofline"])
print({k: (v["us_per_launch"], v.get("frac_of_peak")) for k, v in (d.get("kernels") or {}).items()})
print("eager", d.get("torch_eager_gpu")); print("cpu", d.get("cpu_baseline")); print("clocks", d["clocks"])
PY

#!/bin/bash
# round 2, final 8-GPU visit: BASELINE.json configs 4 and 5 (quoted on 8 x B200) with the final build
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
for C in 4 5; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --config $C --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02zm_bench_cfg${C}_n8.log 2>&1
  echo "config $C n8 exit=$?"; grep '^{' gpurun_out/r02zm_bench_cfg${C}_n8.log | python -c "
import json,sys
for ln in sys.stdin:
    d=json.loads(ln); print('cfg', d['config']['workload'][:60], 'n', d['n_gpus'], 'value %.2f e2e %.2f ms/step %.1f frac %.4f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['tensor_frac_of_peak']), d['clocks'], (d.get('decode') or {}).get('decode_ms_per_step'))
"
done

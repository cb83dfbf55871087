#!/bin/bash
# attention: issuer program order (natural: QK0 QK1 PV0 PV1 per step; new: QK0(j+1) PV1(j-1) QK1(j+1) PV0(j)) x stagger
export ORVB_NO_BUILD=1
mkdir -p gpurun_out
echo "== natural order (-DORVB_ATT_NATURAL_ORDER)"; ORVB_SWEEP_LIB=orv_b200/liborv_b200_nat.so timeout 300 python tools/sweep_attention.py 600,1 1000,1 2>&1 | tee gpurun_out/r03a_attn_order_natural.log
echo "== event order"; ORVB_SWEEP_LIB=orv_b200/liborv_b200_exp.so timeout 400 python tools/sweep_attention.py 0,1 300,1 600,1 900,1 1200,1 1500,1 600,2 2>&1 | tee gpurun_out/r03a_attn_order_event.log

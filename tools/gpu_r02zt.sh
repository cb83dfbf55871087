#!/bin/bash
export ORVB_NO_BUILD=1
echo "== config 5, default"; timeout 300 python tools/time_clips.py 5 8 2>&1 | tail -2
echo "== config 5, generic gated epilogue"; ORVB_GEMM_FAST_RESID=0 timeout 300 python tools/time_clips.py 5 8 2>&1 | tail -2
echo "== config 5, pageable uploads + cuda-core tables"; ORVB_PINNED_UPLOADS=0 ORVB_MOD_TABLES_TC=0 timeout 300 python tools/time_clips.py 5 8 2>&1 | tail -2

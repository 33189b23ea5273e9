#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call10.log
{
echo "== fwd gen 2 counters (b 4096 x N 16384, 8 x 4)"
COSMOS_B200_DBG=1024 timeout 200 python tools/bwd_e_check.py 4096 16384 8 4 14.2857 2>&1 | grep -E "fwd2 prof|^fwd " | tail -14
echo "== fwd gen 1 counters"
COSMOS_B200_FWD=1 COSMOS_B200_DBG=1024 timeout 200 python tools/bwd_e_check.py 4096 16384 8 4 14.2857 2>&1 | grep -E "fwd prof|^fwd " | tail -8
} > $L 2>&1
cat $L

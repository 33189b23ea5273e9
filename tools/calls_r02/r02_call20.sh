#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call20.log
{
echo "== parity gen 3 (ragged, scale 100)"
timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 100 2>&1 | tail -8 | head -6
echo "== fwd gen 1 vs 3 (b 16384 x N 16384, 16 x 4)"
for g in 1 3 1 3; do COSMOS_B200_FWD=$g timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -2 | head -1; done
} > $L 2>&1
cat $L

#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call11.log
{
for d in 0 4096 2048; do
echo "== fwd gen 2, DBG=$d (b 16384 x N 16384, 16 x 4)"
COSMOS_B200_DBG=$d timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | grep "^fwd " | tail -1
done
} > $L 2>&1
cat $L

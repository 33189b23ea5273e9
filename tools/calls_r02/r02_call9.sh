#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call9.log
{
echo "== fwd gen 1 vs 2 (b 16384 x N 16384, 16 x 4)"
COSMOS_B200_FWD=1 timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -2
COSMOS_B200_FWD=2 timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -5
echo "== parity (ragged, scale 100), gen 2"
timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 100 2>&1 | tail -8
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
} > $L 2>&1
cat $L

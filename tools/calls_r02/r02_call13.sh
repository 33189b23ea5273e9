#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call13.log
{
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== per-rank shapes of an 8-GPU job (b 4096, N 32768): distillation 16 x 4, sweep slices 1 vs auto"
COSMOS_B200_ROWS_SPLITS=1 timeout 200 python tools/bwd_e_check.py 4096 32768 16 4 14.2857 t 2>&1 | tail -1
timeout 200 python tools/bwd_e_check.py 4096 32768 16 4 14.2857 t 2>&1 | tail -1
echo "== CLIP 8 x 2"
COSMOS_B200_ROWS_SPLITS=1 timeout 200 python tools/cols_check.py 4096 32768 2>&1 | tail -1
timeout 200 python tools/cols_check.py 4096 32768 2>&1 | tail -1
} > $L 2>&1
cat $L

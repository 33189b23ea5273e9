#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call14.log
{
echo "== parity, quad clusters (ragged, scale 100)"
timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 100 2>&1 | tail -8 | head -6
for q in 0 1; do
echo "== QUAD=$q: b 16384 x N 16384, 16 x 4"
COSMOS_B200_QUAD=$q timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -4 | head -3
echo "== QUAD=$q: per-rank shape of 8 GPUs b 4096 x N 32768, 16 x 4"
COSMOS_B200_QUAD=$q timeout 200 python tools/bwd_e_check.py 4096 32768 16 4 14.2857 t 2>&1 | tail -1
done
echo "== pytest infonce + fullsize"
timeout 600 python -m pytest tests/test_gpu_infonce.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -4
} > $L 2>&1
cat $L

#!/bin/bash
# forward epilogue: maximum tree + one-instruction warp maximum (CREDUX) - parity, then A/B against the packed-pair library
mkdir -p gpurun_out
L=gpurun_out/f2b_ab.log
P=tools/ab/libcosmos_b200_f2.so
{
echo "== parity (ragged, scale 100), new"
timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 100 2>&1 | tail -8 | head -7
echo "== f2 / new, three times (b 16384 x N 16384, 16 x 4)"
for k in 1 2 3; do
COSMOS_B200_LIB=$P timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -2 | head -1
timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -2 | head -1
done
echo "== pytest -m gpu (infonce)"
timeout 600 python -m pytest tests/test_gpu_infonce.py -m gpu -x -q 2>&1 | tail -3
} > $L 2>&1
cat $L

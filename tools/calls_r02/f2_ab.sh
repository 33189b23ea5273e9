#!/bin/bash
# packed fp32 pairs (FFMA2 / FADD2 / FMUL2) in the forward epilogue, the backward's factors and the retrieval sweep:
# parity, then A/B against the previous library in the same call
mkdir -p gpurun_out
L=gpurun_out/f2_ab.log
P=tools/ab/libcosmos_b200_prev.so
{
echo "== parity (ragged, scale 100), new"
timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 100 2>&1 | tail -8 | head -7
echo "== prev / new, three times (b 16384 x N 16384, 16 x 4)"
for k in 1 2 3; do
COSMOS_B200_LIB=$P timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -2 | head -1
timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -2 | head -1
done
echo "== retrieval prev / new"
COSMOS_B200_LIB=$P timeout 100 python tools/retrieval_time.py 2>&1 | tail -1
timeout 100 python tools/retrieval_time.py 2>&1 | tail -1
echo "== pytest -m gpu (infonce + retrieval + primitives)"
timeout 600 python -m pytest tests/test_gpu_infonce.py tests/test_gpu_retrieval.py tests/test_gpu_primitives.py -m gpu -x -q 2>&1 | tail -3
} > $L 2>&1
cat $L

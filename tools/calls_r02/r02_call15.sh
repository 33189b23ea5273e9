#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call15.log
{
for lib in tools/ab/libcosmos_b200_prev.so ""; do
echo "== lib=${lib:-current}"
COSMOS_B200_LIB=$lib timeout 200 python tools/cols_check.py 32768 32768 2>&1 | tail -1
COSMOS_B200_LIB=$lib timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 t 2>&1 | tail -1
COSMOS_B200_LIB=$lib timeout 200 python tools/bwd_e_check.py 4096 32768 16 4 14.2857 t 2>&1 | tail -1
done
echo "== pytest infonce (current)"
timeout 600 python -m pytest tests/test_gpu_infonce.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -3
} > $L 2>&1
cat $L

#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call8.log
{
echo "== pytest -m gpu (infonce + fullsize + dropin)"
timeout 900 python -m pytest tests/test_gpu_infonce.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -12
echo "== bench N=32768, column side from E"
timeout 400 python bench.py --no-extras --no-cpu-baseline --no-parity-check > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; tail -2 gpurun_out/bench_r02c.err
echo "== bench N=32768, column side through G tiles + GEMM"
COSMOS_B200_COLS=gemm timeout 400 python bench.py --no-extras --no-cpu-baseline --no-parity-check > gpurun_out/bench_r02c_gemm.json 2> gpurun_out/bench_r02c_gemm.err; tail -2 gpurun_out/bench_r02c_gemm.err
} > $L 2>&1
cat $L

#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call12.log
{
for f in 2 1; do
echo "== bench N=32768, forward generation $f"
COSMOS_B200_FWD=$f timeout 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench_r02d_fwd$f.json 2> gpurun_out/bench_r02d_fwd$f.err; tail -2 gpurun_out/bench_r02d_fwd$f.err
done
echo "== pytest infonce"
timeout 600 python -m pytest tests/test_gpu_infonce.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4
} > $L 2>&1
cat $L

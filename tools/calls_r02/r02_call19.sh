#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call19.log
{
echo "== parity gen 3 (ragged, scale 100 / 14)"
timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 100 2>&1 | tail -8
timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 14.2857 2>&1 | tail -8
echo "== fwd gen 1 vs 3 (b 16384 x N 16384, 16 x 4)"
COSMOS_B200_FWD=1 timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -3
COSMOS_B200_FWD=3 timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -3
COSMOS_B200_FWD=1 timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -3
COSMOS_B200_FWD=3 timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -3
echo "== pytest -m gpu (infonce + fullsize + dropin)"
timeout 900 python -m pytest tests/test_gpu_infonce.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -12
echo "== bench gen 3"
timeout 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench_r02_g3.json 2> gpurun_out/bench_r02_g3.err; tail -2 gpurun_out/bench_r02_g3.err; cat gpurun_out/bench_r02_g3.json | cut -c1-400
echo "== bench gen 1"
COSMOS_B200_FWD=1 timeout 400 python bench.py --no-extras --no-cpu-baseline --no-parity-check > gpurun_out/bench_r02_g1.json 2> gpurun_out/bench_r02_g1.err; tail -2 gpurun_out/bench_r02_g1.err; cat gpurun_out/bench_r02_g1.json | cut -c1-400
} > $L 2>&1
cat $L

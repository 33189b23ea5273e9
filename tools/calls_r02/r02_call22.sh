#!/bin/bash
# in-CTA merge of the forward's column partials: parity, A/B against the previous library in the same call
mkdir -p gpurun_out
L=gpurun_out/r02_call22.log
{
echo "== parity (ragged, scale 100 / 14.3), new"
timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 100 2>&1 | tail -8 | head -6
timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 14.2857 2>&1 | tail -8 | head -3
echo "== fwd prev / new / prev / new (b 16384 x N 16384, 16 x 4)"
for k in 1 2; do
COSMOS_B200_LIB=tools/ab/libcosmos_b200_prev.so timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -2 | head -1
timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -2 | head -1
done
echo "== pytest -m gpu (infonce + fullsize + dropin + primitives)"
timeout 900 python -m pytest tests/test_gpu_infonce.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py tests/test_gpu_primitives.py -m gpu -x -q 2>&1 | tail -5
echo "== bench new"
timeout 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench_r02_merge.json 2> gpurun_out/bench_r02_merge.err; tail -2 gpurun_out/bench_r02_merge.err; cut -c1-260 gpurun_out/bench_r02_merge.json
echo "== bench prev"
COSMOS_B200_LIB=tools/ab/libcosmos_b200_prev.so timeout 400 python bench.py --no-extras --no-cpu-baseline --no-parity-check > gpurun_out/bench_r02_premerge.json 2> gpurun_out/bench_r02_premerge.err; tail -2 gpurun_out/bench_r02_premerge.err; cut -c1-260 gpurun_out/bench_r02_premerge.json
} > $L 2>&1
cat $L

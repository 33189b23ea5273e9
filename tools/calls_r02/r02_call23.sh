#!/bin/bash
# column-partial merge, mbarrier variant: parity, A/B/C against the pre-merge and the bar.sync libraries in the same call
mkdir -p gpurun_out
L=gpurun_out/r02_call23.log
{
echo "== parity (ragged, scale 100), new"
timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 100 2>&1 | tail -8 | head -6
echo "== fwd premerge / bar.sync / mbarrier, twice (b 16384 x N 16384, 16 x 4)"
for k in 1 2; do
COSMOS_B200_LIB=tools/ab/libcosmos_b200_prev.so timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -2 | head -1
COSMOS_B200_LIB=tools/ab/libcosmos_b200_merge1.so timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -2 | head -1
timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -2 | head -1
done
echo "== pytest -m gpu (infonce + fullsize + dropin)"
timeout 900 python -m pytest tests/test_gpu_infonce.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -3
for v in new prev new prev; do
echo "== bench $v"
if [ $v = prev ]; then export COSMOS_B200_LIB=tools/ab/libcosmos_b200_prev.so; else unset COSMOS_B200_LIB; fi
timeout 400 python bench.py --no-extras --no-cpu-baseline --no-parity-check --no-e2e > gpurun_out/bench_r02_m2_$v.json 2> gpurun_out/bench_r02_m2_$v.err; tail -1 gpurun_out/bench_r02_m2_$v.err; cut -c1-230 gpurun_out/bench_r02_m2_$v.json
done
} > $L 2>&1
cat $L

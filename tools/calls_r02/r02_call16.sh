#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call16.log
{
for lib in tools/ab/libcosmos_b200_prev.so ""; do
echo "== lib=${lib:-current}"
COSMOS_B200_LIB=$lib timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -2 | head -1
COSMOS_B200_LIB=$lib timeout 200 python tools/cols_check.py 4096 32768 2>&1 | tail -1
done
echo "== pytest infonce (current)"
timeout 600 python -m pytest tests/test_gpu_infonce.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -3
for lib in tools/ab/libcosmos_b200_prev.so ""; do
echo "== bench lib=${lib:-current}"
COSMOS_B200_LIB=$lib timeout 300 python bench.py --no-extras --no-cpu-baseline --no-parity-check 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(r['ms_per_step'],2), round(r['value']), r['clocks']['sm_mhz'], {k:(round(v['ms_total']/r['steps'],2), round(v['flops_avg']/v['ms_avg']/1e9)) for k,v in r['roofline']['kernels'].items()})"
done
} > $L 2>&1
cat $L

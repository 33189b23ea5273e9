#!/bin/bash
# round 2, first GPU call: baseline of the unmeasured switches of infonce_bwd_e_kernel + in-kernel stall counters
mkdir -p gpurun_out
L=gpurun_out/r02_call1.log
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,driver_version --format=csv
bash tools/bwd_e_sweep.sh
echo "== stall counters, tensor prefetch"
COSMOS_B200_DBG=1024 timeout 100 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 t 2>&1 | tail -12
echo "== stall counters, bulk prefetch"
COSMOS_B200_DBG=1024 COSMOS_B200_EPREFETCH=bulk timeout 100 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 t 2>&1 | tail -12
echo "== both routes, all kernels"
timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -6
echo "== fwd stall counters (with E)"
COSMOS_B200_DBG=1024 timeout 100 python tools/bwd_e_check.py 4096 16384 8 4 14.2857 2>&1 | grep -E "fwd prof|fwd " | tail -12
echo "== pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
} > $L 2>&1
cat $L

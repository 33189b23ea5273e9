#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call3.log
{
run() { echo "== $*"; env "$@" timeout 100 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 t 2>&1 | tail -${TAILN:-1}; }
run COSMOS_B200_DBG=0
run COSMOS_B200_DBG=4096
run COSMOS_B200_DBG=16384
TAILN=6 run COSMOS_B200_DBG=1024
TAILN=6 run COSMOS_B200_DBG=5120
echo "== parity with DBG=4096 (ragged, scale 100)"
COSMOS_B200_DBG=4096 timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 100 2>&1 | tail -8 | head -6
echo "== parity with DBG=4096 (b 4096, N 16384, 8 x 4)"
COSMOS_B200_DBG=4096 timeout 100 python tools/bwd_e_check.py 4096 16384 8 4 14.2857 2>&1 | tail -5
} > $L 2>&1
cat $L

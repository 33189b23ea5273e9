#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call4.log
{
run() { echo "== $*"; env "$@" timeout 100 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 t 2>&1 | tail -${TAILN:-1}; }
run COSMOS_B200_DBG=0
TAILN=6 run COSMOS_B200_DBG=1024
TAILN=6 run COSMOS_B200_DBG=3072
} > $L 2>&1
cat $L

#!/bin/bash
# round 2: second-generation stored-exponential backward (16 scaling warps) against the first
mkdir -p gpurun_out
L=gpurun_out/r02_call2.log
{
run() { echo "== $*"; env "$@" timeout 100 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 t 2>&1 | tail -${TAILN:-1}; }
run COSMOS_B200_BWDE=1
run COSMOS_B200_BWDE=2
TAILN=14 run COSMOS_B200_BWDE=2 COSMOS_B200_DBG=1024
echo "== parity (ragged, scale 100)"
timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 100 2>&1 | tail -8
echo "== parity (b 4096, N 16384, 8 x 4)"
timeout 100 python tools/bwd_e_check.py 4096 16384 8 4 14.2857 2>&1 | tail -5
echo "== pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
} > $L 2>&1
cat $L

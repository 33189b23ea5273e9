#!/bin/bash
# forward epilogue variants in one call on one box: parity of the current library, then timing rounds
#   usage: fwd_abc.sh <log tag> <lib or "-" for the in-tree library> ...
TAG=$1; shift
mkdir -p gpurun_out
L=gpurun_out/fwd_abc_$TAG.log
{
echo "== parity (ragged, scale 100), in-tree library"
timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 100 2>&1 | tail -8 | head -7
echo "== parity (full tiles), in-tree library"
timeout 100 python tools/bwd_e_check.py 2048 4096 3 2 14.2857 2>&1 | tail -8 | head -4
for k in 1 2 3; do
for lib in "$@"; do
if [ "$lib" = "-" ]; then unset COSMOS_B200_LIB; else export COSMOS_B200_LIB=tools/ab/libcosmos_b200_$lib.so; fi
echo -n "$lib: "; timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -2 | head -1
done; done
unset COSMOS_B200_LIB
echo "== pytest -m gpu (infonce)"
timeout 600 python -m pytest tests/test_gpu_infonce.py -m gpu -x -q 2>&1 | tail -3
} > $L 2>&1
cat $L

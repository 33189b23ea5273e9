#!/bin/bash
# the whole step (bench.py, N = 32768, one GPU) with several libraries in one call on one box
#   usage: step_abc.sh <log tag> <lib or "-" for the in-tree library> ...
TAG=$1; shift
mkdir -p gpurun_out
L=gpurun_out/step_abc_$TAG.log
{
for k in 1 2; do
for lib in "$@"; do
if [ "$lib" = "-" ]; then unset COSMOS_B200_LIB; else export COSMOS_B200_LIB=tools/ab/libcosmos_b200_$lib.so; fi
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-parity-check --no-e2e 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels']
print('$lib: step %.2f ms  fwd %.3f  bwd_e %.3f  colgrad %.3f ms per launch  clocks %s' % (d['ms_per_step'], k['fwd']['ms_avg'], k['bwd_e']['ms_avg'], k['colgrad']['ms_avg'], d['clocks']['sm_mhz']))"
done; done
} > $L 2>&1
cat $L

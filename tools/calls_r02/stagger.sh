#!/bin/bash
# forward epilogue: start offsets between the four column groups (COSMOS_B200_FWD_STAGGER, cycles), whole step on one box
mkdir -p gpurun_out
L=gpurun_out/stagger_$1.log; shift
{
for k in 1 2; do
for st in "$@"; do
COSMOS_B200_FWD_STAGGER=$st timeout 300 python bench.py --no-extras --no-cpu-baseline --no-parity-check --no-e2e 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels']
print('stagger $st: step %.2f ms  fwd %.3f  bwd_e %.3f  colgrad %.3f ms per launch  clocks %s  loss %s' % (d['ms_per_step'], k['fwd']['ms_avg'], k['bwd_e']['ms_avg'], k['colgrad']['ms_avg'], d['clocks']['sm_mhz'], d['config']['loss']))"
done; done
} > $L 2>&1
cat $L

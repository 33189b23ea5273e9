#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call21.log
{
echo "== parity gen 3 (ragged, scale 100)"
timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 100 2>&1 | tail -8 | head -6
echo "== fwd gen 1 vs 3 (b 16384 x N 16384, 16 x 4)"
for g in 1 3 1 3; do COSMOS_B200_FWD=$g timeout 200 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 2>&1 | tail -2 | head -1; done
PROF="python bench.py --steps 1 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --no-parity-check"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:infonce_fwd3 -s 15 -c 1 -f -o gpurun_out/prof_fwd_r02g3 $PROF > gpurun_out/ncu_fwd_r02g3.log 2>&1
tail -2 gpurun_out/ncu_fwd_r02g3.log
} > $L 2>&1
cat $L

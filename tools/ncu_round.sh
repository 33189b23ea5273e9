#!/bin/bash
# ncu --set full captures of the three big kernels of the step (after the same command exited 0 without ncu)
TAG=${1:-r02b}
mkdir -p gpurun_out
PROF="python bench.py --steps 1 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --no-parity-check"
timeout 200 $PROF > gpurun_out/plain_$TAG.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:infonce_bwd_e2_kernel -s 15 -c 1 -f -o gpurun_out/prof_bwd_e_$TAG $PROF > gpurun_out/ncu_bwd_e_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:infonce_fwd -s 15 -c 1 -f -o gpurun_out/prof_fwd_$TAG $PROF > gpurun_out/ncu_fwd_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:infonce_bwd_e2t -s 3 -c 1 -f -o gpurun_out/prof_colgrad_$TAG $PROF > gpurun_out/ncu_colgrad_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*

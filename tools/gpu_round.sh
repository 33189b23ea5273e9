#!/bin/bash
# One gpurun call: GPU tests, smoke, the bench lines and the ncu captures that profiles/ summarises.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r02'        then here:  python profiles/summarize.py r02
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,driver_version --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
PROF="python bench.py --steps 1 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --no-parity-check"
# (the same command has to exit 0 without ncu first)
timeout 200 $PROF > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $PROF > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:infonce_bwd_e2_kernel -s 15 -c 1 -f -o gpurun_out/prof_bwd_e_$TAG $PROF > gpurun_out/ncu_bwd_e_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:infonce_fwd -s 15 -c 1 -f -o gpurun_out/prof_fwd_$TAG $PROF > gpurun_out/ncu_fwd_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:infonce_bwd_e2t -s 3 -c 1 -f -o gpurun_out/prof_colgrad_$TAG $PROF > gpurun_out/ncu_colgrad_$TAG.log 2>&1
ls -la gpurun_out | tail -20
tail -3 gpurun_out/pytest_gpu_$TAG.log
cut -c1-400 gpurun_out/bench_$TAG.json

#!/bin/bash
# One gpurun call: GPU tests, smoke, the bench lines and the ncu captures that profiles/ summarises.
#   gpurun --timeout 800 -- 'bash tools/gpu_round.sh r01b'
TAG=${1:-r01b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,driver_version --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
./tools/cluster_occ > gpurun_out/cluster_occ_$TAG.txt 2>&1
timeout 480 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 90 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_$TAG.log
timeout 300 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 90 python bench.py --global-batch 4096 --no-extras --no-cpu-baseline > gpurun_out/bench_4096_$TAG.json 2> gpurun_out/bench_4096_$TAG.err
PROF="python bench.py --steps 1 --warmup 3 --no-extras --no-e2e --no-cpu-baseline"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $PROF > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:infonce_bwd_e -s 15 -c 1 -f -o gpurun_out/prof_bwd_$TAG $PROF > gpurun_out/ncu_bwd_$TAG.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:infonce_fwd -s 15 -c 1 -f -o gpurun_out/prof_fwd_$TAG $PROF > gpurun_out/ncu_fwd_$TAG.log 2>&1
ls -la gpurun_out
tail -5 gpurun_out/pytest_gpu_$TAG.log
cat gpurun_out/bench_$TAG.json | cut -c1-600

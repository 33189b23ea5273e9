#!/bin/bash
# last single-GPU call of round 2 (the InfoNCE kernels are those of the r02g captures): GPU tests, smoke, the full bench line
TAG=r02j
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_$TAG.log
timeout 700 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/pytest_gpu_$TAG.log; tail -2 gpurun_out/smoke_$TAG.log; cut -c1-300 gpurun_out/bench_$TAG.json

#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call7.log
{
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== smoke"
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench N=32768 (no extras)"
timeout 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; tail -3 gpurun_out/bench_r02b.err
} > $L 2>&1
cat $L

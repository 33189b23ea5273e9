#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_call18.log
{
echo "== pytest pooler + primitives + retrieval"
timeout 600 python -m pytest tests/test_gpu_pooler.py tests/test_gpu_primitives.py tests/test_gpu_retrieval.py -m gpu -q 2>&1 | tail -5
echo "== pooler step under the ncu launch list (196 tokens)"
timeout 120 python tools/pooler_step.py 196 > gpurun_out/pooler_plain.log 2>&1 && timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pooler_r02.csv python tools/pooler_step.py 196 > gpurun_out/ncu_pooler.log 2>&1
tail -2 gpurun_out/ncu_pooler.log
} > $L 2>&1
cat $L

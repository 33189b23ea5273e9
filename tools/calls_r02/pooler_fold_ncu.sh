#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/pooler_step.py 196 > gpurun_out/pooler_plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pooler_fold.csv python tools/pooler_step.py 196 > gpurun_out/ncu_pooler_fold.log 2>&1
tail -2 gpurun_out/ncu_pooler_fold.log

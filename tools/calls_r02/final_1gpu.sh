#!/bin/bash
# final single-GPU call of round 2: tests, smoke, full bench line, launch list + three ncu captures, reference arm
bash tools/gpu_round.sh r02g > gpurun_out/final_1gpu.log 2>&1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_r02g.json 2> gpurun_out/bench_reference_r02g.err
timeout 200 python tools/timeline.py --free-run > gpurun_out/timeline_1gpu_r02g.txt 2> gpurun_out/timeline_1gpu_r02g.err
tail -25 gpurun_out/final_1gpu.log; cut -c1-300 gpurun_out/bench_reference_r02g.json; head -5 gpurun_out/timeline_1gpu_r02g.txt

#!/bin/bash
# final single-GPU call of round 2, second session (packed fp32 pairs, add_zero_attn, full-size headline parity test):
# tests, smoke, full bench line, launch list + three ncu captures, reference arm, timeline
bash tools/gpu_round.sh r02m > gpurun_out/final_1gpu_m.log 2>&1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_r02m.json 2> gpurun_out/bench_reference_r02m.err
timeout 200 python tools/timeline.py --free-run > gpurun_out/timeline_1gpu_r02m.txt 2> gpurun_out/timeline_1gpu_r02m.err
tail -12 gpurun_out/final_1gpu_m.log | cut -c1-600; grep -a "fullsize headline" gpurun_out/pytest_gpu_r02m.log; cut -c1-300 gpurun_out/bench_reference_r02m.json; head -5 gpurun_out/timeline_1gpu_r02m.txt

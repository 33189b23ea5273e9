#!/bin/bash
# final 8-GPU call of round 2: the bench line (with its parity_check leg) and the free-running timeline
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $RUN --master-port 29541 bench.py --gpus 8 > gpurun_out/bench_8gpu_r02f.json 2> gpurun_out/bench_8gpu_r02f.err
echo "bench exit $?"; tail -2 gpurun_out/bench_8gpu_r02f.err; cut -c1-300 gpurun_out/bench_8gpu_r02f.json
timeout 300 $RUN --master-port 29543 tools/timeline.py --free-run > gpurun_out/timeline_8gpu_r02f.txt 2> gpurun_out/timeline_8gpu_r02f.err
echo "timeline exit $?"; head -30 gpurun_out/timeline_8gpu_r02f.txt

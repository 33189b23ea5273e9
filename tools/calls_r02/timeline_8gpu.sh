#!/bin/bash
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29543 tools/timeline.py --free-run > gpurun_out/timeline_8gpu_r02f.txt 2> gpurun_out/timeline_8gpu_r02f.err
echo "timeline exit $?"; head -60 gpurun_out/timeline_8gpu_r02f.txt

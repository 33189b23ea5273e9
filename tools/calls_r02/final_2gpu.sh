#!/bin/bash
# final 2-GPU call of round 2: multi-rank NCCL parity tests, the bench line, the free-running timeline
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/pytest_multirank_2gpu_r02f.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multirank_2gpu_r02f.log; tail -3 gpurun_out/pytest_multirank_2gpu_r02f.log
timeout 500 $RUN --master-port 29541 bench.py --gpus 2 > gpurun_out/bench_2gpu_r02f.json 2> gpurun_out/bench_2gpu_r02f.err
echo "bench exit $?"; cut -c1-300 gpurun_out/bench_2gpu_r02f.json
timeout 300 $RUN --master-port 29543 tools/timeline.py --free-run > gpurun_out/timeline_2gpu_r02f.txt 2> gpurun_out/timeline_2gpu_r02f.err
echo "timeline exit $?"; head -40 gpurun_out/timeline_2gpu_r02f.txt

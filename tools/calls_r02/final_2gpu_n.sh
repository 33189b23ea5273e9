#!/bin/bash
# 2 B200s, final code of round 2 (second session): multi-rank NCCL parity tests, then the bench line
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/pytest_multirank_2gpu_r02n.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multirank_2gpu_r02n.log; tail -3 gpurun_out/pytest_multirank_2gpu_r02n.log
timeout 300 $RUN --master-port 29541 bench.py --gpus 2 > gpurun_out/bench_2gpu_r02n.json 2> gpurun_out/bench_2gpu_r02n.err
echo "bench exit $?"; cut -c1-260 gpurun_out/bench_2gpu_r02n.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_2gpu_r02n.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['parity_check']['ok'], d['parity_check']['min_grad_cos'], d['clocks'])"

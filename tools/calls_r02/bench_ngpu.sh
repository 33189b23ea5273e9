#!/bin/bash
# bench line (with its parity_check leg) at N GPUs:  bash tools/calls_r02/bench_ngpu.sh N tag
N=$1; TAG=$2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $RUN --master-port 29541 bench.py --gpus $N > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err
echo "bench exit $?"; grep -v "^W\|Warn\|warn" gpurun_out/bench_${N}gpu_$TAG.err | tail -5; cut -c1-250 gpurun_out/bench_${N}gpu_$TAG.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu_$TAG.json').read().strip().splitlines()[-1])
r=d['roofline']
print(d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'minus', r.get('step_minus_kernels_ms'), {k:round(v['ms_total']/d['steps'],2) for k,v in r['kernels'].items()}, d['clocks'], 'parity', d['parity_check']['ok'], d['parity_check']['min_grad_cos'], d['parity_check']['max_loss_rel'], 'frac', r['step_frac'])
PY

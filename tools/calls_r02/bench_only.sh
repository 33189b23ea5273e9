#!/bin/bash
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 700 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "exit $?"; tail -3 gpurun_out/bench_$TAG.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['traffic'])
c=d['config2_global_batch_4096']; print('cfg2', c['ms_per_step'], c['cuda_graph'])
s=d['same_size_cpu_vs_gpu']; print('same', s['gpu_ms_per_step'], s['gpu_cuda_graph_ms_per_step'], s['cpu_port_ms_per_step'])
print({k:(round(v['ms'],3), round(v['cuda_graph_ms'],3)) for k,v in d['xattn'].items()})
PY

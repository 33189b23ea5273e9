#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/pooler_fold2.log
{
echo "== pytest tests/test_gpu_pooler.py"
timeout 600 python -m pytest tests/test_gpu_pooler.py -m gpu -x -q 2>&1 | tail -4
for mode in folded unfolded; do
echo "== timings (bench_xattn): $mode"
if [ $mode = unfolded ]; then export COSMOS_B200_POOLER=unfolded; fi
timeout 300 python - <<'PY' 2>&1 | tail -5
import torch, bench
dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
out = bench.bench_xattn(dev, flush)
for k, v in out.items():
    print(k, "ms %.3f best %.3f" % (v["ms"], v["ms_best"]))
PY
done
unset COSMOS_B200_POOLER
timeout 200 python tools/pooler_step.py 196 > gpurun_out/pooler_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_pooler_fold.csv python tools/pooler_step.py 196 > gpurun_out/ncu_pooler_fold.log 2>&1
} > $L 2>&1
cat $L

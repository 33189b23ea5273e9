#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/pooler_fold3.log
{
echo "== pytest tests/test_gpu_pooler.py, every shape folded"
COSMOS_B200_POOLER_FOLD_MAX=8192 timeout 600 python -m pytest tests/test_gpu_pooler.py -m gpu -q 2>&1 | tail -4
echo "== pytest tests/test_gpu_pooler.py, default"
timeout 600 python -m pytest tests/test_gpu_pooler.py -m gpu -q 2>&1 | tail -2
for mx in 8192 128; do
echo "== timings (bench_xattn): fold max $mx"
COSMOS_B200_POOLER_FOLD_MAX=$mx timeout 300 python - <<'PY' 2>&1 | tail -5
import torch, bench
dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
out = bench.bench_xattn(dev, flush)
for k, v in out.items():
    print(k, "ms %.3f best %.3f" % (v["ms"], v["ms_best"]), "peak mem GB %.1f" % (torch.cuda.max_memory_allocated() / 2**30))
PY
done
} > $L 2>&1
cat $L

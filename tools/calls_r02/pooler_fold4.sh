#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/pooler_fold4.log
{
echo "== pytest tests/test_gpu_pooler.py"
timeout 600 python -m pytest tests/test_gpu_pooler.py -m gpu -q 2>&1 | tail -6
for core in gemm cuda_cores; do
echo "== timings (bench_xattn): key / value route, attention core $core"
COSMOS_B200_POOLER=unfolded COSMOS_B200_POOLER_CORE=$core timeout 300 python - <<'PY' 2>&1 | tail -5
import torch, bench
dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
out = bench.bench_xattn(dev, flush)
for k, v in out.items():
    print(k, "ms %.3f best %.3f" % (v["ms"], v["ms_best"]), "peak mem GB %.1f" % (torch.cuda.max_memory_allocated() / 2**30))
PY
done
echo "== timings (bench_xattn): default"
timeout 300 python - <<'PY' 2>&1 | tail -5
import torch, bench
dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
out = bench.bench_xattn(dev, flush)
for k, v in out.items():
    print(k, "ms %.3f best %.3f" % (v["ms"], v["ms_best"]))
PY
} > $L 2>&1
cat $L

#!/bin/bash
# folded attention of the pooler: GPU parity tests, parity tool, timings folded vs unfolded in the same call
mkdir -p gpurun_out
L=gpurun_out/pooler_fold.log
{
echo "== pytest tests/test_gpu_pooler.py"
timeout 600 python -m pytest tests/test_gpu_pooler.py -m gpu -x -q 2>&1 | tail -15
echo "== pooler_parity (folded)"
timeout 200 python tools/pooler_parity.py 2>&1 | tail -12
echo "== timings (bench_xattn): folded"
timeout 300 python - <<'PY' 2>&1 | tail -8
import json, torch, bench
dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
out = bench.bench_xattn(dev, flush)
for k, v in out.items():
    print(k, "ms %.3f best %.3f" % (v["ms"], v["ms_best"]), {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("algorithmic_tflops", "frac_of_bf16_peak")}, v.get("hbm_roofline", {}).get("frac"))
PY
echo "== timings (bench_xattn): unfolded"
COSMOS_B200_POOLER=unfolded timeout 300 python - <<'PY' 2>&1 | tail -8
import json, torch, bench
dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
out = bench.bench_xattn(dev, flush)
for k, v in out.items():
    print(k, "ms %.3f best %.3f" % (v["ms"], v["ms_best"]))
PY
} > $L 2>&1
cat $L

"""Ad-hoc: retrieval ranks at 32768 x 32768 x 512 fp32, event-timed - for A/B runs (COSMOS_B200_LIB selects the library)."""
import sys, torch
sys.path.insert(0, ".")
from cosmos_b200.retrieval import retrieval_ranks
n, D = 32768, 512
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(21)
z = torch.randn(n, D, generator=g, device=dev)
img = torch.nn.functional.normalize(z + 2.0 * torch.randn(n, D, generator=g, device=dev), dim=-1)
txt = torch.nn.functional.normalize(z + 2.0 * torch.randn(n, D, generator=g, device=dev), dim=-1)
r = retrieval_ranks(img, txt)
ts = []
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = retrieval_ranks(img, txt); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("retrieval ranks %d^2 x %d: median %.2f ms best %.2f ms (%.1f TFLOP/s fp32)  checksum %d" % (
    n, D, sorted(ts)[len(ts) // 2], min(ts), 2.0 * n * n * D / (min(ts) * 1e-3) / 1e12, int(r.long().sum())))

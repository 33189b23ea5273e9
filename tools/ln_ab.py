"""Ad-hoc: LayerNorm backward / forward on the pooler's token shape, event-timed (L2 flushed) - for A/B runs of row maps."""
import sys, torch
sys.path.insert(0, ".")
from cosmos_b200 import pooler
dev = torch.device("cuda", 0)
rows, dim = 1024 * 196, 512
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(rows, dim, generator=g, device=dev).bfloat16()
dy = torch.randn(rows, dim, generator=g, device=dev).bfloat16()
w = torch.ones(dim, device=dev); b = torch.zeros(dim, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
y, mean, rstd = pooler._ln_fwd(x, w, b, torch.bfloat16)
dx = torch.empty_like(x)
dw = torch.zeros(dim, device=dev); db = torch.zeros(dim, device=dev)
def timed(fn, n=10):
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]
f = timed(lambda: pooler._ln_fwd(x, w, b, torch.bfloat16))
bw = timed(lambda: pooler._ln_bwd(dy, x, w, mean, rstd, dx, False, dw, db))
print("ln_fwd median %.1f us best %.1f (%.0f GB/s)   ln_bwd median %.1f us best %.1f (%.0f GB/s)" % (
    f[0] * 1e3, f[1] * 1e3, 2 * rows * dim * 2 / f[0] / 1e6, bw[0] * 1e3, bw[1] * 1e3, 3 * rows * dim * 2 / bw[0] / 1e6))

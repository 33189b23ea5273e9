"""Ad-hoc GPU debug: run fwd / bwd of one 8x4 group at growing sizes with a sync after each launch."""
import sys, time, torch
sys.path.insert(0, ".")
from cosmos_b200 import infonce as K

def run(b, D, gx=8, gy=4):
    g = torch.Generator().manual_seed(0)
    x = torch.nn.functional.normalize(torch.randn(gx, b, D, generator=g), dim=-1).bfloat16().cuda()
    y = torch.nn.functional.normalize(torch.randn(gy, b, D, generator=g), dim=-1).bfloat16().cuda()
    sc = torch.tensor([14.2857], device="cuda")
    up = torch.ones(1, device="cuda")
    torch.cuda.synchronize()
    t0 = time.time()
    row, diag, col = K._k_fwd(x, y, 0, sc)
    torch.cuda.synchronize()
    t1 = time.time()
    print(f"b={b} D={D} fwd ok {1e3*(t1-t0):.2f} ms", flush=True)
    dx, ds = K._k_bwd(x, y, 0, sc, row, col, 1.0, 1.0, 1.0, 1.0, 1.0 / (2 * b * gx * gy), up, True, True)
    torch.cuda.synchronize()
    t2 = time.time()
    print(f"b={b} D={D} bwd ok {1e3*(t2-t1):.2f} ms  dscale={float(ds):.5f} dxnorm={float(dx.float().norm()):.5f}", flush=True)

for b, D in [(1024, 512), (2048, 512), (4096, 512), (4096, 512), (8192, 512)]:
    run(b, D)

"""Ad-hoc GPU check: 4-CTA-cluster backward (COSMOS_B200_DBG=128) against the pair kernel on the same inputs."""
import os, sys, torch
sys.path.insert(0, ".")
from cosmos_b200 import infonce as K

def run(b, gx, gy, W=1, rank=0, reps=5):
    g = torch.Generator().manual_seed(b + gx)
    D = 512
    N = W * b
    z = torch.randn(b, D, generator=g)
    x = torch.nn.functional.normalize(z + 2 * torch.randn(gx, b, D, generator=g), dim=-1).bfloat16().cuda()
    y = torch.nn.functional.normalize(torch.randn(gy, N, D, generator=g), dim=-1)
    y[:, rank * b:(rank + 1) * b] = torch.nn.functional.normalize(z + 2 * torch.randn(gy, b, D, generator=g), dim=-1)
    y = y.bfloat16().cuda()
    sc = torch.tensor([14.2857], device="cuda")
    up = torch.ones(1, device="cuda")
    row, diag, col = K._k_fwd(x, y, rank * b, sc)
    w = 1.0 / (2 * N * gx * gy)
    os.environ["COSMOS_B200_DBG"] = "128"
    dx0, ds0 = K._k_bwd(x, y, rank * b, sc, row, col, 1.0, 1.0, 1.0, 1.0, w, up, True, True)
    torch.cuda.synchronize()
    os.environ["COSMOS_B200_DBG"] = "0"
    dx1, ds1 = K._k_bwd(x, y, rank * b, sc, row, col, 1.0, 1.0, 1.0, 1.0, w, up, True, True)
    torch.cuda.synchronize()
    a, c = dx0.float().flatten().double(), dx1.float().flatten().double()
    cos = float(a @ c / (a.norm() * c.norm()))
    print(f"b={b} gx={gx} gy={gy} W={W}: cos={cos:.7f} norm ratio={float(c.norm()/a.norm()):.5f} dscale {float(ds0):.6e} vs {float(ds1):.6e} "
          f"maxdiff={float((dx0.float()-dx1.float()).abs().max()):.3e} nan={bool(torch.isnan(dx1.float()).any())}", flush=True)
    tot = {"0": 0.0, "128": 0.0, "512": 0.0}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for rep in range(3 + 2 * reps):
        for mode in ("0", "128", "512"):
            os.environ["COSMOS_B200_DBG"] = mode
            ev[0].record(); K._k_bwd(x, y, rank * b, sc, row, col, 1.0, 1.0, 1.0, 1.0, w, up, True, True); ev[1].record()
            torch.cuda.synchronize()
            if rep >= 3:
                tot[mode] += ev[0].elapsed_time(ev[1])
    print(f"   cluster-of-4 (default): {tot['0']/(2*reps):.3f} ms   pair kernel (DBG=128): {tot['128']/(2*reps):.3f} ms  quad, 2-slot ring A: {tot['512']/(2*reps):.3f} ms", flush=True)

for args in [(256, 1, 1), (384, 2, 3), (4096, 8, 4), (8192, 8, 4)]:
    run(*args)
run(512, 2, 2, W=2, rank=1)

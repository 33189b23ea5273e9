"""Ad-hoc GPU diagnostics: can the 16 SMs a cluster-of-4 launch cannot use (33 clusters = 132 of 148 SMs, tools/cluster_occ.cu)
do other work while infonce_bwd_quad_kernel runs?  Times, with CUDA events on the default stream:
  * the quad kernel alone on 16 / 15 / 14 row tensors, the pair kernel (cluster of 2) alone on 1 / 2 tensors,
  * quad on 16-k tensors on one stream next to the pair kernel on k tensors on another, in both launch orders,
  * quad on 16 tensors next to the column-gradient GEMM restricted to 16 persistent CTAs, in both launch orders.
    python tools/overlap_probe.py [rows_per_tensor] [reps]
"""
import os
import sys

import torch

sys.path.insert(0, ".")
from cosmos_b200 import infonce as K  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
D, GX, GY = 512, 16, 4
g = torch.Generator().manual_seed(0)


def feats(n):
    return torch.nn.functional.normalize(torch.randn(n, b, D, generator=g), dim=-1).bfloat16().cuda()


x, y = feats(GX), feats(GY)
sc = torch.tensor([14.2857], device="cuda")
up = torch.ones(1, device="cuda")
row, diag, col = K._k_fwd(x, y, 0, sc)
cur = torch.cuda.current_stream()
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
sH = torch.cuda.Stream(priority=-1)          # CTAs of a higher-priority stream are dispatched before pending lower-priority ones
MODE = os.environ.get("PROBE_MODE", "order")
out = {}


def bwd(lo, hi, pair_kernel, key):
    def run():
        if pair_kernel:
            os.environ["COSMOS_B200_DBG"] = "128"
        try:
            dx, _ = K._k_bwd(x[lo:hi], y, 0, sc, row[lo * GY:hi * GY], col[lo * GY:hi * GY], 1.0, 1.0, 1.0, 1.0, 1.0, up, True, False)
        finally:
            os.environ.pop("COSMOS_B200_DBG", None)
        out[key] = dx
    return run


n_c = 2
gt = torch.randn(8 * b, n_c * b, generator=torch.Generator(device="cuda").manual_seed(1), device="cuda", dtype=torch.bfloat16)
x2d = x[:8].reshape(8 * b, D)


def gemm(ctas):
    def run():
        if ctas:
            os.environ["COSMOS_B200_GEMM_CTAS"] = str(ctas)
        try:
            out["gemm"] = K._k_colgrad(gt, x2d, n_c, b)
        finally:
            os.environ.pop("COSMOS_B200_GEMM_CTAS", None)
    return run


def alone(fn):
    return lambda: fn()


def both(first, second, s2=None, spin_us=0):
    s2 = s2 or sB

    def run():
        sA.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(sA):
            first()
        with torch.cuda.stream(s2):
            if spin_us:
                torch.cuda._sleep(int(spin_us * 1900))      # ~cycles: let the first grid fill the machine
            second()
        cur.wait_stream(sA)
        cur.wait_stream(s2)
    return run


def timed(name, fn):
    ms = []
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = sorted(ms[1:])
    print(f"{name:42s} median {ms[len(ms) // 2]:8.3f} ms   min {ms[0]:8.3f}", flush=True)
    return ms[len(ms) // 2]


for _ in range(3):      # warm the clocks / power state
    bwd(0, 16, False, "w")()
torch.cuda.synchronize()
timed("quad 16 tensors", alone(bwd(0, 16, False, "q16")))
ref = out["q16"].float()
if MODE == "priority":
    timed("colgrad GEMM, 148 CTAs", alone(gemm(0)))
    for ctas in (12, 16, 20):
        timed("colgrad GEMM, %d CTAs" % ctas, alone(gemm(ctas)))
        for spin in (0, 20, 200):
            timed("quad 16 | GEMM %d CTAs high prio, +%d us" % (ctas, spin), both(bwd(0, 16, False, "a"), gemm(ctas), sH, spin))
    timed("quad 15 | pair 1 high prio, +20 us", both(bwd(0, 15, False, "a"), bwd(15, 16, True, "b"), sH, 20))
    timed("quad 16 | GEMM 148 CTAs high prio, +20 us", both(bwd(0, 16, False, "a"), gemm(0), sH, 20))
    sys.exit(0)
timed("quad 15 tensors", alone(bwd(0, 15, False, "q15")))
timed("quad 14 tensors", alone(bwd(0, 14, False, "q14")))
timed("pair kernel 1 tensor", alone(bwd(15, 16, True, "p1")))
timed("pair kernel 2 tensors", alone(bwd(14, 16, True, "p2")))
timed("quad 15 | pair 1   (quad first)", both(bwd(0, 15, False, "a"), bwd(15, 16, True, "b")))
got = torch.cat([out["a"], out["b"]]).float()
print("   hybrid dx vs quad-only dx: cosine %.7f, max abs diff %.3e" % (
    float((got * ref).sum() / (got.norm() * ref.norm())), float((got - ref).abs().max())), flush=True)
timed("quad 15 | pair 1   (pair first)", both(bwd(15, 16, True, "b"), bwd(0, 15, False, "a")))
timed("quad 14 | pair 2   (quad first)", both(bwd(0, 14, False, "a"), bwd(14, 16, True, "b")))
timed("quad 14 | pair 2   (pair first)", both(bwd(14, 16, True, "b"), bwd(0, 14, False, "a")))
timed("colgrad GEMM, 148 CTAs", alone(gemm(0)))
timed("colgrad GEMM, 16 CTAs", alone(gemm(16)))
timed("colgrad GEMM, 8 CTAs", alone(gemm(8)))
timed("quad 16 | GEMM 16 CTAs (quad first)", both(bwd(0, 16, False, "a"), gemm(16)))
timed("quad 16 | GEMM 16 CTAs (GEMM first)", both(gemm(16), bwd(0, 16, False, "a")))
timed("quad 16 | GEMM 8 CTAs (quad first)", both(bwd(0, 16, False, "a"), gemm(8)))
timed("quad 16 | GEMM 148 CTAs (quad first)", both(bwd(0, 16, False, "a"), gemm(0)))
timed("quad 16 then GEMM 148 CTAs (one stream)", lambda: (bwd(0, 16, False, "a")(), gemm(0)()))

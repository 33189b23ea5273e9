"""Ad-hoc GPU check of the stored-exponential route (fwd_e + bwd_e) against the recompute route (fwd + quad backward):
same inputs, same mode scalars; reports cosine / max difference of dX, relative difference of d(scale), of the stored G
tiles, and the time of both routes.    python tools/bwd_e_check.py [b] [N] [gx] [gy] [scale]"""
import sys

import torch

sys.path.insert(0, ".")
from cosmos_b200 import infonce as K  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
gx = int(sys.argv[3]) if len(sys.argv) > 3 else 3
gy = int(sys.argv[4]) if len(sys.argv) > 4 else 2
scale = float(sys.argv[5]) if len(sys.argv) > 5 else 14.2857
D = 512
g = torch.Generator().manual_seed(b + N)
z = torch.randn(N, D, generator=g)


def feats(n, rows):
    return torch.nn.functional.normalize(z[None, :rows] + 1.5 * torch.randn(n, rows, D, generator=g), dim=-1).bfloat16().cuda()


y = feats(gy, N)
x = feats(gx, b)
sc = torch.tensor([scale], device="cuda")
up = torch.tensor([1.7], device="cuda")
mix = (1.0, 1.0, 1.0, 1.0, 0.25)      # a_row, a_col, s_row, s_col, weight

row, diag, col = K._k_fwd(x, y, 0, sc)
dx_ref, ds_ref = K._k_bwd(x, y, 0, sc, row, col, *mix, up, True, True)
row2, diag2, col2, e, off = K._k_fwd(x, y, 0, sc, keep_e=True)
torch.cuda.synchronize()
print("fwd statistics identical:", torch.equal(row, row2), torch.equal(col, col2), torch.equal(diag, diag2), flush=True)
dx, ds = K._k_bwd_e(x, y, 0, sc, e, off, diag2, row2, col2, *mix, up, True)
torch.cuda.synchronize()
a, r = dx.float().flatten(), dx_ref.float().flatten()
print("dX cosine %.7f  max|diff| %.3e  max|ref| %.3e  |dX|/|ref| %.5f" % (
    float(a @ r / (a.norm() * r.norm())), float((a - r).abs().max()), float(r.abs().max()), float(a.norm() / r.norm())), flush=True)
print("dscale %.6e vs %.6e (rel %.2e)" % (float(ds), float(ds_ref), abs(float(ds) - float(ds_ref)) / abs(float(ds_ref))), flush=True)

# fp64 reference of one pair's gradient from first principles
SMALL = b * N <= (1 << 24)
if SMALL:
    xi, yj = x[1].double(), y[gy - 1].double()
    S = scale * xi @ yj.t()
    R = torch.softmax(S, dim=1)
    lse_c = torch.log2(torch.exp2(col[1 * gy + gy - 1].double()))          # complete here: one rank
    Cm = torch.exp(S - lse_c[None, :] * 0.6931471805599453)
    Gm = R + Cm
    Gm[torch.arange(b), torch.arange(b)] -= 2.0
    want = (Gm @ yj) * scale * 0.25 * 1.7
    # the kernels sum over the gy column tensors: isolate tensor gy-1 by a second launch with gy = 1
    r1, d1, c1, e1, o1 = K._k_fwd(x[1:2], y[gy - 1:gy], 0, sc, keep_e=True)
    dx1, _ = K._k_bwd_e(x[1:2], y[gy - 1:gy], 0, sc, e1, o1, d1, r1, c1, *mix, up, False)
    dxq, _ = K._k_bwd(x[1:2], y[gy - 1:gy], 0, sc, r1, c1, *mix, up, True, False)
    w = want.flatten()
    for name, t in (("stored-e", dx1), ("recompute", dxq)):
        v = t[0].double().flatten()
        print("%-10s vs fp64: cosine %.8f  rel L2 %.3e" % (name, float(v @ w / (v.norm() * w.norm())), float((v - w).norm() / w.norm())), flush=True)

if N % 8 == 0 and SMALL:
    g1 = torch.zeros(gx * b, gy * N, dtype=x.dtype, device="cuda")
    g2 = torch.zeros_like(g1)
    K._k_bwd(x, y, 0, sc, row, col, *mix, up, True, False, g1)
    K._k_bwd_e(x, y, 0, sc, e, off, diag2, row2, col2, *mix, up, False, g2)
    torch.cuda.synchronize()
    print("G tiles: cosine %.7f  max|diff| %.3e" % (float((g1.float() * g2.float()).sum() / (g1.float().norm() * g2.float().norm())),
                                                   float((g1.float() - g2.float()).abs().max())), flush=True)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


fl = 2.0 * gx * gy * b * N * D
if len(sys.argv) > 6:      # timing of the stored-exponential backward only (diagnostic flags in the environment)
    t_be = timed(lambda: K._k_bwd_e(x, y, 0, sc, e, off, diag2, row2, col2, *mix, up, True))
    print("bwd stored-e %.3f ms (%.0f TF/s)" % (t_be, fl / t_be * 1e-9), flush=True)
    sys.exit(0)
t_f = timed(lambda: K._k_fwd(x, y, 0, sc))
t_fe = timed(lambda: K._k_fwd(x, y, 0, sc, keep_e=True))
t_b = timed(lambda: K._k_bwd(x, y, 0, sc, row, col, *mix, up, True, True))
t_be = timed(lambda: K._k_bwd_e(x, y, 0, sc, e, off, diag2, row2, col2, *mix, up, True))
print("fwd %.3f ms (%.0f TF/s)   fwd+E %.3f ms (%.0f TF/s)   bwd recompute %.3f ms   bwd stored-e %.3f ms (%.0f TF/s)" % (
    t_f, fl / t_f * 1e-9, t_fe, fl / t_fe * 1e-9, t_b, t_be, fl / t_be * 1e-9), flush=True)
print("route total: recompute %.3f ms, stored-e %.3f ms" % (t_f + t_b, t_fe + t_be), flush=True)

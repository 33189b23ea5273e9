"""Ad-hoc: time the two stored-exponential backward kernels on the CLIP group's shapes (8 row tensors x 2 column tensors).
    python tools/cols_check.py [b] [N]"""
import sys
import torch
sys.path.insert(0, ".")
from cosmos_b200 import infonce as K  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
gx, gy, D = 8, 2, 512
g = torch.Generator(device="cuda").manual_seed(1)
y = torch.nn.functional.normalize(torch.randn(gy, N, D, generator=g, device="cuda"), dim=-1).bfloat16()
x = torch.nn.functional.normalize(torch.randn(gx, b, D, generator=g, device="cuda"), dim=-1).bfloat16()
sc = torch.tensor([14.2857], device="cuda")
up = torch.tensor([1.0], device="cuda")
row, diag, col, e, off = K._k_fwd(x, y, 0, sc, True)
torch.cuda.synchronize()


def timed(fn, reps=3):
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
t_rows = timed(lambda: K._k_bwd_e(x, y, 0, sc, e, off, diag, row, col, 1.0, 1.0, 1.0, 1.0, 0.125, up, True))
t_cols = timed(lambda: K._k_bwd_e_cols(x, y, 0, sc, e, off, diag, row, col, 1.0, 1.0))
t_fwd = timed(lambda: K._k_fwd(x, y, 0, sc, True))
print("fwd+E %.2f ms (%.0f TF/s)  rows (dX = G Y) %.2f ms (%.0f TF/s)   columns (dY = G^T X) %.2f ms (%.0f TF/s)" % (
    t_fwd, fl / t_fwd * 1e-9, t_rows, fl / t_rows * 1e-9, t_cols, fl / t_cols * 1e-9))

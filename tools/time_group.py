"""Ad-hoc GPU diagnostics: event-timed fwd / bwd of one 8x4 group (COSMOS_B200_DBG selects what is skipped)."""
import os, sys, torch
sys.path.insert(0, ".")
from cosmos_b200 import infonce as K

def run(b, D, gx=8, gy=4, reps=5):
    g = torch.Generator().manual_seed(0)
    x = torch.nn.functional.normalize(torch.randn(gx, b, D, generator=g), dim=-1).bfloat16().cuda()
    y = torch.nn.functional.normalize(torch.randn(gy, b, D, generator=g), dim=-1).bfloat16().cuda()
    sc = torch.tensor([14.2857], device="cuda")
    up = torch.ones(1, device="cuda")
    row, diag, col = K._k_fwd(x, y, 0, sc)
    K._k_bwd(x, y, 0, sc, row, col, 1.0, 1.0, 1.0, 1.0, 1.0, up, True, True)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(reps):
        ev[0].record(); K._k_fwd(x, y, 0, sc); ev[1].record()
        K._k_bwd(x, y, 0, sc, row, col, 1.0, 1.0, 1.0, 1.0, 1.0, up, True, True); ev[2].record()
        torch.cuda.synchronize()
        tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
    fl = 2.0 * gx * gy * b * b * D
    print(f"DBG={os.environ.get('COSMOS_B200_DBG','0')} b={b}: fwd {tf/reps:.3f} ms ({fl/(tf/reps)*1e-9:.0f} TF/s)  "
          f"bwd {tb/reps:.3f} ms (exec {3*fl/(tb/reps)*1e-9:.0f} TF/s)", flush=True)

run(int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 512, reps=int(sys.argv[2]) if len(sys.argv) > 2 else 5)

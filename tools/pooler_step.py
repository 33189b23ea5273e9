"""Ad-hoc GPU diagnostics: one pooler fwd+bwd step (config 4 shapes) and EMA launches, for ncu launch lists."""
import sys, torch
sys.path.insert(0, ".")
from cosmos_b200.pooler import AttentionalCrossPooler, crossmodal_features
from cosmos_b200 import ema_update_
from oracle import cosmos_oracle as O
import bench

dev = torch.device("cuda", 0)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 196
B, n, d, heads = 1024, 8, 512, 8
params, _, _, _ = O.make_pooler_case(d, 4, 1, 1, seed=3)
mod = AttentionalCrossPooler(d, d, heads).to(dev)
mod.load_state_dict(params)
g = torch.Generator(device=dev).manual_seed(L)
tokens = torch.randn(B, L, d, generator=g, device=dev).bfloat16().requires_grad_(True)
feats = torch.randn(n * B, d, generator=g, device=dev).bfloat16().requires_grad_(True)
w = torch.randn(n * B, d, generator=g, device=dev).bfloat16()
for _ in range(3):
    xm = crossmodal_features(mod, tokens, feats, B)
    xm.backward(w)
torch.cuda.synchronize()
shapes = bench.vitb16_cosmos_param_shapes()
student = [torch.randn(s, device=dev) * 0.02 for s in shapes]
teacher = [torch.randn(s, device=dev) * 0.02 for s in shapes]
for _ in range(3):
    ema_update_(student, teacher, 0.99)
torch.cuda.synchronize()
print("ok")

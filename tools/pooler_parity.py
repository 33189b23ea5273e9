"""Ad-hoc: how close is the pooler (bf16 tensor-core contractions) to the fp32 oracle evaluated on the SAME 16-bit-valued
inputs and weights?  Prints, per case and tensor, relative L2 error / gradient cosine / norm ratio - the numbers the
tolerances of tests/test_gpu_pooler.py are set from.    python tools/pooler_parity.py"""
import sys
import torch
sys.path.insert(0, ".")
from cosmos_b200.pooler import AttentionalCrossPooler, crossmodal_features  # noqa: E402
from oracle import cosmos_oracle as O  # noqa: E402


def cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


for (d, L, B, n, heads, seed) in [(512, 77, 16, 8, 8, 1), (512, 196, 16, 8, 8, 2), (512, 49, 5, 8, 8, 3), (768, 197, 4, 2, 12, 4)]:
    params, tokens, feats, w = O.make_pooler_case(d, L, B, n, seed)
    r16 = lambda t: t.bfloat16().float()
    # oracle: fp32 arithmetic on bf16-valued tokens / features / matrix weights (norm weights and biases stay fp32, as in the kernels)
    p32 = {k: (r16(v) if v.dim() == 2 else v.clone()).requires_grad_(True) for k, v in params.items()}
    t32, f32 = r16(tokens).requires_grad_(True), r16(feats).requires_grad_(True)
    ref = O.cosmos_crossmodal(f32, t32, p32, heads, B)
    (ref * w).sum().backward()
    mod = AttentionalCrossPooler(d, d, heads).cuda()
    mod.load_state_dict({k: (r16(v) if v.dim() == 2 else v) for k, v in params.items()})
    tok = tokens.bfloat16().cuda().requires_grad_(True)
    f = feats.bfloat16().cuda().requires_grad_(True)
    xm = crossmodal_features(mod, tok, f, B)
    (xm.float() * w.cuda()).sum().backward()
    print("case d=%d L=%d B=%d n=%d: out rel-L2 %.2e | d feats cos %.6f | d tokens cos %.6f norm %.4f" % (
        d, L, B, n, rel(xm.float(), ref), cos(f.grad.float(), f32.grad), cos(tok.grad.float(), t32.grad),
        float(tok.grad.float().norm().cpu() / t32.grad.norm())))
    for k, p in mod.named_parameters():
        g, gr = p.grad, p32[k].grad
        if k == "attn.in_proj_bias":
            sel = torch.cat([torch.arange(0, d), torch.arange(2 * d, 3 * d)])
            g, gr = g[sel], gr[sel]
        print("    %-24s cos %.6f  norm ratio %.4f" % (k, cos(g, gr), float(g.float().norm().cpu() / gr.norm())))

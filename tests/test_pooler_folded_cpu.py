"""CPU: the algebra and the stride bookkeeping of the pooler's folded attention (cosmos_b200/pooler.py:_folded_fwd/_folded_bwd)
against the oracle's written-out AttentionalCrossPooler (src/open_clip/transformer.py:210-230, src/open_clip/model.py:366-387),
with the kernel entry points replaced by the documented-contract emulation tests/emulation_pooler.py.  The kernels themselves
are covered by tests/test_gpu_pooler.py."""
import pytest
import torch

from tests import emulation_pooler
from oracle import cosmos_oracle as O


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()))


@pytest.mark.parametrize("d,L,B,n,heads,seed", [(64, 13, 5, 4, 4, 1), (128, 70, 3, 8, 8, 2), (96, 9, 2, 2, 12, 3), (64, 10, 3, 3, 4, 4)])
def test_folded_crossmodal_matches_oracle(monkeypatch, d, L, B, n, heads, seed):
    from cosmos_b200 import pooler
    emulation_pooler.install(monkeypatch)
    assert pooler._fold_ok(B, n, 1, B, heads, d)
    params, tokens, feats, w = O.make_pooler_case(d, L, B, n, seed)
    r16 = lambda t: t.bfloat16().float()
    p32 = {k: (r16(v) if v.dim() == 2 else v.clone()).requires_grad_(True) for k, v in params.items()}
    t32, f32 = r16(tokens).requires_grad_(True), r16(feats).requires_grad_(True)
    ref = O.cosmos_crossmodal(f32, t32, p32, heads, B)
    (ref * w).sum().backward()
    mod = pooler.AttentionalCrossPooler(d, d, heads)
    mod.load_state_dict({k: (r16(v) if v.dim() == 2 else v) for k, v in params.items()})
    tok = tokens.bfloat16().requires_grad_(True)
    f = feats.bfloat16().requires_grad_(True)
    xm = pooler.crossmodal_features(mod, tok, f, B)                   # crop-major query rows (model.py:373-376)
    (xm.float() * w).sum().backward()
    assert float((xm.detach().float() - ref.detach()).norm() / ref.detach().norm()) < 1e-2
    assert cosine(f.grad, f32.grad) >= 0.9995 and cosine(tok.grad, t32.grad) >= 0.9995
    for k, p in mod.named_parameters():
        g, gr = p.grad, p32[k].grad
        if k == "attn.in_proj_bias":          # the key third is exactly zero (softmax shift invariance): the folded route says so
            assert float(g[d:2 * d].abs().max()) == 0.0
            sel = torch.cat([torch.arange(0, d), torch.arange(2 * d, 3 * d)])
            g, gr = g[sel], gr[sel]
        assert cosine(g, gr) >= 0.999, (k, cosine(g, gr))
        assert abs(float(g.float().norm() / gr.norm()) - 1) < 2e-2, k


def test_folded_module_forward_set_major_matches_oracle(monkeypatch):
    """forward(x, q) (transformer.py:225-230): queries of a sample are consecutive rows - no reordering copies."""
    from cosmos_b200 import pooler
    emulation_pooler.install(monkeypatch)
    d, L, B, Lq, heads = 64, 11, 3, 6, 8
    params, tokens, _, _ = O.make_pooler_case(d, L, B, 1, 7)
    g = torch.Generator().manual_seed(5)
    q = torch.randn(B, Lq, d, generator=g)
    w = torch.randn(B, Lq, d, generator=g)
    r16 = lambda t: t.bfloat16().float()
    p32 = {k: (r16(v) if v.dim() == 2 else v.clone()).requires_grad_(True) for k, v in params.items()}
    t32, q32 = r16(tokens).requires_grad_(True), r16(q).requires_grad_(True)
    ref = O.cross_pool(t32, q32, p32, heads)
    (ref * w).sum().backward()
    mod = pooler.AttentionalCrossPooler(d, d, heads)
    mod.load_state_dict({k: (r16(v) if v.dim() == 2 else v) for k, v in params.items()})
    assert pooler._fold_ok(B, Lq, Lq, 1, heads, d) and pooler._row_order(B, Lq, Lq, 1) == "set"
    tok, qq = tokens.bfloat16().requires_grad_(True), q.bfloat16().requires_grad_(True)
    out = mod(tok, qq)
    (out.float() * w).sum().backward()
    assert float((out.detach().float() - ref.detach()).norm() / ref.detach().norm()) < 1e-2
    assert cosine(qq.grad, q32.grad) >= 0.9995 and cosine(tok.grad, t32.grad) >= 0.9995
    for k, p in mod.named_parameters():
        if k == "attn.in_proj_bias":
            continue
        assert cosine(p.grad, p32[k].grad) >= 0.999, k


def test_fold_decision():
    from cosmos_b200 import pooler
    assert pooler._fold_ok(1024, 8, 1, 1024, 8, 512)             # COSMOS: 8 crops x 8 heads = 64 score columns
    assert pooler._fold_ok(4, 2, 1, 4, 12, 768)
    assert not pooler._fold_ok(1024, 77, 77, 1, 12, 768)         # BASELINE config 4 literal, 77 queries: 924 columns, 1.1x the flops
    assert not pooler._fold_ok(1024, 197, 197, 1, 12, 768)       # 197 queries: 2364 columns -> key / value route with the GEMM core
    assert pooler._core_ok(1024, 197, 197, 1, 12, 768) and not pooler._core_ok(8, 4, 4, 1, 8, 96)      # head dim 12
    assert pooler._fold_ok(4, 20, 20, 1, 12, 768)
    assert not pooler._fold_ok(8, 4, 2, 3, 8, 512)               # an unknown row pattern


@pytest.mark.parametrize("d,L,B,Lq,heads", [(64, 11, 3, 6, 8), (96, 7, 2, 13, 12)])
def test_key_value_route_with_gemm_core_matches_oracle(monkeypatch, d, L, B, Lq, heads):
    """The attention core of the key / value route as batched GEMMs with samples as the outer and heads as the inner batch
    dimension (cosmos_b200/pooler.py:_core_fwd/_core_bwd), forced by switching the fold off."""
    from cosmos_b200 import pooler
    emulation_pooler.install(monkeypatch)
    monkeypatch.setattr(pooler, "_FOLD_MAX_COLS", 0)
    assert not pooler._fold_ok(B, Lq, Lq, 1, heads, d) and pooler._core_ok(B, Lq, Lq, 1, heads, d)
    params, tokens, _, _ = O.make_pooler_case(d, L, B, 1, 17)
    g = torch.Generator().manual_seed(6)
    q = torch.randn(B, Lq, d, generator=g)
    w = torch.randn(B, Lq, d, generator=g)
    r16 = lambda t: t.bfloat16().float()
    p32 = {k: (r16(v) if v.dim() == 2 else v.clone()).requires_grad_(True) for k, v in params.items()}
    t32, q32 = r16(tokens).requires_grad_(True), r16(q).requires_grad_(True)
    ref = O.cross_pool(t32, q32, p32, heads)
    (ref * w).sum().backward()
    mod = pooler.AttentionalCrossPooler(d, d, heads)
    mod.load_state_dict({k: (r16(v) if v.dim() == 2 else v) for k, v in params.items()})
    tok, qq = tokens.bfloat16().requires_grad_(True), q.bfloat16().requires_grad_(True)
    out = mod(tok, qq)
    (out.float() * w).sum().backward()
    assert float((out.detach().float() - ref.detach()).norm() / ref.detach().norm()) < 1e-2
    assert cosine(qq.grad, q32.grad) >= 0.9995 and cosine(tok.grad, t32.grad) >= 0.9995
    for k, p in mod.named_parameters():
        gg, gr = p.grad, p32[k].grad
        if k == "attn.in_proj_bias":
            sel = torch.cat([torch.arange(0, d), torch.arange(2 * d, 3 * d)])
            gg, gr = gg[sel], gr[sel]
        assert cosine(gg, gr) >= 0.999, (k, cosine(gg, gr))


def test_oracle_add_zero_attn_matches_torch_multihead_attention():
    """Pins the oracle's add_zero_attn branch (the reference module is LayerNorm x2 + nn.MultiheadAttention(add_zero_attn=...),
    src/open_clip/transformer.py:219-229) against torch's own implementation."""
    import torch.nn as nn
    d, L, B, Lq, heads = 32, 9, 3, 4, 4
    params, tokens, _, _ = O.make_pooler_case(d, L, B, 1, 23)
    g = torch.Generator().manual_seed(9)
    q = torch.randn(B, Lq, d, generator=g)
    attn = nn.MultiheadAttention(d, heads, kdim=d, vdim=d, add_zero_attn=True)
    ln_q, ln_k = nn.LayerNorm(d), nn.LayerNorm(d)
    with torch.no_grad():
        attn.in_proj_weight.copy_(params["attn.in_proj_weight"]); attn.in_proj_bias.copy_(params["attn.in_proj_bias"])
        attn.out_proj.weight.copy_(params["attn.out_proj.weight"]); attn.out_proj.bias.copy_(params["attn.out_proj.bias"])
        ln_q.weight.copy_(params["ln_q.weight"]); ln_q.bias.copy_(params["ln_q.bias"])
        ln_k.weight.copy_(params["ln_k.weight"]); ln_k.bias.copy_(params["ln_k.bias"])
        x = ln_k(tokens).permute(1, 0, 2)
        want = attn(ln_q(q).permute(1, 0, 2), x, x, need_weights=False)[0].permute(1, 0, 2)      # transformer.py:225-229
    got = O.cross_pool(tokens, q, params, heads, add_zero_attn=True)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4)
    assert not torch.allclose(O.cross_pool(tokens, q, params, heads), want, atol=1e-3)


@pytest.mark.parametrize("crop_major", [False, True])
def test_add_zero_attn_route_matches_oracle(monkeypatch, crop_major):
    """add_zero_attn=True: the key / value route with the batched-GEMM core and one more (zero) key in the softmax."""
    from cosmos_b200 import pooler
    emulation_pooler.install(monkeypatch)
    d, L, B, n, heads = 64, 7, 3, 4, 8
    params, tokens, feats, w = O.make_pooler_case(d, L, B, n, 31)
    r16 = lambda t: t.bfloat16().float()
    p32 = {k: (r16(v) if v.dim() == 2 else v.clone()).requires_grad_(True) for k, v in params.items()}
    mod = pooler.AttentionalCrossPooler(d, d, heads, add_zero_attn=True)
    mod.load_state_dict({k: (r16(v) if v.dim() == 2 else v) for k, v in params.items()})
    t32 = r16(tokens).requires_grad_(True)
    tok = tokens.bfloat16().requires_grad_(True)
    if crop_major:          # the call site of model.py:375-380 (one query per crop, crop-major rows, fused add + normalise)
        f32 = r16(feats).requires_grad_(True)
        rep = t32[:B].repeat(n, 1, 1)
        ref = torch.nn.functional.normalize(f32 + O.cross_pool(rep, f32.unsqueeze(1), p32, heads, add_zero_attn=True).squeeze(1), dim=-1)
        f = feats.bfloat16().requires_grad_(True)
        out = pooler.crossmodal_features(mod, tok, f, B)
        qg, qr = f, f32
    else:
        q = feats.view(n, B, d).transpose(0, 1).contiguous()
        q32 = r16(q).requires_grad_(True)
        ref = O.cross_pool(t32, q32, p32, heads, add_zero_attn=True)
        qq = q.bfloat16().requires_grad_(True)
        out = mod(tok, qq)
        w = w.view(n, B, d).transpose(0, 1).contiguous()
        qg, qr = qq, q32
    (ref * w).sum().backward()
    (out.float() * w).sum().backward()
    assert float((out.detach().float() - ref.detach()).norm() / ref.detach().norm()) < 1e-2
    assert cosine(qg.grad, qr.grad) >= 0.9995 and cosine(tok.grad, t32.grad) >= 0.9995
    for k, p in mod.named_parameters():
        assert cosine(p.grad, p32[k].grad) >= 0.999, (k, cosine(p.grad, p32[k].grad))      # incl. the key bias: not zero any more

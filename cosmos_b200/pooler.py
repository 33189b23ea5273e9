"""Cross-attention pooler of COSMOS on B200 kernels.

Drop-in for `AttentionalCrossPooler` (src/open_clip/transformer.py:210-230) and for the cross-modal
block of `CLIP.forward` that calls it (src/open_clip/model.py:366-387):

    txt_pooled = self.text_attn_cross_pool(txt_tokens.repeat(img_num, 1, 1), img_features.unsqueeze(1))
    img_crossmodal_features = F.normalize(img_features + txt_pooled.squeeze(), dim=-1)

`AttentionalCrossPooler` keeps the reference's parameter names (`attn.in_proj_weight`, `attn.in_proj_bias`,
`attn.out_proj.{weight,bias}`, `ln_q`, `ln_k`) so reference checkpoints load unchanged, and its
`forward(x, q)` has the reference's meaning.  `crossmodal_features(pooler, tokens, features, batch_size)`
replaces the three call-site lines: it takes the UN-repeated tokens of the first `batch_size` samples, so
LayerNorm and the key/value projection run once per unique sample instead of once per crop (the
reference executes them on 8x duplicated rows), and it fuses the residual add + L2 normalisation
(SURVEY.md §8(f) N1).

All contractions (key/value and query in-projections, out-projection, their input and weight
gradients) run on tcgen05 tensor cores (csrc/gemm.cu); LayerNorm, the few-queries attention core and
add+normalise are HBM-bound CUDA kernels (csrc/xpool.cu).  There is no PyTorch fallback.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _code(t: torch.Tensor) -> int:
    return _lib.torch_dtype_code(t.dtype)


def _gemm(a, b, out, M, N, K, lda, ldb, a_kmajor, b_kmajor, bias=None, splits=1, alpha=1.0):
    dev = a.device
    st = _lib.lib().cosmos_gemm(a.data_ptr(), b.data_ptr(), out.data_ptr(), bias.data_ptr() if bias is not None else None,
                                M, N, K, lda, ldb, out.stride(0), int(a_kmajor), int(b_kmajor), _code(a), _code(out), splits,
                                float(alpha), dev.index, _stream(dev))
    _lib.check(st, "gemm")
    return out


def _linear(x, w, bias, out_dtype):
    """x [M, K] @ w[N, K]^T + bias -> [M, N]"""
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty(M, N, dtype=out_dtype, device=x.device)
    return _gemm(x, w, out, M, N, K, x.stride(0), w.stride(0), True, True, bias=bias)


def _dgrad(g, w, out_dtype):
    """g [M, N] @ w[N, K] -> [M, K]   (input gradient of x @ w^T)"""
    M, N = g.shape
    K = w.shape[1]
    out = torch.empty(M, K, dtype=out_dtype, device=g.device)
    return _gemm(g, w, out, M, K, N, g.stride(0), w.stride(0), True, False)


def _wgrad(g, x, out=None):
    """g [R, N]^T @ x [R, K] -> [N, K] fp32   (weight gradient of x @ w^T); contraction over the rows, split for parallelism.
    out: a ZEROED [N, K] fp32 view to accumulate into (one memset for all of a step's parameter gradients)."""
    R, N = g.shape
    K = x.shape[1]
    if out is None:
        out = torch.zeros(N, K, dtype=torch.float32, device=g.device)
    bn = 256 if K >= 256 else 128                      # output tile of the GEMM kernel (csrc/gemm.cu)
    tiles = ((N + 127) // 128) * ((K + bn - 1) // bn)
    splits = max(1, min((R + 1023) // 1024, 148 // tiles if tiles <= 148 else 1))    # one wave of persistent CTAs
    return _gemm(g, x, out, N, K, R, g.stride(0), x.stride(0), False, False, splits=splits)


def _colsum(src, n, dst=None):
    if dst is None:
        dst = torch.zeros(n, dtype=torch.float32, device=src.device)
    st = _lib.lib().cosmos_colsum(src.data_ptr(), _code(src), dst.data_ptr(), src.shape[0], n, src.stride(0), src.device.index,
                                  _stream(src.device))
    _lib.check(st, "colsum")
    return dst


def _ln_fwd(x2d, w, b, out_dtype, eps=1e-5):
    rows, dim = x2d.shape
    dev = x2d.device
    y = torch.empty(rows, dim, dtype=out_dtype, device=dev)
    mean = torch.empty(rows, dtype=torch.float32, device=dev)
    rstd = torch.empty(rows, dtype=torch.float32, device=dev)
    st = _lib.lib().cosmos_layernorm_fwd(x2d.data_ptr(), _code(x2d), w.data_ptr(), b.data_ptr(), y.data_ptr(), _code(y),
                                         mean.data_ptr(), rstd.data_ptr(), rows, dim, float(eps), dev.index, _stream(dev))
    _lib.check(st, "layernorm_fwd")
    return y, mean, rstd


def _ln_bwd(dy, x2d, w, mean, rstd, dx, accumulate, dw=None, db=None):
    rows, dim = x2d.shape
    dev = x2d.device
    if dw is None:
        dw = torch.zeros(dim, dtype=torch.float32, device=dev)
        db = torch.zeros(dim, dtype=torch.float32, device=dev)
    st = _lib.lib().cosmos_layernorm_bwd(dy.data_ptr(), _code(dy), x2d.data_ptr(), _code(x2d), w.data_ptr(), mean.data_ptr(),
                                         rstd.data_ptr(), dx.data_ptr(), _code(dx), int(accumulate), dw.data_ptr(), db.data_ptr(),
                                         rows, dim, dev.index, _stream(dev))
    _lib.check(st, "layernorm_bwd")
    return dw, db


class _MapTokens(torch.autograd.Function):
    """y = x @ w^T + b on the tcgen05 GEMM, x [R, K] 16-bit, w [N, K], b [N] (or None)."""

    @staticmethod
    def forward(ctx, x, w, b):
        w16 = w.detach().to(x.dtype)
        y = _linear(x, w16, b.detach().float() if b is not None else None, x.dtype)
        ctx.save_for_backward(x, w16)
        ctx.has_bias = b is not None
        ctx.w_dtype = w.dtype
        return y

    @staticmethod
    def backward(ctx, g):
        x, w16 = ctx.saved_tensors
        g = g.contiguous()
        dx = _dgrad(g, w16, x.dtype) if ctx.needs_input_grad[0] else None
        dw = _wgrad(g, x).to(ctx.w_dtype) if ctx.needs_input_grad[1] else None
        db = _colsum(g, g.shape[1]).to(ctx.w_dtype) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db


def map_tokens(tokens: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], batch_size: int) -> torch.Tensor:
    """`text_token_mapping` / `image_token_mapping` (nn.Linear, src/open_clip/model.py:285-287, 306, 331) applied ONLY to
    the tokens the cross-attention poolers read: the first `batch_size` samples (model.py:370, 372 take `[:B]` of the mapped
    tokens; the reference maps all 2B global crops / 8B captions first).  tokens [n, L, K] 16-bit -> [batch_size, L, N]
    (SURVEY.md §8(f) N3).  Differentiable w.r.t. tokens, weight and bias; rows past `batch_size` get zero gradient, as in the
    reference, because they never reach the loss."""
    if tokens.dim() != 3 or tokens.shape[0] < batch_size or weight.dim() != 2 or weight.shape[1] != tokens.shape[2]:
        raise RuntimeError("cosmos_b200.pooler.map_tokens: tokens [n >= batch_size, L, K], weight [N, K]")
    if not tokens.is_cuda:
        raise RuntimeError("cosmos_b200.pooler.map_tokens: CUDA tensors only (no CPU fallback)")
    if tokens.dtype not in (torch.bfloat16, torch.float16):
        raise RuntimeError("cosmos_b200.pooler.map_tokens: 16-bit tokens only (run under autocast or cast the tokens)")
    x = tokens[:batch_size]
    L, K = x.shape[1], x.shape[2]
    y = _MapTokens.apply(x.reshape(batch_size * L, K), weight, bias)
    return y.view(batch_size, L, weight.shape[0])


class _CrossPool(torch.autograd.Function):
    """tokens [n_sets, L, C], queries [n_q, d]; query c of set s is row s * qs + c * qq.
    fuse_norm: return normalize(queries + pooled) (model.py:379-380) instead of pooled."""

    @staticmethod
    def forward(ctx, tokens, queries, lnq_w, lnq_b, lnk_w, lnk_b, in_w, in_b, out_w, out_b, heads, q_per_set, qs, qq, fuse_norm,
                eps_q=1e-5, eps_k=1e-5):
        for t in (tokens, queries, lnq_w, lnq_b, lnk_w, lnk_b, in_w, in_b, out_w, out_b):
            _lib.require_cuda(t, "pooler tensor")
        n_sets, L, C = tokens.shape
        n_q, d = queries.shape
        if C != d:
            raise RuntimeError("cosmos_b200.pooler: context_dim must equal d_model (the COSMOS configuration)")
        if n_q != n_sets * q_per_set:
            raise RuntimeError("cosmos_b200.pooler: number of queries does not match sets x queries-per-set")
        dev = tokens.device
        cd = torch.float16 if (tokens.dtype == torch.float16 or queries.dtype == torch.float16) else torch.bfloat16
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        tokens2d = tokens.detach().contiguous().view(n_sets * L, C)
        q_in = queries.detach().contiguous()
        w_in = in_w.detach().to(cd).contiguous()       # one cast of the packed projection; its query / key-value parts are row ranges
        w_q, w_kv = w_in[:d], w_in[d:]
        w_o = out_w.detach().to(cd).contiguous()
        b_in = f32(in_b)
        lnq_w32, lnq_b32, lnk_w32, lnk_b32 = f32(lnq_w), f32(lnq_b), f32(lnk_w), f32(lnk_b)

        xn, mean_k, rstd_k = _ln_fwd(tokens2d, lnk_w32, lnk_b32, cd, eps_k)        # once per unique token set
        kv = _linear(xn, w_kv, b_in[d:], cd)                                       # [n_sets*L, 2d]
        fn, mean_q, rstd_q = _ln_fwd(q_in, lnq_w32, lnq_b32, cd, eps_q)
        qp = _linear(fn, w_q, b_in[:d], cd)                                        # [n_q, d]
        o = torch.empty(n_q, d, dtype=cd, device=dev)
        lse = torch.empty(n_q, heads, dtype=torch.float32, device=dev)
        st = _lib.lib().cosmos_attn_core_fwd(qp.data_ptr(), kv.data_ptr(), o.data_ptr(), lse.data_ptr(), _code(qp), n_sets, L, d,
                                             heads, q_per_set, qs, qq, dev.index, _stream(dev))
        _lib.check(st, "attn_core_fwd")
        pooled = _linear(o, w_o, f32(out_b), torch.float32)                        # [n_q, d] fp32
        ctx.cfg = (n_sets, L, d, heads, q_per_set, qs, qq, fuse_norm, cd)
        ctx.dtypes = (tokens.dtype, queries.dtype, lnq_w.dtype, lnk_w.dtype, in_w.dtype, in_b.dtype, out_w.dtype, out_b.dtype)
        if fuse_norm:
            out = torch.empty_like(q_in)
            inv_norm = torch.empty(n_q, dtype=torch.float32, device=dev)
            st = _lib.lib().cosmos_addnorm_fwd(q_in.data_ptr(), _code(q_in), pooled.data_ptr(), out.data_ptr(),
                                               inv_norm.data_ptr(), n_q, d, dev.index, _stream(dev))
            _lib.check(st, "addnorm_fwd")
        else:
            out = pooled.to(queries.dtype)
            inv_norm = torch.empty(0, device=dev)
        ctx.save_for_backward(tokens2d, q_in, xn, mean_k, rstd_k, fn, mean_q, rstd_q, kv, qp, o, lse, w_q, w_kv, w_o,
                              lnq_w32, lnk_w32, out, inv_norm)
        return out

    @staticmethod
    def backward(ctx, g_out):
        (tokens2d, q_in, xn, mean_k, rstd_k, fn, mean_q, rstd_q, kv, qp, o, lse, w_q, w_kv, w_o, lnq_w32, lnk_w32, out,
         inv_norm) = ctx.saved_tensors
        n_sets, L, d, heads, q_per_set, qs, qq, fuse_norm, cd = ctx.cfg
        dt_tok, dt_q, dt_lnq, dt_lnk, dt_inw, dt_inb, dt_ow, dt_ob = ctx.dtypes
        dev = tokens2d.device
        n_q = q_in.shape[0]
        lib = _lib.lib()
        g_out = g_out.contiguous()
        # every parameter gradient of the step accumulates (split-K GEMMs, column sums, LayerNorm partials use fp32 atomics)
        # into ONE zeroed workspace: a single memset instead of ten, and the packed in-projection gradient needs no torch.cat
        ws = torch.zeros(3 * d * d + d * d + 3 * d + d + 4 * d, dtype=torch.float32, device=dev)
        g_in_w, g_wo = ws[:3 * d * d].view(3 * d, d), ws[3 * d * d:4 * d * d].view(d, d)
        vec = ws[4 * d * d:]
        g_in_b, g_bo = vec[:3 * d], vec[3 * d:4 * d]
        g_lnq_w, g_lnq_b, g_lnk_w, g_lnk_b = vec[4 * d:5 * d], vec[5 * d:6 * d], vec[6 * d:7 * d], vec[7 * d:8 * d]
        if fuse_norm:
            g_z32 = torch.empty(n_q, d, dtype=torch.float32, device=dev)
            g_p = torch.empty(n_q, d, dtype=cd, device=dev)
            st = lib.cosmos_addnorm_bwd(g_out.to(out.dtype).data_ptr(), out.data_ptr(), _code(out), inv_norm.data_ptr(),
                                        g_z32.data_ptr(), g_p.data_ptr(), _code(g_p), n_q, d, dev.index, _stream(dev))
            _lib.check(st, "addnorm_bwd")
            g_queries = g_z32                                  # residual branch; the LN_q path is accumulated below
            _colsum(g_z32, d, g_bo)
        else:
            g_p = g_out.to(cd)
            g_queries = torch.zeros(n_q, d, dtype=torch.float32, device=dev)
            _colsum(g_out, d, g_bo)
        # out-projection
        _wgrad(g_p, o, g_wo)                                   # [d, d]
        g_o = _dgrad(g_p, w_o, cd)                             # [n_q, d]
        # attention core
        dq = torch.empty(n_q, d, dtype=cd, device=dev)
        dkv = torch.empty_like(kv)
        st = lib.cosmos_attn_core_bwd(qp.data_ptr(), kv.data_ptr(), g_o.data_ptr(), lse.data_ptr(), dq.data_ptr(), dkv.data_ptr(),
                                      _code(qp), n_sets, L, d, heads, q_per_set, qs, qq, dev.index, _stream(dev))
        _lib.check(st, "attn_core_bwd")
        # query in-projection + LayerNorm_q
        _wgrad(dq, fn, g_in_w[:d])
        _colsum(dq, d, g_in_b[:d])
        g_fn = _dgrad(dq, w_q, cd)
        _ln_bwd(g_fn, q_in, lnq_w32, mean_q, rstd_q, g_queries, True, g_lnq_w, g_lnq_b)
        # key/value in-projection + LayerNorm_k (once per unique token set)
        _wgrad(dkv, xn, g_in_w[d:])
        _colsum(dkv, 2 * d, g_in_b[d:])
        g_xn = _dgrad(dkv, w_kv, cd)
        g_tokens = torch.empty(n_sets * L, d, dtype=dt_tok, device=dev)
        _ln_bwd(g_xn, tokens2d, lnk_w32, mean_k, rstd_k, g_tokens, False, g_lnk_w, g_lnk_b)
        return (g_tokens.view(n_sets, L, d), g_queries.to(dt_q), g_lnq_w.to(dt_lnq), g_lnq_b.to(dt_lnq), g_lnk_w.to(dt_lnk),
                g_lnk_b.to(dt_lnk), g_in_w.to(dt_inw), g_in_b.to(dt_inb), g_wo.to(dt_ow), g_bo.to(dt_ob), None, None, None, None,
                None, None, None)


class AttentionalCrossPooler(nn.Module):
    """Same constructor, parameters and forward contract as the reference module
    (src/open_clip/transformer.py:210-230)."""

    def __init__(self, d_model: int, context_dim: int, n_head: int = 8, norm_layer=nn.LayerNorm, add_zero_attn: bool = False):
        super().__init__()
        if add_zero_attn:
            raise NotImplementedError("cosmos_b200.pooler: add_zero_attn=True is not used by the COSMOS recipes and is unsupported")
        if context_dim != d_model:
            raise NotImplementedError("cosmos_b200.pooler: context_dim must equal d_model (COSMOS maps tokens to embed_dim first)")
        # parameter containers with the reference's names; their own forward() is never called
        self.attn = nn.MultiheadAttention(d_model, n_head, kdim=context_dim, vdim=context_dim, add_zero_attn=False)
        self.ln_q = norm_layer(d_model)
        self.ln_k = norm_layer(context_dim)
        self.n_head = n_head

    def _params(self):
        return (self.ln_q.weight, self.ln_q.bias, self.ln_k.weight, self.ln_k.bias, self.attn.in_proj_weight,
                self.attn.in_proj_bias, self.attn.out_proj.weight, self.attn.out_proj.bias)

    def _eps(self):
        """The eps of the two norm layers (a custom norm_layer may carry its own; nn.LayerNorm: 1e-5)."""
        return (float(getattr(self.ln_q, "eps", 1e-5)), float(getattr(self.ln_k, "eps", 1e-5)))

    def forward(self, x: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
        """x: [N, L, C] keys/values, q: [N, Lq, d] queries -> [N, Lq, d] (transformer.py:225-230)."""
        N, Lq, d = q.shape
        out = _CrossPool.apply(x, q.reshape(N * Lq, d), *self._params(), self.n_head, Lq, Lq, 1, False, *self._eps())
        return out.view(N, Lq, d)


def crossmodal_features(pooler: AttentionalCrossPooler, tokens: torch.Tensor, features: torch.Tensor, batch_size: int) -> torch.Tensor:
    """normalize(features + pooler(tokens[:batch_size].repeat(n, 1, 1), features[:, None]).squeeze(), dim=-1)
    for features [n * batch_size, d] laid out crop-major (model.py:366-384), without materialising the repeat."""
    n = features.shape[0] // batch_size
    if n * batch_size != features.shape[0]:
        raise RuntimeError("cosmos_b200.pooler: features rows must be a multiple of batch_size")
    return _CrossPool.apply(tokens[:batch_size], features, *pooler._params(), pooler.n_head, n, 1, batch_size, True, *pooler._eps())

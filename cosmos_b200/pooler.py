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

All contractions run on tcgen05 tensor cores (csrc/gemm.cu).  With the few queries per sample COSMOS uses (one per crop:
8 x 8 heads = 64 score columns) the attention is FOLDED: the key projection moves into the queries and the value
projection behind the pooling (`_folded_fwd`), so scores, pooling and all their gradients are batched GEMMs over the
samples and no key / value tensor exists.  Where folding would not at least halve the flops (the module's general
forward(x, q) with many queries per sample) the key / value projection GEMM stays and the attention core itself runs as
batched GEMMs with samples as the outer and heads as the inner batch dimension (`_core_fwd`); the CUDA-core attention
kernel is only left for head dims that are not a multiple of 8.  LayerNorm, the column softmax and
add+normalise are HBM-bound CUDA kernels (csrc/xpool.cu).  There is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os
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


def _bgemm(a, b, out, M, N, K, lda, ldb, ldd, batch, sa, sb, sd, a_kmajor, b_kmajor, bias=None, sbias=0, splits=1, accumulate=False,
           alpha=1.0, second=None, inner=None):
    """`batch` problems D_t = alpha * opA_t opB_t^T (+ bias_t) through cosmos_gemm_ex; a, b, out, bias are tensors (views) whose
    first element is the first problem's, every stride is in elements.  a_kmajor: A_t stored [M, K] (row stride lda), else
    [K, M]; likewise B_t [N, K] / [K, N].  second = (a2, b2, K2, lda2, ldb2, sa2, sb2): D_t = alpha * (A_t B_t^T + A2_t B2_t^T).
    inner = (batch_in, sa_in, sb_in, sd_in): an inner batch dimension, problem (t1, t2) at a + t1 * sa + t2 * sa_in, ..."""
    dev = a.device
    g = _lib.GemmDesc()
    g.a, g.b, g.d = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.bias = bias.data_ptr() if bias is not None else None
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldd = lda, ldb, ldd
    g.batch, g.batch_in = batch, 1
    g.stride_a, g.stride_b, g.stride_d, g.stride_bias = sa, sb, sd, sbias
    if second is not None:
        a2, b2, g.K2, g.lda2, g.ldb2, g.stride_a2, g.stride_b2 = second
        g.a2, g.b2 = a2.data_ptr(), b2.data_ptr()
    if inner is not None:
        g.batch_in, g.stride_a_in, g.stride_b_in, g.stride_d_in = inner
    g.a_kmajor, g.b_kmajor = int(a_kmajor), int(b_kmajor)
    g.in_dtype, g.out_dtype = _code(a), _code(out)
    g.splits, g.accumulate, g.alpha = splits, int(accumulate), float(alpha)
    st = _lib.lib().cosmos_gemm_ex(C.byref(g), dev.index, _stream(dev))
    _lib.check(st, "gemm_ex")
    return out


def _colsoftmax_fwd(scores, p_out, n_sets, L, n_cols, lds, ldp, zero_key=False):
    """softmax over the L keys of every column of scores [n_sets, L, n_cols] fp32 (row stride lds) -> p_out (16-bit view, row
    stride ldp).  zero_key: one more key with score 0 and value 0 (add_zero_attn)."""
    dev = scores.device
    st = _lib.lib().cosmos_colsoftmax_fwd(scores.data_ptr(), L * lds, lds, p_out.data_ptr(), L * ldp, ldp, _code(p_out), n_sets, L,
                                          n_cols, int(zero_key), dev.index, _stream(dev))
    _lib.check(st, "colsoftmax_fwd")


def _colsoftmax_bwd(p_in, d_p, ds_out, n_sets, L, n_cols, lds, ldp):
    """ds = p * (dp - sum_l p dp) per column; p_in / ds_out 16-bit views with row stride ldp, d_p [n_sets, L, n_cols] fp32 with
    row stride lds."""
    dev = d_p.device
    st = _lib.lib().cosmos_colsoftmax_bwd(p_in.data_ptr(), L * ldp, ldp, d_p.data_ptr(), L * lds, lds, ds_out.data_ptr(), L * ldp,
                                          ldp, _code(p_in), n_sets, L, n_cols, dev.index, _stream(dev))
    _lib.check(st, "colsoftmax_bwd")


def _wgrad_splits(R, tiles):
    return max(1, min((R + 1023) // 1024, 148 // tiles if tiles <= 148 else 1))    # one wave of persistent CTAs


def _row_order(n_sets, q_per_set, qs, qq):
    """How query c of set s (row s * qs + c * qq) relates to set-major order: 'set' (already), 'crop' (crop-major: the layout of
    model.py:373-376, features of crop c of all samples together) or None (anything else: no folded attention)."""
    if qq == 1 and qs == q_per_set:
        return "set"
    if qs == 1 and qq == n_sets:
        return "crop"
    return None


def _to_set_major(t, order, n_sets, q_per_set):
    if order == "set":
        return t
    return t.view(q_per_set, n_sets, t.shape[1]).transpose(0, 1).reshape(n_sets * q_per_set, t.shape[1])


def _from_set_major(t, order, n_sets, q_per_set):
    if order == "set":
        return t
    return t.view(n_sets, q_per_set, t.shape[1]).transpose(0, 1).reshape(n_sets * q_per_set, t.shape[1])


# Folded attention (few queries per token set): COSMOS_B200_POOLER=unfolded keeps the key / value projection route (diagnostics)
# COSMOS_B200_POOLER_FOLD_MAX: largest number of score columns (queries x heads) per sample that is folded (diagnostics;
# default: the flop comparison in _fold_ok)
_FOLD_MAX_COLS = 0 if os.environ.get("COSMOS_B200_POOLER", "") == "unfolded" else int(os.environ.get("COSMOS_B200_POOLER_FOLD_MAX", "-1"))


def _fold_ok(n_sets, q_per_set, qs, qq, heads, d):
    hd = d // heads
    n_cols = q_per_set * heads
    # Folded, the two large products cost 4 L n_cols d flop per sample against 4 L d (d + queries) on the key / value route
    # (both all on tensor cores: the attention core of that route is batched GEMMs too, _core_fwd).  Measured at batch 1024:
    # 8 queries x 8 heads, width 512 (64 columns, 0.12x the flops): 1.50 vs 2.68 ms; 77 queries x 12 heads, width 768 (924
    # columns, 1.1x the flops, 1.5 GB of folded queries): 9.0 vs 7.2 ms - so: fold up to half the flops.
    limit = _FOLD_MAX_COLS if _FOLD_MAX_COLS >= 0 else max(128, (d + q_per_set) // 2)
    return (n_cols <= limit and hd % 8 == 0 and hd * heads == d and _row_order(n_sets, q_per_set, qs, qq) is not None)


def _folded_fwd(xn, qp, w_kv, b_in, n_sets, L, d, heads, q_per_set, order, cd):
    """Attention of q_per_set queries per token set WITHOUT key / value tensors.  With k_l = W_k x_l + b_k the score of query q_h
    against key l is q_h . k_l = (W_k,h^T q_h) . x_l + const, and sum_l p_l v_l = W_v,h (sum_l p_l x_l) + b_v,h, so per set
        Q~ = kappa * W_k,h^T q_h   [n_cols = queries x heads, d]      S = xn Q~^T  [L, n_cols]
        P  = softmax over l        Z = P^T xn  [n_cols, d]            o_h = W_v,h Z_h + b_v,h
    - batched tcgen05 GEMMs over sets (S, Z) and over heads (Q~, o), 8x fewer flops than the [L, d] x [d, 2d] projection at 8
    queries, and the tokens are the only large tensor read.  -> o [n_q, d] in set-major order, and what backward needs."""
    dev = xn.device
    hd, n_cols, n_q = d // heads, q_per_set * heads, n_sets * q_per_set
    w_k, w_v = w_kv[:d], w_kv[d:]
    qp_sm = _to_set_major(qp, order, n_sets, q_per_set).contiguous()
    qt = torch.empty(n_q, heads, d, dtype=cd, device=dev)                       # Q~, rows (set, query), then head
    _bgemm(qp_sm, w_k, qt, n_q, d, hd, d, d, heads * d, heads, hd, hd * d, d, True, False, alpha=hd ** -0.5)
    ncp = (n_cols + 7) // 8 * 8                                                 # row pitch of the score blocks (16-byte rows and halves)
    scores = torch.empty(n_sets, L, ncp, dtype=torch.float32, device=dev)
    _bgemm(xn, qt, scores, L, n_cols, d, d, d, ncp, n_sets, L * d, n_cols * d, L * ncp, True, True)
    pd = torch.empty(n_sets, L, 2 * ncp, dtype=cd, device=dev)                  # [P | dS]: the second half is filled by backward
    _colsoftmax_fwd(scores, pd, n_sets, L, n_cols, ncp, 2 * ncp)
    z = torch.empty(n_q, heads, d, dtype=cd, device=dev)
    _bgemm(pd, xn, z, n_cols, d, L, 2 * ncp, d, d, n_sets, L * 2 * ncp, L * d, n_cols * d, False, False)
    o_sm = torch.empty(n_q, d, dtype=cd, device=dev)
    _bgemm(z, w_v, o_sm, n_q, hd, d, heads * d, d, d, heads, d, hd * d, hd, True, True, bias=b_in[2 * d:], sbias=hd)
    return o_sm, qp_sm, qt, pd, z


def _folded_bwd(g_o_sm, xn, qp_sm, qt, pd, z, w_kv, g_in_w, g_in_b, n_sets, L, d, heads, q_per_set, cd):
    """-> (g_xn [n_sets * L, d], dq_sm [n_q, d] set-major); accumulates dW_k, dW_v into g_in_w[d:3d] and db_v into g_in_b[2d:]
    (db_k is exactly zero: a shift of every key's score by the same amount leaves the softmax unchanged)."""
    dev = xn.device
    hd, n_cols, n_q = d // heads, q_per_set * heads, n_sets * q_per_set
    w_k, w_v = w_kv[:d], w_kv[d:]
    kappa = hd ** -0.5
    splits = _wgrad_splits(n_q, heads * ((d + 255) // 256))
    # o_h = W_v,h Z_h + b_v,h
    _bgemm(g_o_sm, z, g_in_w[2 * d:], hd, d, n_q, d, heads * d, d, heads, hd, d, hd * d, False, False, splits=splits)
    _colsum(g_o_sm, d, g_in_b[2 * d:])
    dz = torch.empty(n_q, heads, d, dtype=cd, device=dev)
    _bgemm(g_o_sm, w_v, dz, n_q, d, hd, d, d, heads * d, heads, hd, hd * d, d, True, False)
    # Z = P^T xn
    ncp = pd.shape[2] // 2
    d_p = torch.empty(n_sets, L, ncp, dtype=torch.float32, device=dev)
    _bgemm(xn, dz, d_p, L, n_cols, d, d, d, ncp, n_sets, L * d, n_cols * d, L * ncp, True, True)
    ds = pd[:, :, ncp:]
    _colsoftmax_bwd(pd, d_p, ds, n_sets, L, n_cols, ncp, 2 * ncp)
    # d xn = P dZ + dS Q~   (two operand pairs, one pass over the output)
    g_xn = torch.empty(n_sets * L, d, dtype=cd, device=dev)
    _bgemm(pd, dz, g_xn, L, d, n_cols, 2 * ncp, d, d, n_sets, L * 2 * ncp, n_cols * d, L * d, True, False,
           second=(ds, qt, n_cols, 2 * ncp, d, L * 2 * ncp, n_cols * d))
    # S = xn Q~^T,  Q~ = kappa W_k,h^T q_h
    dqt = torch.empty(n_q, heads, d, dtype=cd, device=dev)
    _bgemm(ds, xn, dqt, n_cols, d, L, 2 * ncp, d, d, n_sets, L * 2 * ncp, L * d, n_cols * d, False, False)
    dq_sm = torch.empty(n_q, d, dtype=cd, device=dev)
    _bgemm(dqt, w_k, dq_sm, n_q, hd, d, heads * d, d, d, heads, d, hd * d, hd, True, True, alpha=kappa)
    _bgemm(qp_sm, dqt, g_in_w[d:2 * d], hd, d, n_q, d, heads * d, d, heads, hd, d, hd * d, False, False, splits=splits, alpha=kappa)
    return g_xn, dq_sm


def _core_ok(n_sets, q_per_set, qs, qq, heads, d):
    """The attention core of the key / value route as batched GEMMs (any number of queries)."""
    hd = d // heads
    return (os.environ.get("COSMOS_B200_POOLER_CORE", "") != "cuda_cores" and hd % 8 == 0 and hd * heads == d
            and _row_order(n_sets, q_per_set, qs, qq) is not None)


def _core_fwd(qp_sm, kv, n_sets, L, d, heads, q_per_set, cd, zero_key=False):
    """softmax(q_h k_h^T / sqrt(hd)) v_h per (sample, head) - F.multi_head_attention_forward's core - on tensor cores: per
    (sample, head) problem  S^T = kappa K_h Q_h^T  [L, queries]  (keys-major, so that the softmax over the keys is the column
    softmax of the folded route),  O_h = P^T V_h  [queries, hd].  Samples are the outer, heads the inner batch dimension of the
    operand tensor maps (head h of a [rows, d] matrix is the column block h * hd).  qp_sm [n_q, d] set-major, kv [n_sets * L, 2d]
    -> o_sm [n_q, d] and pd = [P | room for dS] [n_sets * heads, L, 2 * pitch]."""
    dev = kv.device
    hd, q = d // heads, q_per_set
    qpad = (q + 7) // 8 * 8
    scores = torch.empty(n_sets * heads, L, qpad, dtype=torch.float32, device=dev)
    _bgemm(kv, qp_sm, scores, L, q, hd, 2 * d, d, qpad, n_sets, L * 2 * d, q * d, heads * L * qpad, True, True, alpha=hd ** -0.5,
           inner=(heads, hd, hd, L * qpad))
    pd = torch.empty(n_sets * heads, L, 2 * qpad, dtype=cd, device=dev)
    _colsoftmax_fwd(scores, pd, n_sets * heads, L, q, qpad, 2 * qpad, zero_key)
    o_sm = torch.empty(n_sets * q, d, dtype=cd, device=dev)
    _bgemm(pd, kv[:, d:], o_sm, q, hd, L, 2 * qpad, 2 * d, d, n_sets, heads * L * 2 * qpad, L * 2 * d, q * d, False, False,
           inner=(heads, L * 2 * qpad, hd, hd))
    return o_sm, pd


def _core_bwd(g_o_sm, qp_sm, kv, pd, n_sets, L, d, heads, q_per_set, cd):
    """-> (dq_sm [n_q, d] set-major, dkv [n_sets * L, 2d])"""
    dev = kv.device
    hd, q = d // heads, q_per_set
    qpad = pd.shape[2] // 2
    kappa = hd ** -0.5
    v = kv[:, d:]
    d_p = torch.empty(n_sets * heads, L, qpad, dtype=torch.float32, device=dev)          # dP^T = V_h dO_h^T
    _bgemm(v, g_o_sm, d_p, L, q, hd, 2 * d, d, qpad, n_sets, L * 2 * d, q * d, heads * L * qpad, True, True,
           inner=(heads, hd, hd, L * qpad))
    ds = pd[:, :, qpad:]
    _colsoftmax_bwd(pd, d_p, ds, n_sets * heads, L, q, qpad, 2 * qpad)
    dkv = torch.empty_like(kv)
    # dV_h = P dO_h,  dK_h = kappa dS Q_h   ([L, queries] x [queries, hd]: the second operands are read as stored, [K, N])
    _bgemm(pd, g_o_sm, dkv[:, d:], L, hd, q, 2 * qpad, d, 2 * d, n_sets, heads * L * 2 * qpad, q * d, L * 2 * d, True, False,
           inner=(heads, L * 2 * qpad, hd, hd))
    _bgemm(ds, qp_sm, dkv, L, hd, q, 2 * qpad, d, 2 * d, n_sets, heads * L * 2 * qpad, q * d, L * 2 * d, True, False, alpha=kappa,
           inner=(heads, L * 2 * qpad, hd, hd))
    # dQ_h = kappa dS^T K_h
    dq_sm = torch.empty(n_sets * q, d, dtype=cd, device=dev)
    _bgemm(ds, kv, dq_sm, q, hd, L, 2 * qpad, 2 * d, d, n_sets, heads * L * 2 * qpad, L * 2 * d, q * d, False, False, alpha=kappa,
           inner=(heads, L * 2 * qpad, hd, hd))
    return dq_sm, dkv


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


def _addnorm_fwd(q_in, pooled):
    """normalize(q_in + pooled) row-wise (model.py:379-380) -> (out in q_in's dtype, 1 / norm fp32)"""
    n_q, d = q_in.shape
    dev = q_in.device
    out = torch.empty_like(q_in)
    inv_norm = torch.empty(n_q, dtype=torch.float32, device=dev)
    st = _lib.lib().cosmos_addnorm_fwd(q_in.data_ptr(), _code(q_in), pooled.data_ptr(), out.data_ptr(), inv_norm.data_ptr(), n_q, d,
                                       dev.index, _stream(dev))
    _lib.check(st, "addnorm_fwd")
    return out, inv_norm


def _addnorm_bwd(g_out, out, inv_norm, cd):
    """-> (gradient of the sum in fp32, the same in the compute dtype)"""
    n_q, d = out.shape
    dev = out.device
    g_z32 = torch.empty(n_q, d, dtype=torch.float32, device=dev)
    g_p = torch.empty(n_q, d, dtype=cd, device=dev)
    g_out = g_out.to(out.dtype)
    st = _lib.lib().cosmos_addnorm_bwd(g_out.data_ptr(), out.data_ptr(), _code(out), inv_norm.data_ptr(), g_z32.data_ptr(),
                                       g_p.data_ptr(), _code(g_p), n_q, d, dev.index, _stream(dev))
    _lib.check(st, "addnorm_bwd")
    return g_z32, g_p


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
                eps_q=1e-5, eps_k=1e-5, zero_key=False):
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
        fn, mean_q, rstd_q = _ln_fwd(q_in, lnq_w32, lnq_b32, cd, eps_q)
        qp = _linear(fn, w_q, b_in[:d], cd)                                        # [n_q, d]
        # add_zero_attn: the extra all-zero key is appended AFTER the key projection, so its score relative to the real keys
        # involves the key bias the fold drops - it takes the key / value route with the batched-GEMM core
        folded = _fold_ok(n_sets, q_per_set, qs, qq, heads, d) and not zero_key
        if zero_key and not _core_ok(n_sets, q_per_set, qs, qq, heads, d):
            raise NotImplementedError("cosmos_b200.pooler: add_zero_attn needs head dims that are a multiple of 8")
        if folded:
            order = _row_order(n_sets, q_per_set, qs, qq)
            o_sm, qp_sm, qt, pd, z = _folded_fwd(xn, qp, w_kv, b_in, n_sets, L, d, heads, q_per_set, order, cd)
            o = _from_set_major(o_sm, order, n_sets, q_per_set).contiguous()
            kv = lse = torch.empty(0, device=dev)
            fold_saved = (qp_sm, qt, pd, z)
        else:
            kv = _linear(xn, w_kv, b_in[d:], cd)                                   # [n_sets*L, 2d]
            empty = torch.empty(0, device=dev)
            if _core_ok(n_sets, q_per_set, qs, qq, heads, d):                      # attention core as batched GEMMs
                order = _row_order(n_sets, q_per_set, qs, qq)
                qp_sm = _to_set_major(qp, order, n_sets, q_per_set).contiguous()
                o_sm, pd = _core_fwd(qp_sm, kv, n_sets, L, d, heads, q_per_set, cd, zero_key)
                o = _from_set_major(o_sm, order, n_sets, q_per_set).contiguous()
                lse = empty
                fold_saved = (qp_sm, empty, pd, empty)
            else:                                                                  # CUDA-core attention kernel (head dims not a multiple of 8, unknown row patterns)
                order = None
                o = torch.empty(n_q, d, dtype=cd, device=dev)
                lse = torch.empty(n_q, heads, dtype=torch.float32, device=dev)
                st = _lib.lib().cosmos_attn_core_fwd(qp.data_ptr(), kv.data_ptr(), o.data_ptr(), lse.data_ptr(), _code(qp), n_sets, L,
                                                     d, heads, q_per_set, qs, qq, dev.index, _stream(dev))
                _lib.check(st, "attn_core_fwd")
                fold_saved = (empty, empty, empty, empty)
        ctx.fold = (folded, order)
        pooled = _linear(o, w_o, f32(out_b), torch.float32)                        # [n_q, d] fp32
        ctx.cfg = (n_sets, L, d, heads, q_per_set, qs, qq, fuse_norm, cd)
        ctx.dtypes = (tokens.dtype, queries.dtype, lnq_w.dtype, lnk_w.dtype, in_w.dtype, in_b.dtype, out_w.dtype, out_b.dtype)
        if fuse_norm:
            out, inv_norm = _addnorm_fwd(q_in, pooled)
        else:
            out = pooled.to(queries.dtype)
            inv_norm = torch.empty(0, device=dev)
        ctx.save_for_backward(tokens2d, q_in, xn, mean_k, rstd_k, fn, mean_q, rstd_q, kv, qp, o, lse, w_q, w_kv, w_o,
                              lnq_w32, lnk_w32, out, inv_norm, *fold_saved)
        return out

    @staticmethod
    def backward(ctx, g_out):
        (tokens2d, q_in, xn, mean_k, rstd_k, fn, mean_q, rstd_q, kv, qp, o, lse, w_q, w_kv, w_o, lnq_w32, lnk_w32, out,
         inv_norm, qp_sm, qt, pd, z) = ctx.saved_tensors
        folded, order = ctx.fold
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
            g_z32, g_p = _addnorm_bwd(g_out, out, inv_norm, cd)
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
        if folded:
            g_o_sm = _to_set_major(g_o, order, n_sets, q_per_set).contiguous()
            g_xn, dq_sm = _folded_bwd(g_o_sm, xn, qp_sm, qt, pd, z, w_kv, g_in_w, g_in_b, n_sets, L, d, heads, q_per_set, cd)
            dq = _from_set_major(dq_sm, order, n_sets, q_per_set).contiguous()
        elif order is not None:
            g_o_sm = _to_set_major(g_o, order, n_sets, q_per_set).contiguous()
            dq_sm, dkv = _core_bwd(g_o_sm, qp_sm, kv, pd, n_sets, L, d, heads, q_per_set, cd)
            dq = _from_set_major(dq_sm, order, n_sets, q_per_set).contiguous()
        else:
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
        if not folded:
            _wgrad(dkv, xn, g_in_w[d:])
            _colsum(dkv, 2 * d, g_in_b[d:])
            g_xn = _dgrad(dkv, w_kv, cd)
        g_tokens = torch.empty(n_sets * L, d, dtype=dt_tok, device=dev)
        _ln_bwd(g_xn, tokens2d, lnk_w32, mean_k, rstd_k, g_tokens, False, g_lnk_w, g_lnk_b)
        return (g_tokens.view(n_sets, L, d), g_queries.to(dt_q), g_lnq_w.to(dt_lnq), g_lnq_b.to(dt_lnq), g_lnk_w.to(dt_lnk),
                g_lnk_b.to(dt_lnk), g_in_w.to(dt_inw), g_in_b.to(dt_inb), g_wo.to(dt_ow), g_bo.to(dt_ob), None, None, None, None,
                None, None, None, None)


class AttentionalCrossPooler(nn.Module):
    """Same constructor, parameters and forward contract as the reference module
    (src/open_clip/transformer.py:210-230)."""

    def __init__(self, d_model: int, context_dim: int, n_head: int = 8, norm_layer=nn.LayerNorm, add_zero_attn: bool = False):
        super().__init__()
        if context_dim != d_model:
            raise NotImplementedError("cosmos_b200.pooler: context_dim must equal d_model (COSMOS maps tokens to embed_dim first)")
        # parameter containers with the reference's names; their own forward() is never called
        self.attn = nn.MultiheadAttention(d_model, n_head, kdim=context_dim, vdim=context_dim, add_zero_attn=add_zero_attn)
        self.add_zero_attn = bool(add_zero_attn)
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
        out = _CrossPool.apply(x, q.reshape(N * Lq, d), *self._params(), self.n_head, Lq, Lq, 1, False, *self._eps(), self.add_zero_attn)
        return out.view(N, Lq, d)


def crossmodal_features(pooler: AttentionalCrossPooler, tokens: torch.Tensor, features: torch.Tensor, batch_size: int) -> torch.Tensor:
    """normalize(features + pooler(tokens[:batch_size].repeat(n, 1, 1), features[:, None]).squeeze(), dim=-1)
    for features [n * batch_size, d] laid out crop-major (model.py:366-384), without materialising the repeat."""
    n = features.shape[0] // batch_size
    if n * batch_size != features.shape[0]:
        raise RuntimeError("cosmos_b200.pooler: features rows must be a multiple of batch_size")
    return _CrossPool.apply(tokens[:batch_size], features, *pooler._params(), pooler.n_head, n, 1, batch_size, True, *pooler._eps(),
                            pooler.add_zero_attn)

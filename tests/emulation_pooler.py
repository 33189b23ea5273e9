"""Test-only CPU emulation of the pooler's kernel entry points (cosmos_b200.pooler._gemm, _bgemm, LayerNorm, column softmax,
add + normalise), restating what include/cosmos_b200.h documents for them in torch.  The batched GEMM is emulated from the
same (pointer, strides) description the kernel gets - views built with as_strided on the operand's storage - so that the
stride bookkeeping of the folded attention (per-head column blocks as a batch, [P | dS] halves, MN-major operands) is
exercised exactly as the tensor maps would read it.  Never imported by the package: the product path has no CPU route."""
import torch


def _mat(t, rows, cols, ld, batch, bs, kmajor, batch_in=1, bs_in=0):
    """the [batch, batch_in, rows, cols] matrices an operand pointer describes: stored [rows, cols] (kmajor) or [cols, rows]"""
    if kmajor:
        return t.as_strided((batch, batch_in, rows, cols), (bs, bs_in, ld, 1))
    return t.as_strided((batch, batch_in, cols, rows), (bs, bs_in, ld, 1)).transpose(2, 3)


def emu_bgemm(a, b, out, M, N, K, lda, ldb, ldd, batch, sa, sb, sd, a_kmajor, b_kmajor, bias=None, sbias=0, splits=1, accumulate=False,
              alpha=1.0, second=None, inner=None):
    bi, sa_in, sb_in, sd_in = inner if inner is not None else (1, 0, 0, 0)
    A = _mat(a, M, K, lda, batch, sa, a_kmajor, bi, sa_in).double()
    B = _mat(b, N, K, ldb, batch, sb, b_kmajor, bi, sb_in).double()
    res = A @ B.transpose(2, 3)
    if second is not None:
        a2, b2, K2, lda2, ldb2, sa2, sb2 = second
        res = res + _mat(a2, M, K2, lda2, batch, sa2, a_kmajor).double() @ _mat(b2, N, K2, ldb2, batch, sb2, b_kmajor).double().transpose(2, 3)
    res = alpha * res
    if bias is not None:
        res = res + bias.as_strided((batch, 1, 1, N), (sbias, 0, 0, 1)).double()
    D = out.as_strided((batch, bi, M, N), (sd, sd_in, ldd, 1))
    if splits > 1 or accumulate:
        D.copy_((D.double() + res).to(out.dtype))
    else:
        D.copy_(res.to(out.dtype))
    return out


def emu_gemm(a, b, out, M, N, K, lda, ldb, a_kmajor, b_kmajor, bias=None, splits=1, alpha=1.0):
    return emu_bgemm(a, b, out, M, N, K, lda, ldb, out.stride(0), 1, 0, 0, 0, a_kmajor, b_kmajor, bias=bias, splits=splits, alpha=alpha)


def emu_colsum(src, n, dst=None):
    if dst is None:
        dst = torch.zeros(n, dtype=torch.float32)
    dst.add_(src[:, :n].double().sum(0).float())
    return dst


def emu_ln_fwd(x2d, w, b, out_dtype, eps=1e-5):
    x = x2d.double()
    mean = x.mean(1)
    rstd = 1.0 / torch.sqrt(((x - mean[:, None]) ** 2).mean(1) + eps)
    y = (x - mean[:, None]) * rstd[:, None] * w.double() + b.double()
    return y.to(out_dtype), mean.float(), rstd.float()


def emu_ln_bwd(dy, x2d, w, mean, rstd, dx, accumulate, dw=None, db=None):
    x, g = x2d.double(), dy.double()
    xh = (x - mean.double()[:, None]) * rstd.double()[:, None]
    gw = g * w.double()
    d = rstd.double()[:, None] * (gw - gw.mean(1, keepdim=True) - xh * (gw * xh).mean(1, keepdim=True))
    if accumulate:
        dx.copy_((dx.double() + d).to(dx.dtype))
    else:
        dx.copy_(d.to(dx.dtype))
    if dw is None:
        dw = torch.zeros(x.shape[1], dtype=torch.float32)
        db = torch.zeros(x.shape[1], dtype=torch.float32)
    dw.add_((g * xh).sum(0).float())
    db.add_(g.sum(0).float())
    return dw, db


def emu_colsoftmax_fwd(scores, p_out, n_sets, L, n_cols, lds, ldp, zero_key=False):
    S = scores.as_strided((n_sets, L, n_cols), (L * lds, lds, 1)).double()
    if zero_key:
        S = torch.cat([S, S.new_zeros(n_sets, 1, n_cols)], dim=1)
    P = torch.softmax(S, dim=1)[:, :L]
    p_out.as_strided((n_sets, L, n_cols), (L * ldp, ldp, 1)).copy_(P.to(p_out.dtype))


def emu_colsoftmax_bwd(p_in, d_p, ds_out, n_sets, L, n_cols, lds, ldp):
    P = p_in.as_strided((n_sets, L, n_cols), (L * ldp, ldp, 1)).double()
    dP = d_p.as_strided((n_sets, L, n_cols), (L * lds, lds, 1)).double()
    dS = P * (dP - (P * dP).sum(1, keepdim=True))
    ds_out.as_strided((n_sets, L, n_cols), (L * ldp, ldp, 1)).copy_(dS.to(ds_out.dtype))


def emu_addnorm_fwd(q_in, pooled):
    z = q_in.double() + pooled.double()
    inv = 1.0 / z.norm(dim=1).clamp_min(1e-12)
    return (z * inv[:, None]).to(q_in.dtype), inv.float()


def emu_addnorm_bwd(g_out, out, inv_norm, cd):
    g, o = g_out.double(), out.double()
    gz = inv_norm.double()[:, None] * (g - o * (g * o).sum(1, keepdim=True))
    return gz.float(), gz.to(cd)


def install(monkeypatch):
    from cosmos_b200 import _lib, pooler
    for name, fn in (("_gemm", emu_gemm), ("_bgemm", emu_bgemm), ("_colsum", emu_colsum), ("_ln_fwd", emu_ln_fwd), ("_ln_bwd", emu_ln_bwd),
                     ("_colsoftmax_fwd", emu_colsoftmax_fwd), ("_colsoftmax_bwd", emu_colsoftmax_bwd),
                     ("_addnorm_fwd", emu_addnorm_fwd), ("_addnorm_bwd", emu_addnorm_bwd)):
        monkeypatch.setattr(pooler, name, fn)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, what: None)

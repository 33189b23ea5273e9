"""Test-only CPU emulation of the kernel entry points of cosmos_b200.infonce.

It restates exactly what include/cosmos_b200.h documents for cosmos_infonce_fwd / _loss_sums / _bwd in
fp64 torch, so that the HOST logic (stacking, gathers, mode coefficients, LSE all-reduce, transposed
pass) can be exercised on CPU ranks with gloo.  It lives under tests/ and is never imported by the
package: the product path has no CPU implementation."""
import math

import torch

LOG2E = math.log2(math.e)
LN2 = math.log(2.0)


def _raw(x, y):
    return torch.einsum("ibd,jnd->ijbn", x.double(), y.double())       # [gx, gy, b, N]


def emu_fwd(x, y, label_offset, scale, keep_e=False, out=None):
    if out is not None:            # write into the caller's buffers (slices of per-group tensors)
        res = emu_fwd(x, y, label_offset, scale, keep_e)
        for dst, src in zip(out, res[:3]):
            dst.copy_(src)
        return tuple(out) + tuple(res[3:])
    if keep_e:
        # stored-exponential route: the emulation keeps the logits themselves (the layout of e / off is private to the kernels)
        return emu_fwd(x, y, label_offset, scale) + (_raw(x, y) * (float(scale) * LOG2E), None)
    gx, b, _ = x.shape
    gy, N, _ = y.shape
    raw = _raw(x, y)
    S2 = raw * (float(scale) * LOG2E)
    row = torch.logsumexp(S2 * LN2, dim=3) / LN2
    col = torch.logsumexp(S2 * LN2, dim=2) / LN2
    idx = torch.arange(b)
    diag = raw[:, :, idx, label_offset + idx]
    return (row.reshape(gx * gy, b).float(), diag.reshape(gx * gy, b).float(), col.reshape(gx * gy, N).float())


def emu_loss_sums(x, y, label_offset, scale, row_lse2, diag_raw, col_lse2):
    b = x.shape[1]
    d = float(scale) * diag_raw.double()
    rs = (LN2 * row_lse2.double() - d).sum(1)
    cs = (LN2 * col_lse2.double()[:, label_offset:label_offset + b] - d).sum(1)
    return torch.stack([rs, cs], dim=1).float()


def emu_bwd(x, y, label_offset, scale, row_lse2, col_lse2, a_row, a_col, s_row, s_col, weight, upstream, want_dx, want_dscale,
            g_out=None):
    gx, b, D = x.shape
    gy, N, _ = y.shape
    raw = _raw(x, y)
    S2 = raw * (float(scale) * LOG2E)
    R = torch.exp2(S2 - row_lse2.double().view(gx, gy, b, 1))
    Cm = torch.exp2(S2 - col_lse2.double().view(gx, gy, 1, N))
    eye = torch.zeros(b, N, dtype=torch.float64)
    eye[torch.arange(b), label_offset + torch.arange(b)] = 1
    G = a_row * R + a_col * Cm - (a_row + a_col) * eye
    Gs = s_row * R + s_col * Cm - (s_row + s_col) * eye
    if g_out is not None:        # cosmos_infonce_bwd_g: block (i, j) of G at rows i * b, columns j * N, in the stack dtype
        g_out[:, :gy * N] = G.permute(0, 2, 1, 3).reshape(gx * b, gy * N).to(g_out.dtype)
    up = float(upstream)
    dx = None
    if want_dx:
        dx = (up * weight * float(scale)) * torch.einsum("ijbn,jnd->ibd", G, y.double())
        dx = dx.to(x.dtype)
    dscale = None
    if want_dscale:
        dscale = ((up * weight) * (Gs * raw).sum()).float().reshape(1)
    return dx, dscale


def emu_scale16(src, num, den):
    """cosmos_scale16: src * (num / den), products in fp32, in src's dtype."""
    return (src.float() * (float(num) / den)).to(src.dtype)


def emu_lse2_merge(parts, out):
    """cosmos_lse2_merge: log2-sum-exp2 over the leading (rank) dimension."""
    out.copy_((torch.logsumexp(parts.double() * LN2, dim=0) / LN2).float())
    return out


def emu_bwd_e(x, y, label_offset, scale, e, off, diag_raw, row_lse2, col_lse2, a_row, a_col, s_row, s_col, weight, upstream,
              want_dscale, g_out=None, dx_out=None):
    """cosmos_infonce_bwd_e: the same gradient, formed from what the forward kept (here: the logits) - no x y^T."""
    assert abs(a_row * s_col - a_col * s_row) < 1e-12, "bwd_e needs proportional d(scale) / gradient mixes"
    gx, b, D = x.shape
    gy, N, _ = y.shape
    S2 = e
    R = torch.exp2(S2 - row_lse2.double().view(gx, gy, b, 1))
    Cm = torch.exp2(S2 - col_lse2.double().view(gx, gy, 1, N))
    eye = torch.zeros(b, N, dtype=torch.float64)
    eye[torch.arange(b), label_offset + torch.arange(b)] = 1
    G = a_row * R + a_col * Cm - (a_row + a_col) * eye
    if g_out is not None:
        g_out[:, :gy * N] = G.permute(0, 2, 1, 3).reshape(gx * b, gy * N).to(g_out.dtype)
    up = float(upstream)
    acc = torch.einsum("ijbn,jnd->ibd", G, y.double())                      # the dX accumulators
    dx = ((up * weight * float(scale)) * acc).to(x.dtype)
    if dx_out is not None:
        dx_out.copy_(dx)
        dx = dx_out
    dscale = None
    if want_dscale:       # sum_r <x_r, (G y)_r>, re-weighted to the d(scale) mix
        dscale = ((up * weight * (s_row + s_col) / (a_row + a_col)) * (acc * x.double()).sum()).float().reshape(1)
    return dx, dscale


def emu_bwd_e_cols(x, y, label_offset, scale, e, off, diag_raw, row_lse2, col_lse2, a_row, a_col, rank_major=0):
    """cosmos_infonce_bwd_e_cols: fp32 [gy, N, D] = sum over row tensors and local rows of G^T x, unit scale."""
    assert not rank_major, "the reduce-scatter send layout is only produced for NCCL groups"
    gx, b, D = x.shape
    gy, N, _ = y.shape
    S2 = e
    R = torch.exp2(S2 - row_lse2.double().view(gx, gy, b, 1))
    Cm = torch.exp2(S2 - col_lse2.double().view(gx, gy, 1, N))
    eye = torch.zeros(b, N, dtype=torch.float64)
    eye[torch.arange(b), label_offset + torch.arange(b)] = 1
    G = a_row * R + a_col * Cm - (a_row + a_col) * eye
    return torch.einsum("ijbn,ibd->jnd", G, x.double()).float()


def emu_colgrad(g, x2d, n_c, n_cols):
    """infonce._k_colgrad: fp32 [n_c, n_cols, D] = sum over rows of G^T x."""
    D = x2d.shape[1]
    return (g[:, :n_c * n_cols].double().T @ x2d.double()).reshape(n_c, n_cols, D).float()


def install(monkeypatch=None):
    """Patch cosmos_b200.infonce to run on CPU tensors through the emulation."""
    from cosmos_b200 import _lib, infonce
    if monkeypatch is not None:
        monkeypatch.setattr(infonce, "_k_fwd", emu_fwd)
        monkeypatch.setattr(infonce, "_k_loss_sums", emu_loss_sums)
        monkeypatch.setattr(infonce, "_k_bwd", emu_bwd)
        monkeypatch.setattr(infonce, "_k_bwd_e", emu_bwd_e)
        monkeypatch.setattr(infonce, "_k_colgrad", emu_colgrad)
        monkeypatch.setattr(infonce, "_k_lse2_merge", emu_lse2_merge)
        monkeypatch.setattr(infonce, "_k_scale16", emu_scale16)
        monkeypatch.setattr(infonce, "_k_bwd_e_cols", emu_bwd_e_cols)
        monkeypatch.setattr(_lib, "require_cuda", lambda t, what: None)
        monkeypatch.setattr(infonce, "compute_dtype", lambda dt: dt)
    else:
        infonce._k_fwd, infonce._k_loss_sums, infonce._k_bwd = emu_fwd, emu_loss_sums, emu_bwd
        infonce._k_colgrad = emu_colgrad
        infonce._k_bwd_e = emu_bwd_e
        infonce._k_lse2_merge = emu_lse2_merge
        infonce._k_scale16 = emu_scale16
        infonce._k_bwd_e_cols = emu_bwd_e_cols
        _lib.require_cuda = lambda t, what: None
        infonce.compute_dtype = lambda dt: dt

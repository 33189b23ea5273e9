"""CPU: the arithmetic of the stored-exponential route, restated in torch and checked against fp64.

The forward keeps e = bf16(2^(s2 - m)) with m = the running maximum of the row inside its 64-column group at the moment the
32-column chunk is processed (csrc/infonce_fwd.cu: one (max, sum) per thread and column group h = (column % 256) / 64), the
backward forms G = e * (a_row 2^(m - lse_row) + a_col 2^(m - lse_col)) - (a_row + a_col) [positive].  This file pins what
that representation can and cannot carry: its error on ordinary inputs, at logit scale 100, and the one documented loss -
column-softmax weights of elements more than 126 log2 units below their row's running maximum."""
import math

import pytest
import torch

LOG2E = math.log2(math.e)


def stored_exponentials(s2: torch.Tensor):
    """s2 [b, N] fp32 logits in log2 units -> (e bf16 [b, N], off fp32 [b, N // 32]) in the order the forward visits them."""
    b, n = s2.shape
    e = torch.zeros(b, n, dtype=torch.bfloat16)
    off = torch.zeros(b, (n + 31) // 32, dtype=torch.float32)
    m_run = torch.full((b, 4), -math.inf)                       # one running maximum per 64-column group of the 256-column tile
    for c0 in range(0, n, 32):
        h = (c0 % 256) // 64
        blk = s2[:, c0:c0 + 32]
        m_run[:, h] = torch.maximum(m_run[:, h], blk.max(dim=1).values)
        off[:, c0 // 32] = m_run[:, h]
        e[:, c0:c0 + 32] = torch.exp2(blk - m_run[:, h:h + 1]).to(torch.bfloat16)     # <= 1; fp32 flushes below 2^-126
    return e, off


def gradient_matrix(e, off, lse_row, lse_col, a_row, a_col, label0):
    """What csrc/infonce_bwd_e2.cu forms (fp32 arithmetic, exponents clamped like its slow path), before the bf16 rounding of G."""
    b, n = e.shape
    o = off.repeat_interleave(32, dim=1)[:, :n]
    g = e.float() * (a_row * torch.exp2(o - lse_row[:, None]) + a_col * torch.exp2(torch.clamp(o - lse_col[None, :], max=126.0)))
    g[torch.arange(b), label0 + torch.arange(b)] -= a_row + a_col
    return g


def exact(s2, a_row, a_col, label0):
    s = s2.double()
    lse_r = torch.logsumexp(s * math.log(2.0), dim=1) / math.log(2.0)
    lse_c = torch.logsumexp(s * math.log(2.0), dim=0) / math.log(2.0)
    g = a_row * torch.exp2(s - lse_r[:, None]) + a_col * torch.exp2(s - lse_c[None, :])
    g[torch.arange(s.shape[0]), label0 + torch.arange(s.shape[0])] -= a_row + a_col
    return g, lse_r.float(), lse_c.float()


def _features(b, n, dim, noise, seed):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(n, dim, generator=g)
    x = torch.nn.functional.normalize(z[:b] + noise * torch.randn(b, dim, generator=g), dim=-1).bfloat16().float()
    y = torch.nn.functional.normalize(z + noise * torch.randn(n, dim, generator=g), dim=-1).bfloat16().float()
    return x, y


@pytest.mark.parametrize("scale,noise", [(14.2857, 2.0), (100.0, 1.5), (100.0, 0.3)])
def test_stored_exponentials_reproduce_the_gradient(scale, noise):
    b, n, dim = 96, 600, 128
    x, y = _features(b, n, dim, noise, seed=int(scale))
    s2 = (x @ y.t()) * (scale * LOG2E)
    want, lse_r, lse_c = exact(s2, 1.0, 1.0, 0)
    e, off = stored_exponentials(s2)
    assert float(e.float().max()) <= 1.0
    got = gradient_matrix(e, off, lse_r, lse_c, 1.0, 1.0, 0).double()
    # element-wise: bf16 rounding of e (2^-9 relative) is the only error of the representation; the floor is the fp32
    # rounding of the log-sum-exps, visible where R + C - 2 nearly cancels at a confident positive
    assert bool(((got - want).abs() <= 2.0 ** -8 * want.abs() + 1e-5).all())
    # the gradient itself (dX = G Y): cosine far inside the north_star bound, norm within 1e-3
    dx_got, dx_want = got @ y.double(), want @ y.double()
    cos = float((dx_got * dx_want).sum() / (dx_got.norm() * dx_want.norm()))
    assert cos >= 0.999999, cos
    assert abs(float(dx_got.norm() / dx_want.norm()) - 1.0) <= 1e-3
    # d(scale) through the accumulator identity: sum_rc G_rc <x_r, y_c> = sum_r <x_r, (G Y)_r>
    ds_identity = float((dx_got * x.double()).sum())
    ds_direct = float((want * (x.double() @ y.double().t())).sum())
    assert abs(ds_identity - ds_direct) <= 2e-3 * abs(ds_direct) + 1e-6 * float(want.abs().sum())


def test_what_the_representation_drops():
    """A row whose running maximum is > 126 log2 units above an element stores that element as zero.  Its row-softmax weight
    is below fp32 resolution anyway; its column-softmax weight is lost - and matters only in a column where EVERY entry is that
    far below its own row's maximum.  Constructed here (logit range 200 log2 units = 139 nats: scale 100, cosines +-0.7)."""
    b, n = 63, 64
    s2 = torch.full((b, n), -100.0)
    s2[torch.arange(b), torch.arange(b)] = 100.0       # every row has a dominant positive in columns 0..62 ...
    low = n - 1                                        # ... and the last column (nobody's positive) is uniformly low
    want, lse_r, lse_c = exact(s2, 1.0, 1.0, 0)
    e, off = stored_exponentials(s2)
    got = gradient_matrix(e, off, lse_r, lse_c, 1.0, 1.0, 0).double()
    assert float((got[:, :low] - want[:, :low]).abs().max()) <= 2.0 ** -8 + 1e-5   # everything else is intact
    assert float(e[:, low].float().abs().max()) == 0.0                          # the whole column was flushed: 2^-200
    assert float(want[:, low].sum()) == pytest.approx(1.0, abs=1e-6)            # exact: its column softmax sums to 1
    assert float(got[:, low].abs().sum()) == 0.0                                # stored route: that weight is gone
    # with a logit range below 126 log2 units (87 nats: any scale <= 43, or realistic cosines at scale 100) nothing is dropped
    s2c = s2.clamp(min=-20.0)
    want, lse_r, lse_c = exact(s2c, 1.0, 1.0, 0)
    e, off = stored_exponentials(s2c)
    got = gradient_matrix(e, off, lse_r, lse_c, 1.0, 1.0, 0).double()
    assert float((got - want).abs().max()) <= 2.0 ** -8 + 1e-5

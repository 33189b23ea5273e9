"""Size-independent properties of the loss head at BASELINE.json's full batch (N = 32768, dim 512), where the CPU oracle
cannot be run: determinism, invariance under a permutation of the batch, the forward/backward identity in the logit
scale, and a problem whose logits are known in closed form.  A reduced number of crops keeps the runtime at seconds;
the tile counts, label offsets and column sweeps are those of the headline configuration."""
import math

import pytest
import torch

N, D = 32768, 512


def _features(n_rows, n_cols, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    z = torch.randn(N, D, device="cuda", generator=g)
    mk = lambda: torch.nn.functional.normalize(z + 2.0 * torch.randn(N, D, device="cuda", generator=g), dim=-1).bfloat16()
    return [mk() for _ in range(n_rows)], [mk() for _ in range(n_cols)]


def _loss_and_grads(rows, cols, scale):
    from cosmos_b200 import pairs_infonce
    rows = [t.detach().clone().requires_grad_(True) for t in rows]
    cols = [t.detach().clone().requires_grad_(True) for t in cols]
    s = torch.tensor(float(scale), device="cuda", requires_grad=True)
    loss = pairs_infonce(rows, cols, s)
    loss.backward()
    torch.cuda.synchronize()
    return loss.detach(), [t.grad for t in rows], [t.grad for t in cols], s.grad


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


@pytest.mark.gpu
def test_fullsize_deterministic_and_permutation_invariant():
    rows, cols = _features(2, 2, 7)
    l0, gr0, gc0, ds0 = _loss_and_grads(rows, cols, 14.2857)
    l1, gr1, gc1, ds1 = _loss_and_grads(rows, cols, 14.2857)
    assert torch.equal(l0, l1) and torch.equal(ds0, ds1)                       # no atomics on this path: bit-identical reruns
    assert all(torch.equal(a, b) for a, b in zip(gr0 + gc0, gr1 + gc1))
    assert math.isfinite(float(l0)) and 0.0 < float(l0) < math.log(N) + 1.0

    perm = torch.randperm(N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    lp, grp, gcp, dsp = _loss_and_grads([t[perm] for t in rows], [t[perm] for t in cols], 14.2857)
    assert abs(float(lp) - float(l0)) <= 2e-5 * abs(float(l0))
    assert abs(float(dsp) - float(ds0)) <= 1e-3 * abs(float(ds0)) + 1e-7
    for a, b in zip(grp + gcp, gr0 + gc0):
        assert _cos(a, b[perm]) > 0.99999
        assert float((a.float() - b[perm].float()).abs().max()) <= 2e-2 * float(b.float().abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [14.2857, 60.0])
def test_fullsize_scale_gradient_matches_forward_difference(scale):
    """d loss / d logit_scale from the backward kernels == central difference of the forward kernels."""
    from cosmos_b200 import pairs_infonce
    rows, cols = _features(2, 1, 11)
    _, _, _, ds = _loss_and_grads(rows, cols, scale)
    h = 0.01 * scale
    with torch.no_grad():
        lp = float(pairs_infonce(rows, cols, scale + h))
        lm = float(pairs_infonce(rows, cols, scale - h))
    fd = (lp - lm) / (2 * h)
    assert abs(fd - float(ds)) <= 2e-2 * abs(fd) + 2e-5, (fd, float(ds))


@pytest.mark.gpu
def test_fullsize_closed_form_logits():
    """Rows and columns are one-hot in 512 directions (sample i -> axis i mod 512): S_ij = scale * [i = j mod 512], so every
    row / column has N/512 logits equal to `scale` (one of them the positive) and the rest 0:
        loss = log(m e^s + N - m) - s,  m = N / 512,  for both directions; d loss / d scale = m e^s / Z - 1."""
    from cosmos_b200 import pairs_infonce
    idx = torch.arange(N, device="cuda") % D
    onehot = torch.zeros(N, D, device="cuda", dtype=torch.bfloat16)
    onehot[torch.arange(N, device="cuda"), idx] = 1.0
    s = 5.0
    rows = [onehot.clone().requires_grad_(True)]
    cols = [onehot.clone().requires_grad_(True)]
    sc = torch.tensor(s, device="cuda", requires_grad=True)
    loss = pairs_infonce(rows, cols, sc)
    loss.backward()
    m = N // D
    Z = m * math.exp(s) + (N - m)
    want = math.log(Z) - s
    assert abs(float(loss.detach()) - want) <= 1e-4 * want
    assert abs(float(sc.grad) - (m * math.exp(s) / Z - 1.0)) <= 1e-4
    # d loss / d x_i = (s / N) * (sum_j p_ij y_j - y_i), p the mean of the row- and column-softmax (identical here):
    # along the sample's own axis (m e^s / Z - 1), along every other axis m / Z
    g = rows[0].grad.float()
    own = g[torch.arange(N, device="cuda"), idx]
    torch.testing.assert_close(own, torch.full_like(own, (s / N) * (m * math.exp(s) / Z - 1.0)), rtol=2e-2, atol=0)
    off = g.sum(dim=1) - own
    torch.testing.assert_close(off, torch.full_like(off, (s / N) * (D - 1) * m / Z), rtol=2e-2, atol=0)

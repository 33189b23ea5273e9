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


@pytest.mark.gpu
def test_fullsize_headline_grouping_against_fp32_eager():
    """The headline configuration itself (BASELINE.json: global batch 32768, dim 512, 8 + 8 student, 8 + 8 cross-modal and 2 + 2
    teacher feature tensors per sample = 64 distillation + 16 CLIP pairs, bf16) through COSMOSLoss on the stored-exponential route,
    against the oracle's formulas (oracle.symmetric_infonce over src/open_clip/loss.py:116-117 logits, composition of
    loss.py:176-207) evaluated in fp32 on the same GPU, one 32768 x 32768 logit matrix at a time (the oracle's own loop keeps
    all 80 alive in one autograd graph; pair by pair with an immediate backward is the same sum).  north_star tolerance:
    loss relative error <= 1e-4, gradient cosine >= 0.9999 for every feature tensor, d logit_scale within 3e-3."""
    from cosmos_b200 import COSMOSLoss
    from oracle import cosmos_oracle as O
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(2024)
    z = torch.randn(N, D, device="cuda", generator=g)

    def view(noise):
        return torch.nn.functional.normalize(z + noise * torch.randn(N, D, device="cuda", generator=g), dim=-1).bfloat16()

    counts = {"s_image": 8, "s_text": 8, "s_img_x": 8, "s_txt_x": 8, "t_image": 2, "t_text": 2}
    feats = {k: [view(1.0 + 0.25 * i) for i in range(n)] for k, n in counts.items()}
    teacher = ("t_image", "t_text")
    x = {k: [t.clone().requires_grad_(k not in teacher) for t in v] for k, v in feats.items()}
    ls = torch.tensor(14.2857, device="cuda", requires_grad=True)
    ds = torch.tensor(30.0, device="cuda", requires_grad=True)
    out = COSMOSLoss(local_loss=False, gather_with_grad=False, cache_labels=True, rank=0, world_size=1)(
        x["s_image"], x["s_text"], ls, t_image_features=x["t_image"], t_text_features=x["t_text"], output_dict=True,
        distill_logit_scale=ds, s_img_crossmodal_features=x["s_img_x"], s_txt_crossmodal_features=x["s_txt_x"])
    (out["distill_loss"] + out["clip_loss"]).backward()
    torch.cuda.synchronize()

    r = {k: [t.float().requires_grad_(k not in teacher) for t in v] for k, v in feats.items()}
    rls = torch.tensor(14.2857, device="cuda", requires_grad=True)
    rds = torch.tensor(30.0, device="cuda", requires_grad=True)
    ref = {"distill_loss": 0.0, "clip_loss": 0.0}

    def group(name, a_list, b_list, scale, weight):
        for a in a_list:
            for b in b_list:
                logits = scale * a @ b.T                                  # loss.py:116: the other direction is its transpose
                term = weight * O.symmetric_infonce(logits, logits.T) / (len(a_list) * len(b_list))
                term.backward()
                ref[name] += float(term.detach())
                del logits, term

    group("distill_loss", r["s_img_x"], r["t_image"], rds, 0.25)          # loss.py:193-203: mean of the four groups
    group("distill_loss", r["s_img_x"], r["t_text"], rds, 0.25)
    group("distill_loss", r["s_txt_x"], r["t_image"], rds, 0.25)
    group("distill_loss", r["s_txt_x"], r["t_text"], rds, 0.25)
    group("clip_loss", r["s_image"][:2], r["s_text"], rls, 1.0)           # loss.py:205-206: the first two image crops only
    torch.cuda.synchronize()

    worst = {"loss_rel": 0.0, "grad_cos_min": 1.0, "grad_norm_rel": 0.0}
    for k in ref:
        worst["loss_rel"] = max(worst["loss_rel"], abs(float(out[k].detach()) - ref[k]) / abs(ref[k]))
        assert abs(float(out[k].detach()) - ref[k]) <= 1e-4 * abs(ref[k]), (k, float(out[k].detach()), ref[k])
    for k in ("s_image", "s_text", "s_img_x", "s_txt_x"):
        for i, (t, rt) in enumerate(zip(x[k], r[k])):
            if rt.grad is None:                                            # student image crops 2.. take no part in any term
                assert t.grad is None or float(t.grad.abs().max()) == 0.0, (k, i)
                continue
            c, nr = _cos(t.grad, rt.grad), abs(float(t.grad.double().norm() / rt.grad.double().norm()) - 1.0)
            worst["grad_cos_min"], worst["grad_norm_rel"] = min(worst["grad_cos_min"], c), max(worst["grad_norm_rel"], nr)
            assert c >= 0.9999, (k, i, c)
            assert nr <= 5e-3, (k, i, nr)
    worst["dscale_rel"] = max(abs(float(ls.grad) - float(rls.grad)) / abs(float(rls.grad)),
                              abs(float(ds.grad) - float(rds.grad)) / abs(float(rds.grad)))
    print("fullsize headline parity (N = 32768, 80 pairs, bf16 vs fp32 eager):", worst, "losses", ref)      # pytest -s
    assert abs(float(ls.grad) - float(rls.grad)) <= 3e-3 * abs(float(rls.grad))
    assert abs(float(ds.grad) - float(rds.grad)) <= 3e-3 * abs(float(rds.grad))

"""GPU parity of the InfoNCE kernels against the CPU oracle (fp32 on the same bf16-valued inputs).

Tolerances are the north_star's: loss relative error <= 1e-4, gradient cosine >= 0.9999."""
import ctypes as C
import math
import os

import pytest
import torch

from oracle import cosmos_oracle as O

LOSS_RTOL = 1e-4
LOSS_ATOL = 3e-6   # fp32 rounding floor of (LSE - positive) when the loss itself is ~0
GRAD_COS = 0.9999


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def lse2_ref(S2, dim):
    return torch.logsumexp(S2 * math.log(2.0), dim=dim) / math.log(2.0)


@pytest.mark.gpu
@pytest.mark.parametrize("gx,gy,b,W,rank,D,scale", [
    (1, 1, 128, 1, 0, 64, 10.0),
    (2, 3, 300, 3, 1, 128, 20.0),
    (3, 2, 77, 1, 0, 512, 100.0),
    (1, 2, 256, 2, 1, 512, 14.2857),
    (2, 1, 513, 1, 0, 192, 50.0),
])
def test_fwd_statistics(gx, gy, b, W, rank, D, scale):
    """row LSE, per-rank column LSE and positives of every pair block vs a direct fp64 computation."""
    from cosmos_b200 import infonce as K
    g = torch.Generator().manual_seed(gx * 100 + gy * 10 + b)
    N = W * b
    x = torch.nn.functional.normalize(torch.randn(gx, b, D, generator=g), dim=-1).bfloat16()
    y = torch.nn.functional.normalize(torch.randn(gy, N, D, generator=g), dim=-1).bfloat16()
    # make positives stand out a little
    y[:, rank * b:(rank + 1) * b] = (y[:, rank * b:(rank + 1) * b].float() * 0.6 + x[0].float() * 0.4).bfloat16()
    sc = torch.tensor([scale], dtype=torch.float32, device="cuda")
    row, diag, col = K._k_fwd(x.cuda(), y.cuda(), rank * b, sc)
    torch.cuda.synchronize()
    xd, yd = x.double(), y.double()
    for i in range(gx):
        for j in range(gy):
            raw = xd[i] @ yd[j].T                                   # [b, N]
            S2 = raw * (float(sc.item()) * math.log2(math.e))
            p = i * gy + j
            torch.testing.assert_close(row[p].cpu().double(), lse2_ref(S2, 1), rtol=0, atol=2e-4)
            torch.testing.assert_close(col[p].cpu().double(), lse2_ref(S2, 0), rtol=0, atol=2e-4)
            want_diag = raw[torch.arange(b), rank * b + torch.arange(b)]
            torch.testing.assert_close(diag[p].cpu().double(), want_diag, rtol=0, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [100.0, 30.0])
def test_fwd_statistics_extreme_range(scale):
    """Logits spanning [-scale, +scale] inside one 32-row block: the one-exp fast path must detect the columns whose
    terms would underflow relative to the block's offset and redo them exactly; same for the backward's per-warp range
    check (gradients compared with the closed form)."""
    from cosmos_b200 import infonce as K
    g = torch.Generator().manual_seed(5)
    b, D = 256, 128
    u = torch.nn.functional.normalize(torch.randn(D, generator=g), dim=0)
    x = torch.nn.functional.normalize(torch.randn(1, b, D, generator=g), dim=-1)
    y = torch.nn.functional.normalize(torch.randn(1, b, D, generator=g), dim=-1)
    x[0, :64] = u                      # rows 0..63 all equal u
    y[0, 0:8] = u                      # columns 0..7: +scale against those rows
    y[0, 8:16] = -u                    # columns 8..15: -scale against them (2*scale below the block's row maxima)
    y[0, 200:204] = -u
    x, y = x.bfloat16(), y.bfloat16()
    sc = torch.tensor([scale], dtype=torch.float32, device="cuda")
    row, diag, col = K._k_fwd(x.cuda(), y.cuda(), 0, sc)
    raw = x[0].double() @ y[0].double().T
    S2 = raw * (scale * math.log2(math.e))
    torch.testing.assert_close(row[0].cpu().double(), lse2_ref(S2, 1), rtol=0, atol=3e-4)
    torch.testing.assert_close(col[0].cpu().double(), lse2_ref(S2, 0), rtol=0, atol=3e-4)
    up = torch.ones(1, device="cuda")
    w = 1.0 / (2 * b)
    dx, dsc = K._k_bwd(x.cuda(), y.cuda(), 0, sc, row, col, 1.0, 1.0, 1.0, 1.0, w, up, True, True)
    _, da, _, dscale = O.pair_closed_form(x[0].double(), y[0].double(), scale)
    assert cosine(dx[0].float().cpu(), da) >= GRAD_COS
    assert abs(float(dx[0].float().cpu().norm() / da.norm()) - 1) < 5e-3
    assert abs(float(dsc) - float(dscale)) <= 3e-3 * abs(float(dscale)) + 1e-6


def _run_ours(inp, ls, ds, up, dtype):
    from cosmos_b200 import COSMOSLoss
    leaf = {k: [t.to(dtype).cuda().requires_grad_(True) for t in v] for k, v in inp.items()}
    lsd = torch.tensor(ls, device="cuda", requires_grad=True)
    dsd = None if ds is None else torch.tensor(ds, device="cuda", requires_grad=True)
    out = COSMOSLoss(cache_labels=True)(leaf["s_image"], leaf["s_text"], lsd, t_image_features=leaf["t_image"],
                                        t_text_features=leaf["t_text"], output_dict=True, distill_logit_scale=dsd,
                                        s_img_crossmodal_features=leaf["s_img_x"], s_txt_crossmodal_features=leaf["s_txt_x"])
    assert set(out) == {"distill_loss", "clip_loss"} and out["clip_loss"].dim() == 0 and out["clip_loss"].dtype == torch.float32
    (up[0] * out["distill_loss"] + up[1] * out["clip_loss"]).backward()
    return out, leaf, lsd, dsd


def _run_oracle(inp, ls, ds, up, dtype):
    leaf = {k: [t.to(dtype).float().requires_grad_(True) for t in v] for k, v in inp.items()}
    lsd = torch.tensor(ls, requires_grad=True)
    dsd = None if ds is None else torch.tensor(ds, requires_grad=True)
    out = O.cosmos_loss_single(leaf["s_image"], leaf["s_text"], lsd, leaf["t_image"], leaf["t_text"], dsd,
                               leaf["s_img_x"], leaf["s_txt_x"])
    (up[0] * out["distill_loss"] + up[1] * out["clip_loss"]).backward()
    return out, leaf, lsd, dsd


def _compare(ours, ref, up):
    out, leaf, ls, ds = ours
    rout, rleaf, rls, rds = ref
    for k in ("distill_loss", "clip_loss"):
        a, b = float(out[k]), float(rout[k])
        assert abs(a - b) <= LOSS_RTOL * abs(b) + LOSS_ATOL, (k, a, b)
    for k, lst in rleaf.items():
        for t, r in zip(leaf[k], lst):
            if r.grad is None:
                assert t.grad is None, k
            else:
                assert t.grad is not None and t.grad.dtype == t.dtype, k
                assert cosine(t.grad.cpu().float(), r.grad) >= GRAD_COS, (k, cosine(t.grad.cpu().float(), r.grad))
                rel = (t.grad.cpu().float().norm() / r.grad.norm()).item()
                assert abs(rel - 1) < 5e-3, (k, rel)
    assert abs(float(ls.grad) - float(rls.grad)) <= 2e-3 * abs(float(rls.grad)) + 1e-6 * max(up), (float(ls.grad), float(rls.grad))
    if ds is not None:
        assert abs(float(ds.grad) - float(rds.grad)) <= 2e-3 * abs(float(rds.grad)) + 1e-6 * max(up)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_cosmos_loss_golden_small(golden_dir, dtype):
    """The committed reference cases (ragged batch sizes 24/40/17, dims 64/128, scales up to 100,
    GradScaler-sized upstream gradients)."""
    cases = torch.load(os.path.join(golden_dir, "cosmos_w1_small.pt"), weights_only=False)
    for case in cases:
        up = case["upstream"]
        ours = _run_ours(case["inputs"], case["logit_scale"], case["distill_logit_scale"], up, dtype)
        ref = _run_oracle(case["inputs"], case["logit_scale"], case["distill_logit_scale"], up, dtype)
        _compare(ours, ref, up)
        # and against what the reference itself produced on the un-rounded fp32 inputs (looser: input rounding)
        for k in ("distill_loss", "clip_loss"):
            assert abs(float(ours[0][k]) - float(case["out"][k])) <= 5e-3 * abs(float(case["out"][k])) + 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("batch,dim,scale,keep_g", [(256, 512, 14.2857, False), (384, 512, 100.0, False), (1000, 256, 30.0, False),
                                                    (256, 512, 14.2857, True), (392, 512, 100.0, True),
                                                    (256, 512, 14.2857, "e"), (392, 512, 100.0, "e"), (77, 512, 50.0, "e")])
def test_cosmos_loss_vs_oracle(batch, dim, scale, keep_g, monkeypatch):
    """BASELINE config 1 shape (batch 256, dim 512, 2 global + 6 local crops) and two more.  keep_g: the image-side CLIP
    gradient through the stored G tiles + GEMM (cosmos_infonce_bwd_g), the route large batches take, forced at this size.
    "e": the stored-exponential route (cosmos_infonce_fwd_e / _bwd_e, gradients formed in forward in chunks of three row
    tensors), the route the headline batch takes, forced at this size (ragged batches included)."""
    if keep_g == "e":
        from cosmos_b200 import infonce
        calls = []
        real = infonce._k_bwd_e
        monkeypatch.setattr(infonce, "_e_store_chunk", lambda x_r, y_c, comm: min(3, x_r.shape[0]))
        monkeypatch.setattr(infonce, "_k_bwd_e", lambda *a, **k: calls.append(1) or real(*a, **k))
    elif keep_g:
        from cosmos_b200 import infonce
        monkeypatch.setattr(infonce, "_G_STORE_MIN_BYTES", 0)
        calls = []
        real = infonce._k_colgrad
        monkeypatch.setattr(infonce, "_k_colgrad", lambda *a: calls.append(1) or real(*a))
    inp = O.make_features(batch, dim, seed=1234)
    up = (1.0, 1.0)
    ours = _run_ours(inp, scale, scale, up, torch.bfloat16)
    ref = _run_oracle(inp, scale, scale, up, torch.bfloat16)
    _compare(ours, ref, up)
    if keep_g:
        assert calls, "the forced route was not taken"


@pytest.mark.gpu
@pytest.mark.parametrize("batch,scale,route", [(200, 14.2857, "g"), (130, 60.0, "g"), (200, 14.2857, "e"), (130, 60.0, "e")])
def test_cosmos_loss_fp16_dim512(batch, scale, route, monkeypatch):
    """fp16 features (the reference's `--precision amp`) through the dim-512 cluster kernels, ragged batch; batch 200 also
    takes the stored-G route for the image-side CLIP gradient (130 is not a multiple of 8 and falls back to the second sweep)."""
    from cosmos_b200 import infonce
    monkeypatch.setattr(infonce, "_G_STORE_MIN_BYTES", 0)
    if route == "e":     # stored exponentials are bf16 whatever the feature dtype is; G is converted to fp16 in shared memory
        monkeypatch.setattr(infonce, "_e_store_chunk", lambda x_r, y_c, comm: x_r.shape[0])
    inp = O.make_features(batch, 512, seed=77)
    up = (65536.0, 65536.0)          # GradScaler's initial scale
    ours = _run_ours(inp, scale, scale * 0.7, up, torch.float16)
    ref = _run_oracle(inp, scale, scale * 0.7, up, torch.float16)
    _compare(ours, ref, up)


@pytest.mark.gpu
def test_config1_fp32_inputs_against_reference_golden(golden_dir):
    """fp32 features as in BASELINE config 1: the kernels round them to bf16; the loss must still sit
    within the north_star tolerance of what the reference produced in fp32."""
    import importlib.util
    rec = torch.load(os.path.join(golden_dir, "cosmos_w1_cfg1.pt"), weights_only=False)
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(golden_dir, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    inp = mg.cosmos_inputs(torch.Generator().manual_seed(rec["seed"]), rec["batch"], rec["dim"])
    ours = _run_ours(inp, rec["logit_scale"], rec["distill_logit_scale"], (1.0, 1.0), torch.float32)
    out, leaf, ls, ds = ours
    for k in ("distill_loss", "clip_loss"):
        assert abs(float(out[k]) - float(rec["out"][k])) <= LOSS_RTOL * abs(float(rec["out"][k])), k
    assert leaf["s_img_x"][0].grad.dtype == torch.float32
    assert cosine(leaf["s_img_x"][0].grad.cpu(), rec["g_s_img_x0"]) >= GRAD_COS
    assert cosine(leaf["s_text"][3].grad.cpu(), rec["g_s_text3"]) >= GRAD_COS
    assert leaf["s_image"][2].grad is None
    assert abs(float(ls.grad) - float(rec["g_logit_scale"])) <= 2e-3 * abs(float(rec["g_logit_scale"]))
    assert abs(float(ds.grad) - float(rec["g_distill_scale"])) <= 2e-3 * abs(float(rec["g_distill_scale"]))


@pytest.mark.gpu
def test_clip_loss_api_and_errors():
    from cosmos_b200 import ClipLoss, COSMOSLoss
    g = torch.Generator().manual_seed(0)
    a = torch.nn.functional.normalize(torch.randn(64, 128, generator=g), dim=-1).bfloat16().cuda().requires_grad_(True)
    b = torch.nn.functional.normalize(torch.randn(64, 128, generator=g), dim=-1).bfloat16().cuda().requires_grad_(True)
    loss = ClipLoss()(a, b, torch.tensor(20.0, device="cuda"))
    ref = O.clip_loss_single(a.detach().cpu().float(), b.detach().cpu().float(), 20.0)
    assert abs(float(loss) - float(ref)) <= LOSS_RTOL * float(ref)
    d = ClipLoss()(a, b, 20.0, output_dict=True)
    assert set(d) == {"contrastive_loss"}
    with pytest.raises(RuntimeError):
        ClipLoss()(a.detach().cpu(), b.detach().cpu(), 20.0)         # no CPU fallback
    with pytest.raises(RuntimeError):
        ClipLoss(use_horovod=True)
    with pytest.raises(AssertionError):
        COSMOSLoss()([a], [b], 20.0, t_image_features=[a], t_text_features=[b, b], s_img_crossmodal_features=[a],
                     s_txt_crossmodal_features=[b])
    with pytest.raises(RuntimeError):
        x = torch.randn(8, 100, device="cuda").bfloat16()            # dim not a multiple of 64
        ClipLoss()(x, x, 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("b,n_all,off,gx,gy,scale,dtype", [(200, 600, 400, 2, 3, 30.0, torch.bfloat16),
                                                           (128, 256, 128, 1, 1, 100.0, torch.bfloat16),
                                                           (77, 231, 77, 3, 2, 14.2857, torch.float16),
                                                           (300, 1200, 0, 2, 2, 60.0, torch.bfloat16)])
def test_stored_exponential_kernels_match_recompute_kernels(b, n_all, off, gx, gy, scale, dtype):
    """cosmos_infonce_fwd_e + _bwd_e against cosmos_infonce_fwd + _bwd on a row block that is NOT the first of its batch
    (label_offset = rank * b as on rank > 0 of a multi-GPU job), ragged sizes, both 16-bit dtypes, a_row != a_col."""
    from cosmos_b200 import infonce as K
    g = torch.Generator().manual_seed(b + n_all + off)
    z = torch.randn(n_all, 512, generator=g)
    y = torch.nn.functional.normalize(z[None] + 1.5 * torch.randn(gy, n_all, 512, generator=g), dim=-1).to(dtype).cuda()
    x = torch.nn.functional.normalize(z[None, off:off + b] + 1.5 * torch.randn(gx, b, 512, generator=g), dim=-1).to(dtype).cuda()
    sc = torch.tensor([scale], device="cuda")
    up = torch.tensor([3.0], device="cuda")
    row, diag, col = K._k_fwd(x, y, off, sc)
    row_e, diag_e, col_e, e, offs = K._k_fwd(x, y, off, sc, True)
    assert torch.equal(row, row_e) and torch.equal(diag, diag_e) and torch.equal(col, col_e)
    for a_row, a_col, ratio in ((1.0, 1.0, 1.0), (1.0, 0.0, 0.5), (0.25, 1.0, 2.0)):
        mix = (a_row, a_col, ratio * a_row, ratio * a_col, 0.125)
        dx_ref, ds_ref = K._k_bwd(x, y, off, sc, row, col, *mix, up, True, True)
        dx, ds = K._k_bwd_e(x, y, off, sc, e, offs, diag, row, col, *mix, up, True)
        assert cosine(dx.float().cpu(), dx_ref.float().cpu()) >= 0.99999
        assert abs(float(dx.float().norm()) / float(dx_ref.float().norm()) - 1.0) <= (5e-3 if scale >= 100.0 else 2e-3)
        # bf16 G in one kernel, fp32 G in the other; the floor covers mixes whose d(scale) nearly cancels (confident rows at
        # scale 100: sum (R - I) raw ~ 1e-3 of its terms' magnitude)
        assert abs(float(ds) - float(ds_ref)) <= 5e-3 * abs(float(ds_ref)) + 1e-4
    if n_all % 8 == 0:
        g1 = torch.zeros(gx * b, gy * n_all, dtype=dtype, device="cuda")
        g2 = torch.zeros_like(g1)
        K._k_bwd(x, y, off, sc, row, col, 1.0, 1.0, 1.0, 1.0, 0.125, up, True, False, g1)
        K._k_bwd_e(x, y, off, sc, e, offs, diag, row, col, 1.0, 1.0, 1.0, 1.0, 0.125, up, False, g2)
        assert cosine(g1.float().cpu(), g2.float().cpu()) >= 0.99999
    with pytest.raises(RuntimeError, match="unsupported"):      # d(scale) weights not proportional to the gradient weights
        K._k_bwd_e(x, y, off, sc, e, offs, diag, row, col, 1.0, 1.0, 1.0, 0.0, 0.125, up, True)
    # column side from the same exponentials (cosmos_infonce_bwd_e_cols: dY = G^T X, nothing transposed, no G in memory)
    # against fp64 from first principles: G = a_row softmax_rows + a_col softmax_cols - (a_row + a_col) positives
    S2 = torch.einsum("ibd,jnd->ijbn", x.double(), y.double()) * (scale * math.log2(math.e))
    pos = torch.zeros(b, n_all, dtype=torch.float64, device="cuda")
    pos[torch.arange(b), off + torch.arange(b)] = 1
    for a_row, a_col in ((1.0, 1.0), (1.0, 0.0), (0.25, 1.0)):
        G = (a_row * torch.exp2(S2 - row.double().view(gx, gy, b, 1)) + a_col * torch.exp2(S2 - col.double().view(gx, gy, 1, n_all))
             - (a_row + a_col) * pos)
        want = torch.einsum("ijbn,ibd->jnd", G, x.double())
        got = K._k_bwd_e_cols(x, y, off, sc, e, offs, diag, row, col, a_row, a_col)
        assert got.shape == (gy, n_all, 512) and got.dtype == torch.float32
        # G is a bf16 operand in the kernel: at scale 100 (confident rows, G = R + C - 2 I nearly cancels) its rounding shows
        floor = 0.9999 if scale >= 100.0 else 0.99999
        assert cosine(got.cpu(), want.cpu()) >= floor, (a_row, a_col, cosine(got.cpu(), want.cpu()))
        assert abs(float(got.double().norm() / want.norm()) - 1.0) <= 2e-3
        # every column tensor on its own (a column tensor with a wrong neighbour's tiles would still pass the global cosine)
        for jj in range(gy):
            assert cosine(got[jj].cpu(), want[jj].cpu()) >= floor, (a_row, a_col, jj)
        if (a_row, a_col) == (1.0, 1.0):
            # the reduce-scatter's send layout (what an NCCL run asks for): [W, gy, N / W, 512], the same values
            for W in (w for w in (1, 2, 4) if n_all % w == 0):
                rm = K._k_bwd_e_cols(x, y, off, sc, e, offs, diag, row, col, a_row, a_col, rank_major=W)
                assert rm.shape == (W, gy, n_all // W, 512) and rm.is_contiguous()
                ref = got.view(gy, W, n_all // W, 512).transpose(0, 1)       # (the sum over sweep slices may associate differently)
                assert float((rm - ref).abs().max()) <= 1e-6 * float(ref.abs().max())
        if n_all % 8 == 0 and (a_row, a_col) == (1.0, 1.0):
            # and against the first version of this route: the row pass's own G tiles (same bf16 values) through a GEMM
            via_gemm = K._k_colgrad(g2, x.reshape(gx * b, 512), gy, n_all)
            assert cosine(got.cpu(), via_gemm.cpu()) >= 0.999999

"""Pins the CPU oracle to fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import pytest
import torch

from oracle import cosmos_oracle as O


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _close(a, b, rtol=2e-5, atol=1e-6):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def _leafs(inp):
    return {k: [t.clone().requires_grad_(True) for t in v] for k, v in inp.items()}


def test_cosmos_single_process_matches_reference(golden_dir):
    for case in _load(golden_dir, "cosmos_w1_small.pt"):
        leaf = _leafs(case["inputs"])
        ls = torch.tensor(case["logit_scale"], requires_grad=True)
        ds = None if case["distill_logit_scale"] is None else torch.tensor(case["distill_logit_scale"], requires_grad=True)
        out = O.cosmos_loss_single(leaf["s_image"], leaf["s_text"], ls, leaf["t_image"], leaf["t_text"], ds,
                                   leaf["s_img_x"], leaf["s_txt_x"])
        up = case["upstream"]
        (up[0] * out["distill_loss"] + up[1] * out["clip_loss"]).backward()
        for k in ("distill_loss", "clip_loss"):
            _close(out[k].detach(), case["out"][k])
        for k, lst in case["grads"].items():
            for t, g in zip(leaf[k], lst):
                if g is None:
                    assert t.grad is None
                else:
                    _close(t.grad, g, rtol=1e-4, atol=1e-6 * max(up))
        _close(ls.grad, case["g_logit_scale"], rtol=1e-4, atol=1e-5 * max(up))
        if ds is not None:
            _close(ds.grad, case["g_distill_scale"], rtol=1e-4, atol=1e-5 * max(up))


def test_cosmos_config1_summary(golden_dir):
    """BASELINE config 1: batch 256, dim 512, 2+6 crops, fp32 (inputs re-generated from the seed)."""
    rec = _load(golden_dir, "cosmos_w1_cfg1.pt")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(golden_dir, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = torch.Generator().manual_seed(rec["seed"])
    leaf = _leafs(mg.cosmos_inputs(g, rec["batch"], rec["dim"]))
    ls = torch.tensor(rec["logit_scale"], requires_grad=True)
    ds = torch.tensor(rec["distill_logit_scale"], requires_grad=True)
    out = O.cosmos_loss_single(leaf["s_image"], leaf["s_text"], ls, leaf["t_image"], leaf["t_text"], ds,
                               leaf["s_img_x"], leaf["s_txt_x"])
    (out["distill_loss"] + out["clip_loss"]).backward()
    _close(out["distill_loss"].detach(), rec["out"]["distill_loss"])
    _close(out["clip_loss"].detach(), rec["out"]["clip_loss"])
    _close(leaf["s_img_x"][0].grad, rec["g_s_img_x0"], rtol=1e-4, atol=1e-8)
    _close(leaf["s_text"][3].grad, rec["g_s_text3"], rtol=1e-4, atol=1e-8)
    assert leaf["s_image"][2].grad is None and rec["grad_summary"]["s_image"][2] is None   # loss.py:205-206
    assert abs(leaf["s_image"][0].grad.norm().item() - rec["grad_summary"]["s_image"][0]["norm"]) < 1e-6
    _close(ls.grad, rec["g_logit_scale"], rtol=2e-4, atol=2e-6)
    _close(ds.grad, rec["g_distill_scale"], rtol=2e-4, atol=2e-6)


def test_closed_form_pair_gradients():
    g = torch.Generator().manual_seed(3)
    a = torch.nn.functional.normalize(torch.randn(19, 48, generator=g, dtype=torch.float64), dim=-1).requires_grad_(True)
    b = torch.nn.functional.normalize(torch.randn(19, 48, generator=g, dtype=torch.float64), dim=-1).requires_grad_(True)
    s = torch.tensor(37.0, dtype=torch.float64, requires_grad=True)
    loss = O.clip_loss_single(a, b, s)
    loss.backward()
    l2, da, db, dsc = O.pair_closed_form(a.detach(), b.detach(), 37.0)
    _close(l2, loss.detach(), rtol=1e-12, atol=1e-12)
    _close(da, a.grad, rtol=1e-10, atol=1e-12)
    _close(db, b.grad, rtol=1e-10, atol=1e-12)
    _close(dsc, s.grad, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("fname", ["multirank_w2.pt", "multirank_w4.pt"])
def test_multirank_semantics_match_reference(golden_dir, fname):
    """Per-rank loss values and gradients of all four (local_loss, gather_with_grad)
    modes, reference run on gloo ranks vs the one-process simulation."""
    rec = _load(golden_dir, fname)
    W = rec["world"]
    for name, spec in rec["payload"].items():
        ll, gwg = spec["local_loss"], spec["gather_with_grad"]
        shards = spec["shards"]
        if spec["kind"] == "clip":
            A = [[t.clone().requires_grad_(True) for t in s["a"]] for s in shards]
            B = [[t.clone().requires_grad_(True) for t in s["b"]] for s in shards]
            sc = [torch.tensor(spec["logit_scale"], requires_grad=True) for _ in range(W)]
            losses = [O.clip_loss_rank(A, B, sc[r], r, ll, gwg) for r in range(W)]
            # autograd through a with-grad all_gather sums every rank's backward
            sum(losses).backward()
            for r in range(W):
                ref = rec["results"][r][name]
                _close(losses[r].detach(), ref["loss"])
                for t, gref in zip(A[r], ref["ga"]):
                    _close(t.grad, gref, rtol=2e-4, atol=2e-6)
                for t, gref in zip(B[r], ref["gb"]):
                    _close(t.grad, gref, rtol=2e-4, atol=2e-6)
                _close(sc[r].grad, ref["gscale"], rtol=1e-4, atol=1e-6)
        else:
            leafs = []
            for s in shards:
                d = _leafs(s)
                d["logit_scale"] = torch.tensor(spec["logit_scale"], requires_grad=True)
                d["distill_logit_scale"] = torch.tensor(spec["distill_logit_scale"], requires_grad=True)
                leafs.append(d)
            outs = [O.cosmos_loss_rank(leafs, r, ll, gwg) for r in range(W)]
            sum(o["distill_loss"] + o["clip_loss"] for o in outs).backward()
            for r in range(W):
                ref = rec["results"][r][name]
                for k in ("distill_loss", "clip_loss"):
                    _close(outs[r][k].detach(), ref["out"][k])
                for k, lst in ref["grads"].items():
                    for t, gref in zip(leafs[r][k], lst):
                        if gref is None:
                            assert t.grad is None or float(t.grad.abs().max()) == 0.0
                        else:
                            _close(t.grad, gref, rtol=2e-4, atol=2e-6)
                _close(leafs[r]["logit_scale"].grad, ref["g_logit_scale"], rtol=1e-4, atol=1e-6)
                _close(leafs[r]["distill_logit_scale"].grad, ref["g_distill_scale"], rtol=1e-4, atol=1e-6)


def test_pooler_matches_reference(golden_dir):
    for rec in _load(golden_dir, "pooler.pt"):
        params, tokens, feats, w = O.make_pooler_case(rec["d"], rec["L"], rec["batch_size"], rec["n"], rec["seed"])
        params = {k: v.requires_grad_(True) for k, v in params.items()}
        tokens.requires_grad_(True)
        feats.requires_grad_(True)
        B, n = rec["batch_size"], rec["n"]
        pooled = O.cross_pool(tokens[:B].repeat(n, 1, 1), feats.unsqueeze(1), params, rec["heads"])
        _close(pooled, rec["pooled"], rtol=1e-4, atol=2e-5)
        xmodal = O.cosmos_crossmodal(feats, tokens, params, rec["heads"], B)
        _close(xmodal, rec["xmodal"], rtol=1e-4, atol=1e-5)
        (xmodal * w).sum().backward()
        _close(feats.grad, rec["g_feats"], rtol=1e-3, atol=1e-5)
        _close(tokens.grad[:, :2], rec["g_tokens_head"], rtol=1e-3, atol=1e-5)
        assert abs(tokens.grad.norm().item() - rec["g_tokens_norm"]) <= 1e-4 * rec["g_tokens_norm"] + 1e-7
        for k, v in params.items():
            assert abs(v.grad.norm().item() - rec["g_param_norm"][k]) <= 1e-3 * rec["g_param_norm"][k] + 1e-6, k
            _close(v.grad.reshape(-1)[:64], rec["g_param_head"][k], rtol=2e-3, atol=2e-5)


def test_pooler_add_zero_attn_matches_reference(golden_dir):
    """add_zero_attn=True (transformer.py:214-221): the oracle's branch against the unmodified reference module - outputs and
    every gradient, key bias included (with the extra zero key its gradient is no longer zero)."""
    for rec in _load(golden_dir, "pooler_zero_attn.pt"):
        params, tokens, feats, w = O.make_pooler_case(rec["d"], rec["L"], rec["batch_size"], rec["n"], rec["seed"])
        params = {k: v.requires_grad_(True) for k, v in params.items()}
        tokens.requires_grad_(True)
        feats.requires_grad_(True)
        B, n = rec["batch_size"], rec["n"]
        pooled = O.cross_pool(tokens[:B].repeat(n, 1, 1), feats.unsqueeze(1), params, rec["heads"], add_zero_attn=True)
        _close(pooled, rec["pooled"], rtol=1e-4, atol=2e-5)
        xmodal = torch.nn.functional.normalize(feats + pooled.squeeze(1), dim=-1)           # model.py:379-380
        _close(xmodal, rec["xmodal"], rtol=1e-4, atol=1e-5)
        (xmodal * w).sum().backward()
        _close(feats.grad, rec["g_feats"], rtol=1e-3, atol=1e-5)
        _close(tokens.grad, rec["g_tokens"], rtol=1e-3, atol=1e-5)
        for k, v in params.items():
            _close(v.grad, rec["g_params"][k], rtol=2e-3, atol=2e-5)
        d = rec["d"]
        assert float(rec["g_params"]["attn.in_proj_bias"][d:2 * d].abs().max()) > 0        # the key bias matters here
        plain = O.cross_pool(tokens[:B].repeat(n, 1, 1), feats.unsqueeze(1), params, rec["heads"])
        assert not torch.allclose(plain, rec["pooled"], rtol=1e-3, atol=1e-4)


def test_ema_matches_reference(golden_dir):
    rec = _load(golden_dir, "ema.pt")
    for m, want in rec["outs"].items():
        k = [t.clone() for t in rec["teacher"]]
        O.ema_update_(k, rec["student"], m)
        for a, b in zip(k, want):
            assert torch.equal(a, b)


def test_retrieval_metrics_match_reference(golden_dir):
    """oracle.get_clip_metrics / compute_retrieval against the reference's own functions (train.py:712-785)."""
    for rec in _load(golden_dir, "retrieval.pt"):
        img, txt, txt2img, img2txt = O.make_retrieval_case(rec["n_img"], rec["caps"], rec["dim"], rec["seed"], rec["noise"])
        got = O.compute_retrieval(14.2857 * img @ txt.t(), txt2img, img2txt)
        assert got.keys() == rec["compute_retrieval"].keys()
        for k, v in rec["compute_retrieval"].items():
            assert got[k] == pytest.approx(v, rel=1e-6, abs=0), k
        if "get_clip_metrics" in rec:
            img_p, txt_p, _, _ = O.make_retrieval_case(rec["n_img"], 1, rec["dim"], rec["seed"], rec["noise"], shuffle=False)
            got = O.get_clip_metrics(img_p, txt_p, torch.tensor(14.2857))
            assert got.keys() == rec["get_clip_metrics"].keys()
            for k, v in rec["get_clip_metrics"].items():
                assert got[k] == pytest.approx(v, rel=1e-12, abs=0), k


def test_clamp_matches_reference(golden_dir):
    rec = _load(golden_dir, "clamp.pt")
    for name, want in rec["outs"].items():
        ts = [torch.tensor(v, dtype=want.dtype) for v in rec["values"]]
        O.clamp_logit_scales_(ts)
        got = torch.stack(ts)
        assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(want, nan=-7.0)), name
        assert torch.isnan(got).tolist() == torch.isnan(want).tolist()

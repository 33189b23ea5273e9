"""GPU: the loss head, the EMA update and the pooler inside a CUDA graph.

The library keeps no host state, takes (device, stream) explicitly and allocates only through torch's caching allocator
(DESIGN.md §1), so a training step that contains it can be captured with torch.cuda.graph and replayed - the way a
launch-bound step (BASELINE config 2: ~130 launches of 3 - 700 us) is made independent of the host.  Each test captures
forward + backward once, replays the graph on NEW inputs copied into the static buffers and compares with an eager call."""
import pytest
import torch

from oracle import cosmos_oracle as O

pytestmark = pytest.mark.gpu

KEYS = ("s_image", "s_text", "s_img_x", "s_txt_x", "t_image", "t_text")


def _inputs(b, seed):
    f = O.make_features(b, 512, seed=seed)
    return {k: torch.stack([t.bfloat16() for t in f[k]]).cuda() for k in KEYS}       # one [n, b, 512] stack per list


def _step(loss_fn, x, ls, ds):
    out = loss_fn([*x["s_image"]], [*x["s_text"]], ls, t_image_features=[*x["t_image"]], t_text_features=[*x["t_text"]],
                  output_dict=True, distill_logit_scale=ds, s_img_crossmodal_features=[*x["s_img_x"]],
                  s_txt_crossmodal_features=[*x["s_txt_x"]])
    total = out["distill_loss"] + out["clip_loss"]
    total.backward()
    return total


@pytest.mark.parametrize("route", ["recompute", "stored-exponentials"])
def test_loss_head_step_replays_from_a_cuda_graph(route, monkeypatch):
    from cosmos_b200 import COSMOSLoss, infonce
    monkeypatch.setattr(infonce, "_E_STORE_MIN_BYTES", 0 if route == "stored-exponentials" else 1 << 60)
    b = 640
    loss_fn = COSMOSLoss(local_loss=False, gather_with_grad=False, cache_labels=True, rank=0, world_size=1)
    grads_of = ("s_image", "s_text", "s_img_x", "s_txt_x")
    static = {k: v.clone().requires_grad_(k in grads_of) for k, v in _inputs(b, 1).items()}
    ls = torch.tensor(14.2857, device="cuda", requires_grad=True)
    ds = torch.tensor(30.0, device="cuda", requires_grad=True)
    leaves = [static[k] for k in grads_of] + [ls, ds]

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                       # warm-up outside the capture: per-shape decisions are cached here
        for _ in range(2):
            for t in leaves:
                t.grad = None
            _step(loss_fn, static, ls, ds)
    torch.cuda.current_stream().wait_stream(side)
    for t in leaves:
        t.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_total = _step(loss_fn, static, ls, ds)

    for seed in (2, 3):                                 # replay on new data
        fresh = _inputs(b, seed)
        with torch.no_grad():
            for k in KEYS:
                static[k].copy_(fresh[k])
        graph.replay()
        torch.cuda.synchronize()
        got_total = float(static_total.detach())
        got = [t.grad.detach().clone() for t in leaves]
        eager = {k: v.clone().requires_grad_(k in grads_of) for k, v in fresh.items()}
        els = torch.tensor(14.2857, device="cuda", requires_grad=True)
        eds = torch.tensor(30.0, device="cuda", requires_grad=True)
        want_total = float(_step(loss_fn, eager, els, eds).detach())
        want = [eager[k].grad for k in grads_of] + [els.grad, eds.grad]
        assert got_total == want_total, (seed, got_total, want_total)           # same kernels, same order: bit-identical
        for g, w in zip(got, want):
            assert torch.equal(g, w)


def test_ema_update_replays_from_a_cuda_graph():
    from cosmos_b200 import ema_update_
    g = torch.Generator().manual_seed(5)
    shapes = [(1000, 33), (77,), (4096, 64), (1,), (513, 17)]
    student = [torch.randn(*s, generator=g).cuda() for s in shapes]
    teacher = [torch.randn(*s, generator=g).cuda() for s in shapes]
    ref = [t.clone() for t in teacher]
    ema_update_(student, teacher, 0.99)                 # builds the chunk table outside the capture
    O.ema_update_(ref, student, 0.99)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ema_update_(student, teacher, 0.99)
    for _ in range(3):
        with torch.no_grad():
            for s in student:
                s.add_(0.01)
        graph.replay()
        O.ema_update_(ref, student, 0.99)
    torch.cuda.synchronize()
    for t, r in zip(teacher, ref):
        assert torch.equal(t, r)


def test_pooler_step_replays_from_a_cuda_graph():
    from cosmos_b200.pooler import AttentionalCrossPooler, crossmodal_features
    d, L, B, n, heads = 512, 77, 16, 8, 8
    params, tokens, feats, w = O.make_pooler_case(d, L, B, n, 7)
    mod = AttentionalCrossPooler(d, d, heads).cuda()
    mod.load_state_dict(params)
    tok = tokens.bfloat16().cuda().requires_grad_(True)
    f = feats.bfloat16().cuda().requires_grad_(True)
    wd = w.cuda()
    leaves = [tok, f, *mod.parameters()]

    def step():
        out = crossmodal_features(mod, tok, f, B)
        (out.float() * wd).sum().backward()
        return out

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            for t in leaves:
                t.grad = None
            step()
    torch.cuda.current_stream().wait_stream(side)
    for t in leaves:
        t.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_out = step()
    _, tokens2, feats2, _ = O.make_pooler_case(d, L, B, n, 8)
    with torch.no_grad():
        tok.copy_(tokens2.bfloat16())
        f.copy_(feats2.bfloat16())
    graph.replay()
    torch.cuda.synchronize()
    got_out = static_out.detach().clone()
    got = [t.grad.detach().clone() for t in leaves]
    for t in leaves:
        t.grad = None
    want_out = step().detach()
    torch.cuda.synchronize()
    assert torch.equal(got_out, want_out)
    for g_, t in zip(got, leaves):
        # weight gradients are split-K sums with fp32 atomics: equal up to the order of the additions
        assert torch.allclose(g_.float(), t.grad.float(), rtol=2e-3, atol=2e-3 * float(t.grad.float().abs().max()))

"""Generate golden fixtures by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py            # writes tests/golden/*.pt

The reference (/root/reference, read-only) is imported through an empty
namespace package so that `open_clip/__init__.py` (which needs ftfy/timm) is
skipped - SURVEY.md §8(c).  Nothing is copied from it: only its *outputs* on
seeded inputs are stored.  The GPU box has no /root/reference; tests read the
fixtures only.
"""
import importlib
import os
import sys
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("COSMOS_REFERENCE", "/root/reference")


def ref_modules():
    pkg = types.ModuleType("open_clip")
    pkg.__path__ = [os.path.join(REF, "src", "open_clip")]
    sys.modules["open_clip"] = pkg
    return importlib.import_module("open_clip.loss"), importlib.import_module("open_clip.transformer")


def unit(g, *shape, dtype=torch.float32):
    return torch.nn.functional.normalize(torch.randn(*shape, generator=g), dim=-1).to(dtype)


def cosmos_inputs(g, b, d, n_img=8, n_txt=8, corr=True, noise=2.0):
    z = torch.randn(b, d, generator=g) if corr else None

    def v():
        x = torch.randn(b, d, generator=g)
        if z is not None:
            x = z + noise * x
        return torch.nn.functional.normalize(x, dim=-1)

    return {"s_image": [v() for _ in range(n_img)], "s_text": [v() for _ in range(n_txt)],
            "s_img_x": [v() for _ in range(n_img)], "s_txt_x": [v() for _ in range(n_txt)],
            "t_image": [v() for _ in range(2)], "t_text": [v() for _ in range(2)]}


def run_cosmos(L, inp, logit_scale, distill_scale, up=(1.0, 1.0), **ctor):
    leaf = {k: [t.clone().requires_grad_(True) for t in v] for k, v in inp.items()}
    ls = torch.tensor(logit_scale, requires_grad=True)
    ds = torch.tensor(distill_scale, requires_grad=True) if distill_scale is not None else None
    loss = L.COSMOSLoss(cache_labels=True, **ctor)
    out = loss(leaf["s_image"], leaf["s_text"], ls, t_image_features=leaf["t_image"], t_text_features=leaf["t_text"],
               output_dict=True, distill_logit_scale=ds, s_img_crossmodal_features=leaf["s_img_x"],
               s_txt_crossmodal_features=leaf["s_txt_x"])
    (up[0] * out["distill_loss"] + up[1] * out["clip_loss"]).backward()
    grads = {k: [None if t.grad is None else t.grad.clone() for t in v] for k, v in leaf.items()}
    return ({k: v.detach().clone() for k, v in out.items()}, grads,
            ls.grad.clone(), None if ds is None else ds.grad.clone())


def case_w1_small(L):
    g = torch.Generator().manual_seed(101)
    cases = []
    for (b, d, ls, dsc, corr, up, noise) in [(24, 64, 14.2857, 9.5, True, (1.0, 1.0), 1.0),
                                             (40, 128, 100.0, None, True, (65536.0, 65536.0), 2.5),
                                             (17, 64, 30.0, 100.0, False, (0.5, 2.0), 0.0)]:
        inp = cosmos_inputs(g, b, d, corr=corr, noise=noise)
        out, grads, gls, gds = run_cosmos(L, inp, ls, dsc, up)
        cases.append(dict(inputs=inp, logit_scale=ls, distill_logit_scale=dsc, upstream=up,
                          out=out, grads=grads, g_logit_scale=gls, g_distill_scale=gds))
    torch.save(cases, os.path.join(HERE, "cosmos_w1_small.pt"))


def case_cfg1(L):
    """BASELINE config 1 (batch 256, dim 512): inputs are re-generated from the seed by
    the test, only summaries of the outputs are stored."""
    g = torch.Generator().manual_seed(1234)
    inp = cosmos_inputs(g, 256, 512)
    out, grads, gls, gds = run_cosmos(L, inp, 14.2857, 14.2857)
    summ = {k: [None if t is None else dict(norm=t.norm().item(), head=t[:2, :8].clone(), rowsum=t.sum(1)[:16].clone())
                for t in v] for k, v in grads.items()}
    torch.save(dict(seed=1234, batch=256, dim=512, logit_scale=14.2857, distill_logit_scale=14.2857,
                    out=out, grad_summary=summ, g_logit_scale=gls, g_distill_scale=gds,
                    # the full grads of two tensors, bf16-packed to stay small
                    g_s_img_x0=grads["s_img_x"][0].clone(), g_s_text3=grads["s_text"][3].clone()),
               os.path.join(HERE, "cosmos_w1_cfg1.pt"))


def _worker(rank, world, port, payload, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L, _ = ref_modules()
    res = {}
    for name, (kind, shards, ls, dsc, local_loss, gwg) in payload.items():
        mine = shards[rank]
        if kind == "clip":
            a = [t.clone().requires_grad_(True) for t in mine["a"]]
            b = [t.clone().requires_grad_(True) for t in mine["b"]]
            s = torch.tensor(ls, requires_grad=True)
            loss = L.ClipLoss(local_loss=local_loss, gather_with_grad=gwg, cache_labels=True, rank=rank, world_size=world)
            val = loss(a, b, s)
            val.backward()
            res[name] = dict(loss=val.detach().clone(), ga=[t.grad.clone() for t in a], gb=[t.grad.clone() for t in b],
                             gscale=s.grad.clone())
        else:
            out, grads, gls, gds = run_cosmos(L, mine, ls, dsc, local_loss=local_loss, gather_with_grad=gwg,
                                              rank=rank, world_size=world)
            res[name] = dict(out=out, grads=grads, g_logit_scale=gls, g_distill_scale=gds)
    torch.save(res, os.path.join(tmpdir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def case_multirank(world, port, fname):
    g = torch.Generator().manual_seed(77 + world)
    payload = {}
    b, d = 6, 32
    clip_shards = [dict(a=[unit(g, b, d) for _ in range(2)], b=[unit(g, b, d) for _ in range(3)]) for _ in range(world)]
    for ll in (False, True):
        for gwg in (False, True):
            payload[f"clip_ll{int(ll)}_gwg{int(gwg)}"] = ("clip", clip_shards, 20.0, None, ll, gwg)
    cos_shards = [cosmos_inputs(g, 8, 32, n_img=3, n_txt=4) for _ in range(world)]
    for ll, gwg in ((False, False), (False, True), (True, True), (True, False)):
        payload[f"cosmos_ll{int(ll)}_gwg{int(gwg)}"] = ("cosmos", cos_shards, 14.2857, 25.0, ll, gwg)
    import tempfile
    ctx = mp.get_context("spawn")
    with tempfile.TemporaryDirectory() as tmpdir:
        procs = [ctx.Process(target=_worker, args=(r, world, port, payload, tmpdir)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
            assert p.exitcode == 0
        got = {r: torch.load(os.path.join(tmpdir, f"rank{r}.pt")) for r in range(world)}
    torch.save(dict(world=world, payload={k: dict(kind=v[0], shards=v[1], logit_scale=v[2], distill_logit_scale=v[3],
                                                  local_loss=v[4], gather_with_grad=v[5]) for k, v in payload.items()},
                    results=[got[r] for r in range(world)]), os.path.join(HERE, fname))


def case_pooler(T):
    """Parameters/inputs come from oracle.make_pooler_case (seeded) so the d=512 cases only
    store outputs and gradient summaries."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle.cosmos_oracle import make_pooler_case
    cases = []
    for idx, (d, h, L_k, B, n) in enumerate([(64, 4, 7, 3, 2), (128, 8, 13, 2, 4), (512, 8, 77, 2, 8), (512, 8, 196, 1, 8)]):
        params, tokens, feats, w = make_pooler_case(d, L_k, B, n, seed=50 + idx)
        mod = T.AttentionalCrossPooler(d, d, h)
        mod.load_state_dict(params)
        tokens.requires_grad_(True)
        feats.requires_grad_(True)
        # model.py:378-380 / 382-384
        pooled = mod(tokens[:B].repeat(n, 1, 1), feats.unsqueeze(1))
        xmodal = torch.nn.functional.normalize(feats + pooled.squeeze(), dim=-1)
        (xmodal * w).sum().backward()
        gp = {k: v.grad.clone() for k, v in mod.named_parameters()}
        rec = dict(d=d, heads=h, L=L_k, batch_size=B, n=n, seed=50 + idx, pooled=pooled.detach().clone(),
                   xmodal=xmodal.detach().clone(), g_feats=feats.grad.clone(),
                   g_param_norm={k: v.norm().item() for k, v in gp.items()},
                   g_param_head={k: v.reshape(-1)[:64].clone() for k, v in gp.items()},
                   g_tokens_norm=tokens.grad.norm().item(), g_tokens_head=tokens.grad[:, :2].clone())
        if d <= 128:
            rec.update(g_tokens=tokens.grad.clone(), g_params=gp)
        cases.append(rec)
    torch.save(cases, os.path.join(HERE, "pooler.pt"))


def case_pooler_zero_attn(T):
    """AttentionalCrossPooler(add_zero_attn=True) (transformer.py:214-221; off in every COSMOS recipe, supported by the drop-in):
    outputs and every gradient at small shapes, through the module's own forward and through the model.py:378-380 lines."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle.cosmos_oracle import make_pooler_case
    cases = []
    for idx, (d, h, L_k, B, n) in enumerate([(64, 4, 7, 3, 2), (128, 8, 13, 2, 4)]):
        params, tokens, feats, w = make_pooler_case(d, L_k, B, n, seed=70 + idx)
        mod = T.AttentionalCrossPooler(d, d, h, add_zero_attn=True)
        mod.load_state_dict(params)
        tokens.requires_grad_(True)
        feats.requires_grad_(True)
        pooled = mod(tokens[:B].repeat(n, 1, 1), feats.unsqueeze(1))
        xmodal = torch.nn.functional.normalize(feats + pooled.squeeze(), dim=-1)
        (xmodal * w).sum().backward()
        cases.append(dict(d=d, heads=h, L=L_k, batch_size=B, n=n, seed=70 + idx, pooled=pooled.detach().clone(),
                          xmodal=xmodal.detach().clone(), g_feats=feats.grad.clone(), g_tokens=tokens.grad.clone(),
                          g_params={k: v.grad.clone() for k, v in mod.named_parameters()}))
    torch.save(cases, os.path.join(HERE, "pooler_zero_attn.pt"))


def case_ema():
    g = torch.Generator().manual_seed(9)
    shapes = [(1,), (), (7,), (33, 5), (4096,), (1000, 3)]
    teacher = [torch.randn(s, generator=g) * 0.02 for s in shapes]
    student = [torch.randn(s, generator=g) * 0.02 for s in shapes]
    t0 = [t.clone() for t in teacher]
    outs = {}
    for m in (0.99, 0.999, 0.5):
        k = [t.clone() for t in t0]
        with torch.no_grad():   # literal train.py:200-203
            for pq, pk in zip(student, k):
                pk.data.mul_(m).add_((1 - m) * pq.detach().data)
        outs[m] = k
    torch.save(dict(teacher=t0, student=student, outs=outs), os.path.join(HERE, "ema.pt"))


def case_clamp():
    """Logit-scale clamps at the end of the training step, literal train.py:237-243 on seeded scalars."""
    import math
    vals = [4.7, -0.3, 2.6593, 4.60517, 100.0, 0.0, 4.6051702, 4.605171, float("nan")]
    out = {}
    for name, dt in (("float32", torch.float32), ("bfloat16", torch.bfloat16), ("float16", torch.float16)):
        ts = [torch.tensor(v, dtype=dt) for v in vals]
        with torch.no_grad():
            for t in ts:
                t.clamp_(0, math.log(100))
        out[name] = torch.stack(ts)
    torch.save(dict(values=vals, outs=out), os.path.join(HERE, "clamp.pt"))


def _gather_worker(rank, world, port, payload, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L, _ = ref_modules()
    res = {}
    img0, txt0 = payload["image"][rank], payload["text"][rank]
    wi, wt = payload["probe_image"], payload["probe_text"]
    for ll in (False, True):
        for gwg in (False, True):
            # gather_features (src/open_clip/loss.py:21-65) + a probe loss that weighs every gathered row differently
            img, txt = img0.clone().requires_grad_(True), txt0.clone().requires_grad_(True)
            all_i, all_t = L.gather_features(img, txt, local_loss=ll, gather_with_grad=gwg, rank=rank, world_size=world)
            probe = (all_i * wi).sum() + (all_t * wt).sum()
            if probe.requires_grad:
                probe.backward()
            rec = dict(all_image=all_i.detach().clone(), all_text=all_t.detach().clone(),
                       g_image=None if img.grad is None else img.grad.clone(), g_text=None if txt.grad is None else txt.grad.clone())
            # ClipLoss.get_logits (loss.py:103-119) in the same mode
            img, txt = img0.clone().requires_grad_(True), txt0.clone().requires_grad_(True)
            lpi, lpt = L.ClipLoss(local_loss=ll, gather_with_grad=gwg, rank=rank, world_size=world).get_logits(img, txt, 7.5)
            rec["logits_per_image"], rec["logits_per_text"] = lpi.detach().clone(), lpt.detach().clone()
            (lpi * payload["probe_logits"][:lpi.shape[0], :lpi.shape[1]]).sum().backward()
            rec["g_logits_image"] = None if img.grad is None else img.grad.clone()
            rec["g_logits_text"] = None if txt.grad is None else txt.grad.clone()
            res[f"ll{int(ll)}_gwg{int(gwg)}"] = rec
    torch.save(res, os.path.join(tmpdir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def case_gather(world, port, fname):
    """gather_features and ClipLoss.get_logits of the reference on gloo ranks, all four (local_loss, gather_with_grad) modes:
    outputs and the gradients a probe loss sends back to the local shards."""
    import tempfile
    g = torch.Generator().manual_seed(500 + world)
    b, d = 5, 16
    payload = dict(image=[torch.randn(b, d, generator=g) for _ in range(world)], text=[torch.randn(b, d, generator=g) for _ in range(world)],
                   probe_image=torch.randn(world * b, d, generator=g), probe_text=torch.randn(world * b, d, generator=g),
                   probe_logits=torch.randn(world * b, world * b, generator=g))
    ctx = mp.get_context("spawn")
    with tempfile.TemporaryDirectory() as tmpdir:
        procs = [ctx.Process(target=_gather_worker, args=(r, world, port, payload, tmpdir)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
            assert p.exitcode == 0
        got = [torch.load(os.path.join(tmpdir, f"rank{r}.pt")) for r in range(world)]
    torch.save(dict(world=world, payload=payload, results=got), os.path.join(HERE, fname))


def ref_function_source(relpath, name):
    """One top-level function of a reference file as an ast module (compiled where the file lies, nothing is copied)."""
    import ast
    path = os.path.join(REF, relpath)
    tree = ast.parse(open(path).read(), filename=path)
    tree.body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name]
    assert tree.body, (relpath, name)
    return compile(tree, path, "exec")


def ref_train_step_lines():
    """The literal statements of the reference's training step that touch the loss head, as source text read from the file
    where it lies: src/training/train.py:162-188 (model_out dict, loss call, sum), 190 (backward) and 195-203 (EMA loop)."""
    import textwrap
    lines = open(os.path.join(REF, "src", "training", "train.py")).read().split("\n")
    block = lambda a, b: textwrap.dedent("\n".join(lines[a - 1:b]))
    return block(162, 188) + "\n" + block(190, 190) + "\n" + block(195, 203) + "\n"


class ToyTowers(torch.nn.Module):
    """Stand-in for the COSMOS student / teacher (src/open_clip/model.py:348-408): per-channel gains that turn fixed inputs
    into the feature dict train_one_epoch reads.  Outputs are rounded to bf16 values (what autocast hands the loss)."""

    def __init__(self, d, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        mk = lambda: torch.nn.Parameter(1.0 + 0.3 * torch.randn(d, generator=g))
        self.w_img, self.w_txt, self.w_imgx, self.w_txtx = mk(), mk(), mk(), mk()
        self.bias = torch.nn.Parameter(0.05 * torch.randn(d, generator=g))
        self.logit_scale = torch.nn.Parameter(torch.tensor(2.6593))            # ln(1 / 0.07), model.py:249
        self.distill_logit_scale = torch.nn.Parameter(torch.tensor(3.0))

    def forward(self, images, texts, batch_size=None):
        f = lambda x, w: torch.nn.functional.normalize(x * w + self.bias, dim=-1).to(torch.bfloat16).to(torch.float32)
        out = {"image_features": f(images, self.w_img), "text_features": f(texts, self.w_txt),
               "logit_scale": self.logit_scale.exp(), "distill_logit_scale": self.distill_logit_scale.exp()}
        if batch_size is not None:
            out["img_crossmodal_features"] = f(images, self.w_imgx)
            out["txt_crossmodal_features"] = f(texts, self.w_txtx)
        return out


def dropin_inputs(d, b, n_img, n_txt, seed):
    """Seeded inputs of the drop-in case (regenerated by the test; the fixture keeps a checksum): correlated views of one
    latent per sample, so that the losses sit where a partly trained model's do."""
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(b, d, generator=g)
    images = torch.cat([z + 1.5 * torch.randn(b, d, generator=g) for _ in range(n_img)])
    texts = torch.cat([z + 1.5 * torch.randn(b, d, generator=g) for _ in range(n_txt)])
    return images, texts, 0.01 * torch.randn(7, d, generator=g)


class FakeScaler:
    """torch.cuda.amp.GradScaler as the step uses it (train.py:62-66): scale(loss) = loss * 65536."""

    def scale(self, x):
        return x * 65536.0


def case_dropin():
    """The loss-head lines of train_one_epoch executed LITERALLY (read from the reference file) with the reference's
    create_loss(args) (factory.py:372-415): feature dict -> loss dict -> sum -> GradScaler.scale().backward() -> EMA loop."""
    import copy
    L, _ = ref_modules()
    ns = {name: getattr(L, name) for name in ("ClipLoss", "COSMOSLoss", "CoCaLoss", "DistillClipLoss", "SigLipLoss")}
    exec(ref_function_source(os.path.join("src", "open_clip", "factory.py"), "create_loss"), ns)
    tns = {"torch": torch}
    exec(ref_function_source(os.path.join("src", "training", "train.py"), "backward"), tns)
    args = types.SimpleNamespace(distill=False, model="ViT-B-16", siglip=False, cosmos=True, local_loss=False, gather_with_grad=False,
                                 rank=0, world_size=1, horovod=False, fix_momentum=True, momentum_teacher=0.99, accum_freq=1)
    d, b, n_img, n_txt, seed = 512, 40, 8, 8, 4242
    images, texts, noise = dropin_inputs(d, b, n_img, n_txt, seed)
    student = ToyTowers(d, 11)
    teacher = copy.deepcopy(student)
    with torch.no_grad():
        for p_, n_ in zip(teacher.parameters(), noise):
            p_.add_(n_[:p_.numel()].reshape(p_.shape))
            p_.requires_grad = False
    s_model_out = student(images, texts, b)
    t_model_out = teacher(torch.cat(images.chunk(n_img)[:2]), texts[:2 * b])
    env = dict(torch=torch, args=args, loss=ns["create_loss"](args), scaler=FakeScaler(), backward=tns["backward"],
               student=student, teacher=teacher, s_model_out=s_model_out, t_model_out=t_model_out,
               logit_scale=s_model_out["logit_scale"], distill_logit_scale=s_model_out["distill_logit_scale"],
               num_images=n_img, num_texts=n_txt, step=0, momentum_scheduler=None)
    exec(compile(ref_train_step_lines(), "train.py:162-203", "exec"), env)
    torch.save(dict(d=d, b=b, n_img=n_img, n_txt=n_txt, seed=seed, checksum=float(images.double().sum() + texts.double().sum()),
                    momentum=0.99, losses={k: v.detach().clone() for k, v in env["losses"].items()},
                    grads={k: p_.grad.clone() for k, p_ in student.named_parameters()},
                    teacher_after={k: v.clone() for k, v in teacher.state_dict().items()}), os.path.join(HERE, "dropin.pt"))


def ref_train_functions(*names):
    """The named top-level functions of the reference's src/training/train.py, compiled from the file where it lies
    (the module itself cannot be imported here: it needs open_clip's package __init__, PIL, tqdm, ...)."""
    import ast
    import numpy as np
    path = os.path.join(REF, "src", "training", "train.py")
    tree = ast.parse(open(path).read(), filename=path)
    tree.body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    ns = {"torch": torch, "np": np}
    exec(compile(tree, path, "exec"), ns)
    return [ns[n] for n in names]


def case_retrieval():
    """Eval metrics (train.py:712-763, 766-785) on seeded features from oracle.make_retrieval_case; only the metric
    dictionaries are stored, the tests regenerate the inputs."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle.cosmos_oracle import make_retrieval_case
    get_clip_metrics, compute_retrieval = ref_train_functions("get_clip_metrics", "compute_retrieval")
    cases = []
    for (n_img, caps, dim, seed, noise) in [(60, 5, 64, 301, 1.5), (131, 3, 96, 302, 2.5), (257, 1, 512, 303, 3.0)]:
        img, txt, txt2img, img2txt = make_retrieval_case(n_img, caps, dim, seed, noise)
        sim = 14.2857 * img @ txt.t()
        rec = dict(n_img=n_img, caps=caps, dim=dim, seed=seed, noise=noise,
                   compute_retrieval={k: float(v) for k, v in compute_retrieval(sim, txt2img, img2txt).items()})
        if caps == 1:
            img_p, txt_p, _, _ = make_retrieval_case(n_img, 1, dim, seed, noise, shuffle=False)
            rec["get_clip_metrics"] = {k: float(v) for k, v in get_clip_metrics(img_p, txt_p, torch.tensor(14.2857)).items()}
        cases.append(rec)
    torch.save(cases, os.path.join(HERE, "retrieval.pt"))


if __name__ == "__main__":
    torch.set_num_threads(4)
    L, T = ref_modules()
    case_w1_small(L)
    case_cfg1(L)
    case_multirank(4, 29611, "multirank_w4.pt")
    case_multirank(2, 29612, "multirank_w2.pt")
    case_gather(2, 29613, "gather_w2.pt")
    case_gather(3, 29614, "gather_w3.pt")
    case_dropin()
    case_pooler(T)
    case_pooler_zero_attn(T)
    case_ema()
    case_retrieval()
    case_clamp()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)))

"""CPU oracle for the COSMOS loss head (TEST INFRASTRUCTURE ONLY).

This file is a plain-PyTorch (CPU, fp32 or fp64) restatement of the reference's
algorithm for the hot path named in BASELINE.json.  It is the *checker*: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import it.  Nothing under `cosmos_b200/` imports it, and the
product path fails loudly when the CUDA extension is missing.

Pinning: the reference ships no tests and no golden vectors, so
the oracle is pinned against outputs of the reference itself, run in the build
container by `tests/golden/make_golden.py` (which imports
/root/reference/src/open_clip/{loss,transformer}.py unmodified) and committed as
fixtures under `tests/golden/`.  `tests/test_oracle_golden.py` checks every
function here against those fixtures.

Reference anchors (paths relative to /root/reference):
  * gather semantics          src/open_clip/loss.py:21-65
  * pairwise InfoNCE          src/open_clip/loss.py:103-142
  * COSMOS loss composition   src/open_clip/loss.py:176-207
  * cross-attention pooler    src/open_clip/transformer.py:210-230 and the call
                              site src/open_clip/model.py:366-387
  * EMA teacher update        src/training/train.py:195-203
  * logit-scale clamps        src/training/train.py:237-243
  * eval retrieval metrics    src/training/train.py:712-763, 766-785
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# InfoNCE over pairs
# ----------------------------------------------------------------------------

def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def symmetric_infonce(logits_ab: torch.Tensor, logits_ba: torch.Tensor, first_label: int = 0) -> torch.Tensor:
    """(CE(rows of logits_ab) + CE(rows of logits_ba)) / 2 with targets
    first_label + arange(n).   loss.py:134-138 (labels: loss.py:90-101)."""
    n = logits_ab.shape[0]
    target = torch.arange(n, dtype=torch.long) + first_label
    lse_ab = torch.logsumexp(logits_ab, dim=1)
    lse_ba = torch.logsumexp(logits_ba, dim=1)
    pos_ab = logits_ab[torch.arange(n), target]
    pos_ba = logits_ba[torch.arange(n), target]
    return 0.5 * ((lse_ab - pos_ab).mean() + (lse_ba - pos_ba).mean())


def pair_logits_single(a: torch.Tensor, b: torch.Tensor, scale) -> Tuple[torch.Tensor, torch.Tensor]:
    """world_size == 1 branch, loss.py:116-117."""
    return scale * a @ b.T, scale * b @ a.T


def clip_loss_single(a_list, b_list, scale) -> torch.Tensor:
    """Mean of the symmetric InfoNCE over the cartesian product of two feature
    lists, single process.  loss.py:121-142."""
    a_list, b_list = _as_list(a_list), _as_list(b_list)
    acc = 0
    for a in a_list:
        for b in b_list:
            lab, lba = pair_logits_single(a, b, scale)
            acc = acc + symmetric_infonce(lab, lba)
    return acc / (len(a_list) * len(b_list))


def cosmos_loss_single(s_image, s_text, logit_scale, t_image, t_text,
                       distill_logit_scale=None, s_img_x=None, s_txt_x=None) -> Dict[str, torch.Tensor]:
    """COSMOSLoss.forward with world_size == 1.  loss.py:176-207.
    Teacher lists are detached (loss.py:187,191); only the first two student
    image crops enter the CLIP term (loss.py:205-206)."""
    s_image, s_text = _as_list(s_image), _as_list(s_text)
    assert len(t_image) == 2 and len(t_text) == 2
    t_image = [t.detach() for t in t_image]
    t_text = [t.detach() for t in t_text]
    dscale = distill_logit_scale if distill_logit_scale is not None else logit_scale
    distill = (clip_loss_single(s_img_x, t_image, dscale) + clip_loss_single(s_img_x, t_text, dscale)
               + clip_loss_single(s_txt_x, t_image, dscale) + clip_loss_single(s_txt_x, t_text, dscale)) / 4
    clip = clip_loss_single(s_image[:2], s_text, logit_scale)
    return {"distill_loss": distill, "clip_loss": clip}


# ----------------------------------------------------------------------------
# Multi-rank semantics, simulated in one process
# ----------------------------------------------------------------------------

def _gathered_view(shards: Sequence[torch.Tensor], rank: int, local_loss: bool, gather_with_grad: bool):
    """What `gather_features` hands rank `rank` for one side (loss.py:49-63):
      gather_with_grad -> every shard keeps its graph (all_gather with autograd)
      else             -> remote shards are constants; the local shard is spliced
                          back in (keeps grad) only when not local_loss."""
    parts = []
    for r, s in enumerate(shards):
        if gather_with_grad or (r == rank and not local_loss):
            parts.append(s)
        else:
            parts.append(s.detach())
    return torch.cat(parts, dim=0)


def clip_loss_rank(a_shards: Sequence[Sequence[torch.Tensor]], b_shards: Sequence[Sequence[torch.Tensor]],
                   scale, rank: int, local_loss: bool, gather_with_grad: bool) -> torch.Tensor:
    """The scalar that rank `rank` computes in ClipLoss.forward when world_size > 1.

    a_shards[r][i] is feature tensor i of the image-like list on rank r (same for
    b_shards).  Gradients of the *sum over ranks* of these scalars w.r.t. the
    shard tensors reproduce what autograd + torch.distributed.nn.all_gather give
    (the all_gather backward is a sum over ranks); for gather_with_grad=False the
    per-rank scalar's own gradient is what that rank sees.  loss.py:103-142."""
    world = len(a_shards)
    n_a, n_b = len(a_shards[0]), len(b_shards[0])
    acc = 0
    for i in range(n_a):
        for j in range(n_b):
            a_loc, b_loc = a_shards[rank][i], b_shards[rank][j]
            a_all = _gathered_view([a_shards[r][i] for r in range(world)], rank, local_loss, gather_with_grad)
            b_all = _gathered_view([b_shards[r][j] for r in range(world)], rank, local_loss, gather_with_grad)
            if local_loss:
                lab = scale * a_loc @ b_all.T
                lba = scale * b_loc @ a_all.T
                first = a_loc.shape[0] * rank
            else:
                lab = scale * a_all @ b_all.T
                lba = lab.T
                first = 0
            acc = acc + symmetric_infonce(lab, lba, first)
    return acc / (n_a * n_b)


def cosmos_loss_rank(shards: Sequence[dict], rank: int, local_loss: bool, gather_with_grad: bool) -> Dict[str, torch.Tensor]:
    """COSMOSLoss.forward as seen by one rank of a world_size = len(shards) job.
    Each shards[r] is a dict with keys s_image, s_text, t_image, t_text, s_img_x,
    s_txt_x (lists of [b, D] tensors), logit_scale and distill_logit_scale."""
    me = shards[rank]
    dscale = me.get("distill_logit_scale")
    if dscale is None:
        dscale = me["logit_scale"]

    def side(key, detach=False, sl=slice(None)):
        out = []
        for s in shards:
            lst = list(s[key])[sl]
            out.append([t.detach() for t in lst] if detach else lst)
        return out

    t_img, t_txt = side("t_image", True), side("t_text", True)
    s_ix, s_tx = side("s_img_x"), side("s_txt_x")
    kw = dict(rank=rank, local_loss=local_loss, gather_with_grad=gather_with_grad)
    distill = (clip_loss_rank(s_ix, t_img, dscale, **kw) + clip_loss_rank(s_ix, t_txt, dscale, **kw)
               + clip_loss_rank(s_tx, t_img, dscale, **kw) + clip_loss_rank(s_tx, t_txt, dscale, **kw)) / 4
    clip = clip_loss_rank(side("s_image", sl=slice(0, 2)), side("s_text"), me["logit_scale"], **kw)
    return {"distill_loss": distill, "clip_loss": clip}


# ----------------------------------------------------------------------------
# Closed-form gradients of one pair (documents the math the CUDA kernels use)
# ----------------------------------------------------------------------------

def pair_closed_form(a: torch.Tensor, b: torch.Tensor, scale: float):
    """Loss and analytic gradients of one symmetric-InfoNCE pair (single process).

    With S = scale * a b^T, R = softmax over rows, C = softmax over columns,
    n = rows:   dL/dS = (R + C - 2 I) / (2 n),
                dL/da = scale * dS b,  dL/db = scale * dS^T a,
                dL/dscale = sum(dS * (a b^T)).
    Returns (loss, da, db, dscale) in the dtype of `a`."""
    n = a.shape[0]
    raw = a @ b.T
    S = scale * raw
    row_lse = torch.logsumexp(S, dim=1)
    col_lse = torch.logsumexp(S, dim=0)
    diag = torch.diagonal(S)
    loss = 0.5 * ((row_lse - diag).mean() + (col_lse - diag).mean())
    dS = (torch.exp(S - row_lse[:, None]) + torch.exp(S - col_lse[None, :])) / (2 * n)
    dS = dS - torch.eye(n, dtype=S.dtype) / n
    return loss, scale * dS @ b, scale * dS.T @ a, (dS * raw).sum()


# ----------------------------------------------------------------------------
# Cross-attention pooler
# ----------------------------------------------------------------------------

def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """transformer.py:24-30 (LayerNorm over the last dim, eps = nn.LayerNorm default)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def cross_pool(tokens: torch.Tensor, query: torch.Tensor, p: Dict[str, torch.Tensor], n_head: int,
               add_zero_attn: bool = False) -> torch.Tensor:
    """AttentionalCrossPooler.forward (transformer.py:225-230) written out.

    tokens [B, L, C] are the keys/values, query [B, Lq, d] the queries,
    p holds the module's parameters by their reference names:
      ln_q.weight/bias, ln_k.weight/bias, attn.in_proj_weight [3d, d],
      attn.in_proj_bias [3d], attn.out_proj.weight [d, d], attn.out_proj.bias.
    No key-padding mask, no dropout, softmax scale 1/sqrt(head_dim)
    (nn.MultiheadAttention defaults with kdim == vdim == d)."""
    B, L, C = tokens.shape
    Lq, d = query.shape[1], query.shape[2]
    hd = d // n_head
    kx = layer_norm(tokens, p["ln_k.weight"], p["ln_k.bias"])
    qx = layer_norm(query, p["ln_q.weight"], p["ln_q.bias"])
    W, bias = p["attn.in_proj_weight"], p["attn.in_proj_bias"]
    q = qx @ W[:d].T + bias[:d]
    k = kx @ W[d:2 * d].T + bias[d:2 * d]
    v = kx @ W[2 * d:].T + bias[2 * d:]
    q = q.view(B, Lq, n_head, hd).transpose(1, 2)          # B h Lq hd
    k = k.view(B, L, n_head, hd).transpose(1, 2)
    v = v.view(B, L, n_head, hd).transpose(1, 2)
    if add_zero_attn:      # transformer.py:221 -> F.multi_head_attention_forward: one all-zero key / value per head, after the projection
        k = torch.cat([k, k.new_zeros(B, n_head, 1, hd)], dim=2)
        v = torch.cat([v, v.new_zeros(B, n_head, 1, hd)], dim=2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, Lq, d)
    return o @ p["attn.out_proj.weight"].T + p["attn.out_proj.bias"]


def cosmos_crossmodal(features: torch.Tensor, tokens: torch.Tensor, p: Dict[str, torch.Tensor], n_head: int,
                      batch_size: int) -> torch.Tensor:
    """The cross-modal block of CLIP.forward for one direction (model.py:375-380
    for images, 382-384 for text): the tokens of the first `batch_size` samples
    are repeated for every crop, each crop's un-normalised feature is the single
    query, and the pooled token is added back before L2 normalisation."""
    n = features.shape[0] // batch_size
    rep = tokens[:batch_size].repeat(n, 1, 1)
    pooled = cross_pool(rep, features.unsqueeze(1), p, n_head)
    return F.normalize(features + pooled.squeeze(1), dim=-1)


# ----------------------------------------------------------------------------
# EMA
# ----------------------------------------------------------------------------

def ema_update_(teacher: Sequence[torch.Tensor], student: Sequence[torch.Tensor], momentum: float) -> None:
    """train.py:200-203: k <- k * m + (1 - m) * q, in place, parameter by parameter.
    (1 - m) is evaluated in Python double precision and multiplies q in q's dtype,
    exactly as the reference expression does."""
    with torch.no_grad():
        for k, q in zip(teacher, student):
            k.mul_(momentum).add_((1 - momentum) * q)


# ----------------------------------------------------------------------------
# Synthetic inputs (shared by tests, smoke and bench so that every leg sees the
# same tensors; SURVEY.md §8(d) "Synthetic inputs")
# ----------------------------------------------------------------------------

def clamp_logit_scales_(scalars: Sequence[torch.Tensor], lo: float = 0.0, hi: float = math.log(100)) -> None:
    """train.py:237-243: every logit scale (student and teacher, clip and distill) is clamped in place to [0, ln 100]."""
    with torch.no_grad():
        for t in scalars:
            t.clamp_(lo, hi)


# ----------------------------------------------------------------------------
# Retrieval metrics of the evaluation (outside the training step)
# ----------------------------------------------------------------------------

def ranks_from_scores(scores: torch.Tensor, gt: Sequence[Sequence[int]]) -> torch.Tensor:
    """0-based position of the best ground-truth item of every row in a descending sort of that row
    (train.py:722-731 for several items per row, 745-748 / 776-777 for one)."""
    out = torch.zeros(scores.shape[0], dtype=torch.long)
    for r, row in enumerate(scores):
        order = torch.argsort(row, descending=True)
        pos = torch.empty_like(order)
        pos[order] = torch.arange(order.numel())
        out[r] = min(int(pos[t]) for t in gt[r])
    return out


def _report(ranks: torch.Tensor, name: str, float32_mean: bool) -> Dict[str, float]:
    import numpy as np
    r = ranks.numpy()
    mean = float(ranks.float().mean().item()) if float32_mean else float(r.mean())
    d = {f"{name}_mean_rank": mean + 1, f"{name}_median_rank": float(np.floor(np.median(r)) + 1)}
    for k in (1, 5, 10):
        d[f"{name}_R@{k}"] = float(np.mean(r < k))
    return d


def get_clip_metrics(image_features: torch.Tensor, text_features: torch.Tensor, logit_scale) -> Dict[str, float]:
    """train.py:766-785: paired features, item r of one side matches item r of the other."""
    per_image = (logit_scale * image_features @ text_features.t()).detach().cpu()
    n = per_image.shape[0]
    gt = [[r] for r in range(n)]
    out = {}
    out.update(_report(ranks_from_scores(per_image, gt), "image_to_text", False))
    out.update(_report(ranks_from_scores(per_image.t(), gt), "text_to_image", False))
    return out


def compute_retrieval(similarity: torch.Tensor, txt2img: Dict[int, int], img2txt: Dict[int, Sequence[int]]) -> Dict[str, float]:
    """train.py:712-763: similarity is [images, captions]; an image has several captions (best one counts),
    a caption one image.  Ranks are held in a float32 tensor there, so the mean is a float32 mean."""
    n_img, n_txt = similarity.shape
    out = {}
    out.update(_report(ranks_from_scores(similarity.t(), [[txt2img[c]] for c in range(n_txt)]), "text_to_image", True))
    out.update(_report(ranks_from_scores(similarity, [list(img2txt[i]) for i in range(n_img)]), "image_to_text", True))
    return out


def make_retrieval_case(n_img: int, caps_per_img: int, dim: int, seed: int, noise: float = 1.5, shuffle: bool = True):
    """Seeded eval-like features: every image has `caps_per_img` captions scattered over the caption list."""
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(n_img, dim, generator=g)
    img = F.normalize(z + noise * torch.randn(n_img, dim, generator=g), dim=-1)
    n_txt = n_img * caps_per_img
    owner = torch.arange(n_img).repeat_interleave(caps_per_img)
    if shuffle:
        owner = owner[torch.randperm(n_txt, generator=g)]
    txt = F.normalize(z[owner] + noise * torch.randn(n_txt, dim, generator=g), dim=-1)
    txt2img = {c: int(owner[c]) for c in range(n_txt)}
    img2txt = {i: [c for c in range(n_txt) if txt2img[c] == i] for i in range(n_img)}
    return img, txt, txt2img, img2txt


def make_features(batch: int, dim: int, seed: int, n_img: int = 8, n_txt: int = 8, correlated: bool = True,
                  dtype=torch.float32, noise: float = 2.0) -> dict:
    """Unit-norm embeddings with the COSMOS list structure.  `correlated` draws a
    shared latent per sample; positives have cosine ~1/(1+noise^2): 0.2 by default (a
    mid-training softmax), 0.8 with noise=0.5 (peaked softmax, loss near zero)."""
    g = torch.Generator().manual_seed(seed)

    def view(z):
        x = torch.randn(batch, dim, generator=g)
        if z is not None:
            x = z + noise * x
        return F.normalize(x, dim=-1).to(dtype)

    z = torch.randn(batch, dim, generator=g) if correlated else None
    mk = lambda n: [view(z) for _ in range(n)]
    return {
        "s_image": mk(n_img), "s_text": mk(n_txt), "s_img_x": mk(n_img), "s_txt_x": mk(n_txt),
        "t_image": mk(2), "t_text": mk(2),
    }


def make_pooler_case(d: int, L: int, batch: int, n: int, seed: int):
    """Seeded parameters (reference state_dict names) and inputs for one pooler case:
    returns (params, tokens [batch, L, d], feats [n*batch, d], w [n*batch, d])."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, scale=1.0: torch.randn(*s, generator=g) * scale
    params = {
        "attn.in_proj_weight": r(3 * d, d, scale=d ** -0.5),
        "attn.in_proj_bias": r(3 * d, scale=0.05),
        "attn.out_proj.weight": r(d, d, scale=d ** -0.5),
        "attn.out_proj.bias": r(d, scale=0.05),
        "ln_q.weight": 1 + r(d, scale=0.1), "ln_q.bias": r(d, scale=0.1),
        "ln_k.weight": 1 + r(d, scale=0.1), "ln_k.bias": r(d, scale=0.1),
    }
    return params, r(batch, L, d), r(n * batch, d), r(n * batch, d)

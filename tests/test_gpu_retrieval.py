"""GPU: retrieval ranks (csrc/retrieval.cu) against the oracle and the reference's own eval metrics
(src/training/train.py:712-763 compute_retrieval, 766-785 get_clip_metrics; fixtures from tests/golden/make_golden.py)."""
import os

import pytest
import torch

from oracle import cosmos_oracle as O

pytestmark = pytest.mark.gpu

EPS = 2e-6   # scores closer than this to the threshold may legitimately fall on either side (fp32 summation order)


def _bounds(q, g, gt_lists):
    """Per row: the rank with near-ties resolved against / in favour of the ground truth (fp64 scores)."""
    s = q.double() @ g.double().t()
    lo, hi = [], []
    for r, items in enumerate(gt_lists):
        thr = s[r, items].max()
        lo.append(int((s[r] > thr + EPS).sum()))
        hi.append(int((s[r] > thr - EPS).sum()) - sum(1 for t in items if abs(float(s[r, t] - thr)) <= EPS))
    return torch.tensor(lo), torch.tensor(hi)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,D", [(1, 1, 1), (5, 7, 3), (128, 128, 16), (130, 257, 100), (300, 1000, 512), (64, 4099, 77)])
def test_paired_and_csr_ranks_match_oracle(dtype, M, N, D):
    from cosmos_b200 import retrieval_ranks
    g = torch.Generator().manual_seed(M * 7 + N + D)
    unit = torch.nn.functional.normalize      # unit-norm rows: scores in [-1, 1], fp32 summation error far below EPS
    q = unit(torch.randn(M, D, generator=g), dim=-1).to(dtype)
    gal = unit(torch.randn(N, D, generator=g), dim=-1).to(dtype)
    if N >= M:                      # paired: item r for query r; every other positive is an exact copy of its query
        gal[:M:2] = q[::2]
        got = retrieval_ranks(q.cuda(), gal.cuda()).cpu()
        lo, hi = _bounds(q.float(), gal.float(), [[r] for r in range(M)])
        assert got.dtype == torch.int32 and got.shape == (M,)
        assert bool(((got >= lo) & (got <= hi)).all()), (got - lo).abs().max()
    # several ground-truth items per row, scattered; one row without any (rank = gallery size)
    lists = [sorted(set(torch.randint(0, N, (1 + r % 4,), generator=g).tolist())) for r in range(M)]
    lists[-1] = [] if M > 1 else lists[-1]
    off = torch.tensor([0] + torch.tensor([len(x) for x in lists]).cumsum(0).tolist(), dtype=torch.int32)
    idx = torch.tensor([t for x in lists for t in x], dtype=torch.int32)
    got = retrieval_ranks(q.cuda(), gal.cuda(), off.cuda(), idx.cuda()).cpu()
    keep = [r for r, x in enumerate(lists) if x]
    lo, hi = _bounds(q.float()[keep], gal.float(), [lists[r] for r in keep])
    assert bool(((got[keep] >= lo) & (got[keep] <= hi)).all())
    if M > 1:
        assert int(got[-1]) == N
    # contiguous ranges of two items (gt_index omitted); rows past the end of the gallery have none
    rng_off = (torch.arange(M + 1) * 2).clamp(max=N).to(torch.int32)
    got = retrieval_ranks(q.cuda(), gal.cuda(), rng_off.cuda()).cpu()
    lists = [list(range(int(rng_off[r]), int(rng_off[r + 1]))) for r in range(M)]
    keep = [r for r, x in enumerate(lists) if x]
    lo, hi = _bounds(q.float()[keep], gal.float(), [lists[r] for r in keep])
    assert bool(((got[keep] >= lo) & (got[keep] <= hi)).all())


def test_strided_rows_and_exact_self_match():
    """Row strides larger than D (views of a wider buffer); a gallery that contains the query itself scores it exactly
    like the threshold kernel does, so the item is never counted against itself: rank 0 for unit-norm duplicates."""
    from cosmos_b200 import retrieval_ranks
    g = torch.Generator().manual_seed(5)
    wide_q = torch.nn.functional.normalize(torch.randn(257, 640, generator=g), dim=-1).cuda()
    q = wide_q[:, 64:576]                                   # stride 640, D = 512
    assert q.stride(0) == 640
    gal = torch.cat([q.clone(), torch.nn.functional.normalize(torch.randn(1000, 512, generator=g), dim=-1).cuda() * 0.9])
    got = retrieval_ranks(q, gal)
    assert int(got.abs().max()) == 0


def test_metrics_match_reference_fixture(golden_dir):
    from cosmos_b200.retrieval import compute_retrieval, get_clip_metrics
    for rec in torch.load(os.path.join(golden_dir, "retrieval.pt"), weights_only=False):
        img, txt, txt2img, img2txt = O.make_retrieval_case(rec["n_img"], rec["caps"], rec["dim"], rec["seed"], rec["noise"])
        got = compute_retrieval(img.cuda(), txt.cuda(), txt2img, img2txt)
        assert list(got.keys()) == list(rec["compute_retrieval"].keys())
        for k, v in rec["compute_retrieval"].items():
            assert float(got[k]) == pytest.approx(v, rel=1e-6, abs=0), k
        if "get_clip_metrics" in rec:
            img_p, txt_p, _, _ = O.make_retrieval_case(rec["n_img"], 1, rec["dim"], rec["seed"], rec["noise"], shuffle=False)
            got = get_clip_metrics(img_p.cuda(), txt_p.cuda(), torch.tensor(14.2857))
            for k, v in rec["get_clip_metrics"].items():
                assert float(got[k]) == pytest.approx(v, rel=1e-12, abs=0), k


def test_full_size_properties():
    """N = 32768 paired features, dim 512 (no oracle at this size): the ranks do not depend on the order of the gallery,
    a query's rank against a gallery that lacks every better item is 0, and sum over rows of rank equals the number of
    (row, column) pairs above the thresholds counted by an independent fp32 matmul on the device (up to near-ties)."""
    from cosmos_b200 import retrieval_ranks
    n, d = 32768, 512
    g = torch.Generator(device="cuda").manual_seed(11)
    z = torch.randn(n, d, generator=g, device="cuda")
    img = torch.nn.functional.normalize(z + 2.0 * torch.randn(n, d, generator=g, device="cuda"), dim=-1)
    txt = torch.nn.functional.normalize(z + 2.0 * torch.randn(n, d, generator=g, device="cuda"), dim=-1)
    ranks = retrieval_ranks(img, txt)
    assert torch.equal(ranks, retrieval_ranks(img, txt))                         # integer counts: run-to-run identical
    perm = torch.randperm(n, generator=g, device="cuda")
    off = torch.arange(n + 1, dtype=torch.int32, device="cuda")
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n, device="cuda")
    ranks_p = retrieval_ranks(img, txt[perm], off, inv.to(torch.int32))           # same items, shuffled gallery
    assert torch.equal(ranks, ranks_p)
    lo = torch.zeros(n, dtype=torch.long, device="cuda")
    hi = torch.zeros(n, dtype=torch.long, device="cuda")
    thr = (img * txt).sum(-1)
    for c0 in range(0, n, 4096):
        s = img @ txt[c0:c0 + 4096].t()
        lo += (s > (thr + 1e-5)[:, None]).sum(1)
        hi += (s > (thr - 1e-5)[:, None]).sum(1)
    r64 = ranks.long()
    assert bool(((r64 >= lo) & (r64 <= hi)).all())
    assert 0.0 < float((ranks == 0).float().mean()) < 1.0


def _stable_ranks(scores: torch.Tensor, gt) -> torch.Tensor:
    """Position of the best ground-truth item of every row in torch's STABLE descending sort (NaN first, ties by index)."""
    out = torch.zeros(scores.shape[0], dtype=torch.long)
    for r, row in enumerate(scores):
        order = torch.argsort(row, descending=True, stable=True)
        pos = torch.empty_like(order)
        pos[order] = torch.arange(order.numel())
        out[r] = min(int(pos[t]) for t in gt[r])
    return out


@pytest.mark.gpu
def test_ties_and_nan_rank_like_a_stable_sort():
    """A collapsed model (every score ties) or a diverged one (NaN features) must not report R@1 = 1: tied items with a
    lower index come first and NaN sorts above every number, as in torch.sort (which the reference's argsort,
    src/training/train.py:722-748, 776-777, goes through)."""
    from cosmos_b200.retrieval import get_clip_metrics, retrieval_ranks
    g = torch.Generator().manual_seed(3)
    M, D = 300, 64
    # (1) collapsed: all rows identical -> all scores tie -> item r sits at position r
    one = torch.randn(1, D, generator=g)
    q = one.expand(M, D).contiguous().cuda()
    ranks = retrieval_ranks(q, q.clone()).cpu().long()
    assert torch.equal(ranks, torch.arange(M))
    m = get_clip_metrics(q, q.clone(), 1.0)
    assert m["image_to_text_R@1"] == pytest.approx(1.0 / M) and m["image_to_text_mean_rank"] == pytest.approx((M - 1) / 2 + 1)
    # (2) duplicated gallery rows: integer-valued features make the duplicates' scores exactly equal
    qi = torch.randint(-3, 4, (M, D), generator=g).float()
    gi = torch.randint(-3, 4, (M, D), generator=g).float()
    gi[1::2] = gi[0::2]                                  # every odd row duplicates the even row before it
    gt = [[r] for r in range(M)]
    want = _stable_ranks(qi @ gi.t(), gt)
    got = retrieval_ranks(qi.cuda(), gi.cuda()).cpu().long()
    assert torch.equal(got, want)
    assert bool((want[1::2] >= 1).all())                 # the duplicate with the higher index never ranks first
    # several ground-truth items per row, some tied with each other and with distractors
    gts = [[(3 * r) % M, (3 * r + 1) % M, (7 * r + 2) % M] for r in range(M)]
    off = torch.arange(0, 3 * M + 1, 3, dtype=torch.int32).cuda()
    idx = torch.tensor([t for row in gts for t in row], dtype=torch.int32).cuda()
    got = retrieval_ranks(qi.cuda(), gi.cuda(), off, idx).cpu().long()
    assert torch.equal(got, _stable_ranks(qi @ gi.t(), gts))
    # (3) NaN: a NaN query row ties everything (position = own index); a NaN gallery row outranks every number
    qn, gn = qi.clone(), gi.clone()
    qn[5, 0] = float("nan")
    gn[17, 3] = float("nan")
    gn[200] = float("nan")
    want = _stable_ranks(qn @ gn.t(), gt)
    got = retrieval_ranks(qn.cuda(), gn.cuda()).cpu().long()
    assert torch.equal(got, want)
    assert int(got[5]) == 5 and int(got[17]) == 0 and int(got[200]) == 1 and int(got[0]) >= 2

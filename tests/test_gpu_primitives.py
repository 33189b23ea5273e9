"""GPU: tcgen05/TMA primitive self-test binary and the EMA kernel (bit-exact)."""
import os
import subprocess

import pytest
import torch

from oracle import cosmos_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_tcgen05_primitives_selftest():
    exe = os.path.join(ROOT, "cosmos_b200", "selftest_sm100")
    assert os.path.exists(exe), "run python -m cosmos_b200.build"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr)
    lines = {l.split(":")[0]: l for l in r.stdout.splitlines() if l.startswith("case")}
    # the product kernels rely on cases 0 and 1 (K-major SS GEMM, MN-major B per slab)
    assert "PASS" in lines.get("case 0", ""), r.stdout
    assert "PASS" in lines.get("case 1", ""), r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_ema_bit_exact(dtype):
    from cosmos_b200 import ema_update_
    g = torch.Generator().manual_seed(1)
    shapes = [(1,), (), (7,), (33, 5), (8192,), (8193,), (3, 8192), (100003,), (512, 768), (2_000_001,)]
    student = [(torch.randn(s, generator=g) * 0.02).to(dtype).cuda() for s in shapes]
    teacher = [(torch.randn(s, generator=g) * 0.02).to(dtype).cuda() for s in shapes]
    # an unaligned view (storage offset of one element)
    base_k = (torch.randn(1001, generator=g) * 0.02).to(dtype).cuda()
    base_q = (torch.randn(1001, generator=g) * 0.02).to(dtype).cuda()
    student.append(base_q[1:])
    teacher.append(base_k[1:])
    for m in (0.99, 0.999, 0.5, 0.0, 1.0):
        want = [t.clone() for t in teacher]
        O.ema_update_(want, student, m)              # the literal reference loop, on the GPU tensors
        got = [t for t in teacher]
        ptrs = [t.data_ptr() for t in got]
        ema_update_(student, got, m)
        torch.cuda.synchronize()
        for a, b, p in zip(got, want, ptrs):
            assert a.data_ptr() == p                  # in place, same storage
            assert torch.equal(a, b)


@pytest.mark.gpu
def test_ema_mixed_dtypes_and_moved_storage():
    """Mixed-precision parameter sets (the reference's --precision bf16/fp16 keeps norms, embeddings and the logit
    scales fp32 next to 16-bit weights, src/open_clip/model.py:156-157): every (dtype, device) bucket must chunk with its
    own element size, incl. tensors over one 8192-element chunk.  Then a parameter whose storage is re-allocated between
    calls must be picked up on the very next call (the cached table holds raw pointers)."""
    from cosmos_b200 import ema_update_
    g = torch.Generator().manual_seed(5)
    spec = [((20000,), torch.float32), ((300, 77), torch.bfloat16), ((9001,), torch.float32), ((8193, 3), torch.float16),
            ((), torch.float32), ((40000,), torch.bfloat16), ((5,), torch.float16), ((3, 8192), torch.float32)]
    for order in (spec, spec[::-1]):          # either dtype may be the "last teacher parameter"
        student = [(torch.randn(s, generator=g) * 0.02).to(dt).cuda() for s, dt in order]
        teacher = [(torch.randn(s, generator=g) * 0.02).to(dt).cuda() for s, dt in order]
        guard = [torch.full((64,), 7.0, device="cuda") for _ in range(4)]        # neighbours an overrun would hit
        for m in (0.99, 0.5):
            want = [t.clone() for t in teacher]
            O.ema_update_(want, student, m)
            ema_update_(student, teacher, m)
            torch.cuda.synchronize()
            for a, b in zip(teacher, want):
                assert torch.equal(a, b)
        assert all(bool((t == 7.0).all()) for t in guard)
    # moved storage: same Parameter objects, new memory (p.data = ...)
    ps = [torch.nn.Parameter(torch.randn(20000, generator=g).cuda()) for _ in range(5)]
    pt = [torch.nn.Parameter(torch.randn(20000, generator=g).cuda(), requires_grad=False) for _ in range(5)]
    ema_update_(ps, pt, 0.9)
    old = pt[1].data
    old_copy = old.clone()
    pt[1].data = old.clone()                    # parameter 1 (neither first, middle nor last) now lives elsewhere
    ps[3].data = ps[3].data.clone() * 2
    want = [t.detach().clone() for t in pt]
    O.ema_update_(want, [p.detach() for p in ps], 0.9)
    ema_update_(ps, pt, 0.9)
    torch.cuda.synchronize()
    for a, b in zip(pt, want):
        assert torch.equal(a.detach(), b)
    assert torch.equal(old, old_copy)            # the abandoned storage was not written through a stale pointer


@pytest.mark.gpu
def test_ema_golden(golden_dir):
    from cosmos_b200 import ema_update_
    rec = torch.load(os.path.join(golden_dir, "ema.pt"), weights_only=False)
    for m, want in rec["outs"].items():
        k = [t.clone().cuda() for t in rec["teacher"]]
        q = [t.cuda() for t in rec["student"]]
        ema_update_(q, k, m)
        for a, b in zip(k, want):
            assert torch.equal(a.cpu(), b)


@pytest.mark.gpu
def test_ema_module_api():
    from cosmos_b200 import ema_update_
    torch.manual_seed(0)
    student = torch.nn.Sequential(torch.nn.Linear(300, 77), torch.nn.LayerNorm(77), torch.nn.Linear(77, 5)).cuda()
    import copy
    teacher = copy.deepcopy(student)
    for p in teacher.parameters():
        p.requires_grad = False
    with torch.no_grad():
        for p in student.parameters():
            p.add_(torch.randn_like(p))
    ref = copy.deepcopy(teacher)
    for _ in range(3):
        O.ema_update_(list(ref.parameters()), list(student.parameters()), 0.99)
        ema_update_(student, teacher, 0.99)
    for a, b in zip(teacher.parameters(), ref.parameters()):
        assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_clamp_logit_scales_bit_exact(dtype):
    """The clamps of src/training/train.py:237-243 in one launch: bit-identical to four clamp_ calls, NaN kept."""
    import math
    from cosmos_b200 import clamp_logit_scales_

    class M(torch.nn.Module):
        def __init__(self, a, b):
            super().__init__()
            self.logit_scale = torch.nn.Parameter(torch.tensor(a, dtype=dtype))
            self.distill_logit_scale = torch.nn.Parameter(torch.tensor(b, dtype=dtype))

    class Wrapped(torch.nn.Module):      # stands in for DistributedDataParallel (train.py: unwrap_model)
        def __init__(self, m):
            super().__init__()
            self.module = m

    for vals in ((4.7, -0.3, 2.6593, 4.60517), (float("nan"), 100.0, 0.0, -0.0), (math.log(100), 4.6051702, 4.605171, 1e-30)):
        student, teacher = M(vals[0], vals[1]).cuda(), M(vals[2], vals[3]).cuda()
        want = [p.detach().clone() for m in (student, teacher) for p in (m.logit_scale, m.distill_logit_scale)]
        for w in want:
            w.clamp_(0, math.log(100))
        clamp_logit_scales_(Wrapped(student), teacher)
        torch.cuda.synchronize()
        got = [p.detach() for m in (student, teacher) for p in (m.logit_scale, m.distill_logit_scale)]
        for g, w in zip(got, want):
            assert torch.equal(g.view(torch.int16 if dtype != torch.float32 else torch.int32),
                               w.view(torch.int16 if dtype != torch.float32 else torch.int32)), (vals, g, w)
    # plain tensors, more than one launch's worth, model without a distill scale
    ts = [torch.tensor(float(i) - 3.0, device="cuda") for i in range(11)]
    clamp_logit_scales_(ts, None, 0.0, 4.0)
    assert [float(t) for t in ts] == [min(max(float(i) - 3.0, 0.0), 4.0) for i in range(11)]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        clamp_logit_scales_([torch.tensor(1.0)])


@pytest.mark.gpu
def test_clamp_golden(golden_dir):
    from cosmos_b200 import clamp_logit_scales_
    rec = torch.load(os.path.join(golden_dir, "clamp.pt"), weights_only=False)
    for name, want in rec["outs"].items():
        ts = [torch.tensor(v, dtype=want.dtype, device="cuda") for v in rec["values"]]
        clamp_logit_scales_(ts)
        got = torch.stack(ts).cpu()
        assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(want, nan=-7.0)), name
        assert torch.isnan(got).tolist() == torch.isnan(want).tolist()

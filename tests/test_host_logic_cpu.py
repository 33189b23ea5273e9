"""CPU tests of the host-side logic of cosmos_b200 (no kernels): the loss API composition on one
process and the four (local_loss, gather_with_grad) modes on gloo ranks, both against the fixtures
the unmodified reference produced (tests/golden/make_golden.py).  Kernel entry points are replaced by
the documented-contract emulation in tests/emulation.py."""
import os
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _close(a, b, rtol=2e-4, atol=2e-6):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def _force_e_route(comm_free=True):
    """Stored-exponential route (forward keeps what the backward needs, gradients formed chunk by chunk in forward) in chunks
    of three row tensors, wherever the product would allow it (not in the local-loss modes)."""
    return lambda x_r, y_c, comm: 0 if (comm.distributed and comm.local_loss) else min(3, x_r.shape[0])


@pytest.mark.parametrize("keep_g", [False, True, "e"])
def test_cosmos_api_single_process(monkeypatch, keep_g):
    from tests import emulation
    emulation.install(monkeypatch)
    from cosmos_b200 import COSMOSLoss, infonce
    used = []
    if keep_g == "e":
        chunker = _force_e_route()
        monkeypatch.setattr(infonce, "_e_store_chunk", lambda *a: used.append(1) or chunker(*a))
    elif keep_g:      # column-side gradient through the stored G tiles + GEMM (the dim-512 route of the GPU path)
        monkeypatch.setattr(infonce, "_g_store_ok", lambda x_r, y_c: used.append(1) or True)
    for case in torch.load(os.path.join(GOLDEN, "cosmos_w1_small.pt"), weights_only=False):
        leaf = {k: [t.clone().requires_grad_(True) for t in v] for k, v in case["inputs"].items()}
        ls = torch.tensor(case["logit_scale"], requires_grad=True)
        ds = None if case["distill_logit_scale"] is None else torch.tensor(case["distill_logit_scale"], requires_grad=True)
        out = COSMOSLoss(cache_labels=True)(tuple(leaf["s_image"]), tuple(leaf["s_text"]), ls, t_image_features=leaf["t_image"],
                                            t_text_features=leaf["t_text"], output_dict=True, distill_logit_scale=ds,
                                            s_img_crossmodal_features=leaf["s_img_x"], s_txt_crossmodal_features=leaf["s_txt_x"])
        up = case["upstream"]
        (up[0] * out["distill_loss"] + up[1] * out["clip_loss"]).backward()
        for k in ("distill_loss", "clip_loss"):
            _close(out[k].detach(), case["out"][k], rtol=2e-5)
        for k, lst in case["grads"].items():
            for t, g in zip(leaf[k], lst):
                if g is None:
                    assert t.grad is None
                else:
                    _close(t.grad, g, rtol=2e-4, atol=2e-6 * max(up))
        _close(ls.grad, case["g_logit_scale"], rtol=2e-4, atol=1e-5 * max(up))
        if ds is not None:
            _close(ds.grad, case["g_distill_scale"], rtol=2e-4, atol=1e-5 * max(up))
        # sum form (output_dict=False) and chunk-view inputs (zero-copy stack path)
        total = COSMOSLoss()(leaf["s_image"], leaf["s_text"], ls, leaf["t_image"], leaf["t_text"], False, ds,
                             leaf["s_img_x"], leaf["s_txt_x"])
        _close(total.detach(), case["out"]["distill_loss"] + case["out"]["clip_loss"], rtol=2e-5)
    assert bool(used) == bool(keep_g)


def test_stack_views_zero_copy():
    from cosmos_b200.infonce import stack_views
    buf = torch.randn(8 * 16, 64).bfloat16()
    chunks = buf.chunk(8)                        # what train.py:171-182 hands the loss
    st = stack_views(chunks, torch.bfloat16)
    assert st.data_ptr() == buf.data_ptr() and st.shape == (8, 16, 64)
    assert torch.equal(st[3], chunks[3])
    st2 = stack_views(chunks[:2], torch.bfloat16)
    assert st2.data_ptr() == buf.data_ptr() and st2.shape == (2, 16, 64)
    sep = [c.clone() for c in chunks]
    st3 = stack_views(sep, torch.bfloat16)
    assert torch.equal(st3, torch.stack(sep))
    st4 = stack_views([c.float() for c in chunks], torch.bfloat16)
    assert st4.dtype == torch.bfloat16


def test_out_of_scope_names_import_and_raise():
    from cosmos_b200.loss import CoCaLoss, DistillClipLoss, SigLipLoss
    for cls in (CoCaLoss, DistillClipLoss, SigLipLoss):
        with pytest.raises(NotImplementedError):
            cls()


def test_product_rejects_cpu_tensors():
    from cosmos_b200 import ClipLoss, ema_update_
    a = torch.randn(8, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ClipLoss()(a, a, 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ema_update_([a], [a.clone()], 0.9)


def _rank_worker(rank, world, port, fname, tmpdir, keep_g=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import emulation
        emulation.install()
        from cosmos_b200 import ClipLoss, COSMOSLoss, infonce
        if keep_g == "e":
            infonce._e_store_chunk = _force_e_route()
        elif keep_g:      # non-local modes: column-side gradient = reduce-scatter of G^T x (stored G tiles) instead of a second sweep
            infonce._g_store_ok = lambda x_r, y_c: True
        rec = torch.load(os.path.join(GOLDEN, fname), weights_only=False)
        for name, spec in rec["payload"].items():
            ref = rec["results"][rank][name]
            mine = spec["shards"][rank]
            kw = dict(local_loss=spec["local_loss"], gather_with_grad=spec["gather_with_grad"], cache_labels=True,
                      rank=rank, world_size=world)
            if spec["kind"] == "clip":
                a = [t.clone().requires_grad_(True) for t in mine["a"]]
                b = [t.clone().requires_grad_(True) for t in mine["b"]]
                s = torch.tensor(spec["logit_scale"], requires_grad=True)
                val = ClipLoss(**kw)(a, b, s)
                val.backward()
                _close(val.detach(), ref["loss"], rtol=2e-5)
                for t, g in zip(a, ref["ga"]):
                    _close(t.grad, g)
                for t, g in zip(b, ref["gb"]):
                    _close(t.grad, g)
                _close(s.grad, ref["gscale"], rtol=2e-4, atol=2e-6)
            else:
                leaf = {k: [t.clone().requires_grad_(True) for t in v] for k, v in mine.items()}
                ls = torch.tensor(spec["logit_scale"], requires_grad=True)
                ds = torch.tensor(spec["distill_logit_scale"], requires_grad=True)
                out = COSMOSLoss(**kw)(leaf["s_image"], leaf["s_text"], ls, t_image_features=leaf["t_image"],
                                       t_text_features=leaf["t_text"], output_dict=True, distill_logit_scale=ds,
                                       s_img_crossmodal_features=leaf["s_img_x"], s_txt_crossmodal_features=leaf["s_txt_x"])
                (out["distill_loss"] + out["clip_loss"]).backward()
                for k in ("distill_loss", "clip_loss"):
                    _close(out[k].detach(), ref["out"][k], rtol=2e-5)
                for k, lst in ref["grads"].items():
                    for t, g in zip(leaf[k], lst):
                        if g is None:
                            assert t.grad is None, (name, k)
                        else:
                            _close(t.grad, g)
                _close(ls.grad, ref["g_logit_scale"], rtol=2e-4, atol=2e-6)
                _close(ds.grad, ref["g_distill_scale"], rtol=2e-4, atol=2e-6)
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world,fname,port,keep_g", [(2, "multirank_w2.pt", 29721, False), (4, "multirank_w4.pt", 29722, False),
                                                     (2, "multirank_w2.pt", 29723, True), (4, "multirank_w4.pt", 29724, True),
                                                     (2, "multirank_w2.pt", 29725, "e"), (4, "multirank_w4.pt", 29726, "e")])
def test_multirank_modes_gloo(world, fname, port, keep_g):
    ctx = mp.get_context("spawn")
    with tempfile.TemporaryDirectory() as tmpdir:
        procs = [ctx.Process(target=_rank_worker, args=(r, world, port, fname, tmpdir, keep_g)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
        for r, p in enumerate(procs):
            assert p.exitcode == 0, f"rank {r} failed"
            assert os.path.exists(os.path.join(tmpdir, f"ok{r}"))


def test_pooler_module_is_checkpoint_compatible():
    """Same parameter names/shapes as the reference AttentionalCrossPooler (transformer.py:210-230), so reference
    state_dicts load; unsupported configurations are rejected loudly; CPU tensors are rejected (no fallback)."""
    from cosmos_b200.pooler import AttentionalCrossPooler, crossmodal_features
    from oracle.cosmos_oracle import make_pooler_case
    params, tokens, feats, _ = make_pooler_case(64, 7, 3, 2, seed=1)
    mod = AttentionalCrossPooler(64, 64, 4)
    assert {k: tuple(v.shape) for k, v in mod.state_dict().items()} == {k: tuple(v.shape) for k, v in params.items()}
    mod.load_state_dict(params)
    assert AttentionalCrossPooler(64, 64, 4, add_zero_attn=True).add_zero_attn          # supported (key / value route)
    with pytest.raises(NotImplementedError):
        AttentionalCrossPooler(64, 32, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crossmodal_features(mod, tokens, feats, 3)


def test_c_abi_exports_every_declared_symbol():
    """Every function declared in include/cosmos_b200.h is exported by libcosmos_b200.so (no compute calls)."""
    import ctypes
    import re
    from cosmos_b200 import _lib
    lib = _lib.lib()
    header = open(os.path.join(ROOT, "include", "cosmos_b200.h")).read()
    names = set(re.findall(r"\b(cosmos_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert lib.cosmos_abi_version() == 3
    assert lib.cosmos_status_string(0) == b"ok"
    # and the other way round: the library (built with -fvisibility=hidden) exports nothing but what the header declares
    import shutil
    import subprocess
    if shutil.which("nm"):
        out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
        exported = {ln.split()[-1] for ln in out.splitlines() if ln.split() and ln.split()[-2] in ("T", "t", "W", "w", "D", "B")}
        exported = {n for n in exported if not n.startswith(("_init", "_fini", "_edata", "_end", "__bss_start"))}
        assert exported == names, (sorted(exported - names), sorted(names - exported))
    # host-only helpers behave
    numel = (ctypes.c_int64 * 3)(8192, 1, 8193)
    assert lib.cosmos_ema_table_entries(3, numel) == 1 + 1 + 2
    assert lib.cosmos_ema_table_entries(-1, numel) == -1
    # clamp launch: argument validation happens before any device call
    two = (ctypes.c_uint64 * 2)(256, 258)
    assert lib.cosmos_clamp_scalars(two, 9, 0.0, 1.0, 0, 0, None) == 1          # more than COSMOS_CLAMP_MAX
    assert lib.cosmos_clamp_scalars(two, 2, 0.0, 1.0, 0, 0, None) == 1          # fp32 scalar at a 2-byte-aligned address
    assert lib.cosmos_clamp_scalars(two, 2, 1.0, 0.0, 1, 0, None) == 1          # lo > hi
    assert lib.cosmos_clamp_scalars(two, 2, 0.0, 1.0, 7, 0, None) == 2          # unknown dtype
    assert lib.cosmos_clamp_scalars(None, 0, 0.0, 1.0, 0, 0, None) == 0         # nothing to do
    bad = _lib.InfoNceProblem(x=16, y=16, gx=1, gy=1, n_rows=8, n_cols=8, dim=100, label_offset=0, dtype=1, reserved=0, scale=4)
    assert lib.cosmos_infonce_workspace_bytes(ctypes.byref(bad)) == -1          # dim not a multiple of 64


def test_retrieval_host_logic_against_reference_fixture(monkeypatch):
    """compute_retrieval / get_clip_metrics of cosmos_b200.retrieval (CSR ground truth, metric expressions, key names)
    with the rank kernel replaced by its documented contract: ranks[r] = #{c : <q_r, g_c> > max_t <q_r, g_t>}."""
    from cosmos_b200 import retrieval
    from oracle import cosmos_oracle as O

    def ranks_contract(q, g, gt_offsets=None, gt_index=None):
        s = q.double() @ g.double().t()
        out = torch.empty(q.shape[0], dtype=torch.int32)
        for r in range(q.shape[0]):
            if gt_offsets is None:
                items = [r]
            else:
                span = range(int(gt_offsets[r]), int(gt_offsets[r + 1]))
                items = [int(gt_index[e]) for e in span] if gt_index is not None else list(span)
            out[r] = int((s[r] > s[r, items].max()).sum())
        return out

    monkeypatch.setattr(retrieval, "retrieval_ranks", ranks_contract)
    for rec in torch.load(os.path.join(GOLDEN, "retrieval.pt"), weights_only=False):
        img, txt, txt2img, img2txt = O.make_retrieval_case(rec["n_img"], rec["caps"], rec["dim"], rec["seed"], rec["noise"])
        got = retrieval.compute_retrieval(img, txt, txt2img, img2txt)
        assert list(got.keys()) == list(rec["compute_retrieval"].keys())
        for k, v in rec["compute_retrieval"].items():
            assert float(got[k]) == pytest.approx(v, rel=1e-6, abs=0), k
        if "get_clip_metrics" in rec:
            img_p, txt_p, _, _ = O.make_retrieval_case(rec["n_img"], 1, rec["dim"], rec["seed"], rec["noise"], shuffle=False)
            got = retrieval.get_clip_metrics(img_p, txt_p, torch.tensor(14.2857))
            assert list(got.keys()) == list(rec["get_clip_metrics"].keys())
            for k, v in rec["get_clip_metrics"].items():
                assert float(got[k]) == pytest.approx(v, rel=1e-12, abs=0), k
    with pytest.raises(RuntimeError, match="outside the gallery"):
        retrieval.compute_retrieval(img, txt, {c: 10 ** 6 for c in txt2img}, img2txt)


def test_retrieval_rejects_cpu_tensors():
    from cosmos_b200 import retrieval_ranks
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        retrieval_ranks(torch.randn(4, 8), torch.randn(5, 8))


def test_stored_exponential_chunking(monkeypatch):
    """_e_store_chunk: route selection and equal passes (shapes only; meta tensors)."""
    from cosmos_b200 import infonce
    from cosmos_b200.infonce import Comm, _e_store_chunk

    def stacks(n_r, b, n_c, n_all, dim=512):
        return (torch.empty(n_r, b, dim, dtype=torch.bfloat16, device="meta"),
                torch.empty(n_c, n_all, dim, dtype=torch.bfloat16, device="meta"))

    agreed = []
    # ranks agree on the smallest chunk with one MIN all-reduce per shape; here: a rank whose peers can only take 5 at a time
    monkeypatch.setattr(infonce.dist, "all_reduce", lambda t, op=None, group=None: (agreed.append(int(t)), t.clamp_(max=5))[1])
    monkeypatch.setattr(infonce, "_e_chunk_cache", {})
    monkeypatch.setattr(infonce, "_E_STORE_MAX_BYTES", 40 << 30)
    monkeypatch.setattr(infonce, "_E_STORE_MIN_BYTES", 1 << 28)
    assert _e_store_chunk(*stacks(16, 32768, 4, 32768), Comm()) == 4        # 8.6 GB per row tensor: 4 + 4 + 4 + 4
    assert _e_store_chunk(*stacks(8, 32768, 2, 32768), Comm()) == 8         # CLIP group, one GPU: one pass
    assert _e_store_chunk(*stacks(16, 4096, 4, 32768), Comm()) == 16        # the per-rank problem of an 8-GPU job: one pass
    assert _e_store_chunk(*stacks(16, 4096, 4, 32768), Comm(rank=1, world_size=8)) == 5 and agreed == [16]
    assert _e_store_chunk(*stacks(16, 4096, 4, 32768), Comm(rank=1, world_size=8)) == 5 and agreed == [16]   # cached: no second collective
    assert _e_store_chunk(*stacks(16, 4096, 4, 32768), Comm(rank=1, world_size=8, local_loss=True)) == 0
    assert _e_store_chunk(*stacks(16, 4096, 4, 32768, dim=256), Comm()) == 0                 # other widths: recompute kernels
    assert _e_store_chunk(*stacks(8, 256, 2, 256), Comm()) == 0                              # tiny: one launch per group wins
    monkeypatch.setattr(infonce, "_E_STORE_MAX_BYTES", 30 << 30)
    monkeypatch.setattr(infonce, "_e_chunk_cache", {})
    assert _e_store_chunk(*stacks(16, 32768, 4, 32768), Comm()) == 3        # at most 3 fit: 6 passes of 3, 3, 3, 3, 3, 1
    monkeypatch.setattr(infonce, "_E_STORE_MAX_BYTES", 0)
    assert _e_store_chunk(*stacks(16, 32768, 4, 32768), Comm()) == 0        # COSMOS_B200_ESTORE_MAX_GB=0 switches the route off


def _chunk_agreement_worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cosmos_b200 import infonce
        infonce._E_STORE_MIN_BYTES = 0
        per_tensor = 2 * 128 * 256 * 2                      # n_c = 2, b = 100 -> 128, N = 200 -> 256
        infonce._E_STORE_MAX_BYTES = per_tensor * (14 if rank == 0 else 6)    # two chunks in flight: the ranks could take 7 and 3 row tensors at a time
        x = torch.empty(7, 100, 512, dtype=torch.bfloat16)
        y = torch.empty(2, 200, 512, dtype=torch.bfloat16)
        comm = infonce.Comm(rank=rank, world_size=world, group=dist.group.WORLD)
        got = infonce._e_store_chunk(x, y, comm)
        assert got == 3, got                                # 3 + 3 + 1 on BOTH ranks: same number of column-statistics all-reduces
        assert infonce._e_store_chunk(x, y, comm) == 3      # cached, no collective (a hang here would mean it was not)
        open(os.path.join(tmpdir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_stored_exponential_chunk_agreement_gloo():
    ctx = mp.get_context("spawn")
    with tempfile.TemporaryDirectory() as tmpdir:
        procs = [ctx.Process(target=_chunk_agreement_worker, args=(r, 2, 29741, tmpdir)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        for r, p in enumerate(procs):
            assert p.exitcode == 0, f"rank {r} failed"
            assert os.path.exists(os.path.join(tmpdir, f"ok{r}"))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path through the oracle restatement): one JSON line with the keys the
    driver reads; non-zero ranks of a multi-rank launch print nothing."""
    import json
    import subprocess
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-batch", "32"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env={**os.environ, "RANK": "0"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["unit"] == "samples/s" and rec["higher_is_better"] is True
    assert rec["metric"].startswith("loss-head fwd+bwd samples/s at global batch")
    assert rec["value"] > 0 and rec["steps"] == 1
    assert rec["cpu_baseline"]["kind"] == "port" and rec["cpu_baseline"]["cores"] >= 1 and rec["cpu_baseline"]["value"] == rec["value"]
    assert rec["e2e"] == {"value": rec["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in rec["config"]
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env={**os.environ, "RANK": "1"})
    assert other.returncode == 0 and other.stdout.strip() == ""


def _gather_worker(rank, world, port, fname, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cosmos_b200.loss import ClipLoss, gather_features
        rec = torch.load(os.path.join(GOLDEN, fname), weights_only=False)
        pay, ref = rec["payload"], rec["results"][rank]
        for ll in (False, True):
            for gwg in (False, True):
                want = ref[f"ll{int(ll)}_gwg{int(gwg)}"]
                img, txt = pay["image"][rank].clone().requires_grad_(True), pay["text"][rank].clone().requires_grad_(True)
                all_i, all_t = gather_features(img, txt, local_loss=ll, gather_with_grad=gwg, rank=rank, world_size=world)
                assert torch.equal(all_i.detach(), want["all_image"]) and torch.equal(all_t.detach(), want["all_text"]), (ll, gwg)
                probe = (all_i * pay["probe_image"]).sum() + (all_t * pay["probe_text"]).sum()
                assert probe.requires_grad == (want["g_image"] is not None), (ll, gwg)
                if probe.requires_grad:
                    probe.backward()
                    _close(img.grad, want["g_image"], rtol=1e-6, atol=1e-6)
                    _close(txt.grad, want["g_text"], rtol=1e-6, atol=1e-6)
                img, txt = pay["image"][rank].clone().requires_grad_(True), pay["text"][rank].clone().requires_grad_(True)
                lpi, lpt = ClipLoss(local_loss=ll, gather_with_grad=gwg, rank=rank, world_size=world).get_logits(img, txt, 7.5)
                _close(lpi.detach(), want["logits_per_image"], rtol=1e-6, atol=1e-6)
                _close(lpt.detach(), want["logits_per_text"], rtol=1e-6, atol=1e-6)
                (lpi * pay["probe_logits"][:lpi.shape[0], :lpi.shape[1]]).sum().backward()
                for got, key in ((img.grad, "g_logits_image"), (txt.grad, "g_logits_text")):
                    if want[key] is None:
                        assert got is None or float(got.abs().max()) == 0.0, (ll, gwg, key)
                    else:
                        _close(got, want[key], rtol=1e-5, atol=1e-5)
        # different shapes on the two sides: two collectives instead of one, same semantics
        a, b2 = torch.full((3, 4), float(rank)), torch.full((3, 6), float(rank) + 10.0)
        ga, gb = gather_features(a, b2, rank=rank, world_size=world)
        assert ga.shape == (3 * world, 4) and gb.shape == (3 * world, 6)
        assert [float(v) for v in ga[::3, 0]] == [float(r) for r in range(world)]
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world,fname,port", [(2, "gather_w2.pt", 29751), (3, "gather_w3.pt", 29752)])
def test_gather_features_and_get_logits_gloo(world, fname, port):
    """gather_features (three behaviours) and ClipLoss.get_logits (four modes) against what the unmodified reference
    returned on the same gloo ranks (src/open_clip/loss.py:21-65, 103-119): values and the gradients of a probe loss."""
    ctx = mp.get_context("spawn")
    with tempfile.TemporaryDirectory() as tmpdir:
        procs = [ctx.Process(target=_gather_worker, args=(r, world, port, fname, tmpdir)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        for r, p in enumerate(procs):
            assert p.exitcode == 0, f"rank {r} failed"
            assert os.path.exists(os.path.join(tmpdir, f"ok{r}"))


def test_get_logits_and_labels_single_process():
    from cosmos_b200.loss import ClipLoss
    g = torch.Generator().manual_seed(9)
    a, b = torch.randn(7, 12, generator=g), torch.randn(7, 12, generator=g)
    mod = ClipLoss(cache_labels=True)
    lpi, lpt = mod.get_logits(a, b, 3.0)
    assert torch.equal(lpi, 3.0 * a @ b.T) and torch.equal(lpt, 3.0 * b @ a.T)
    lab = mod.get_ground_truth(torch.device("cpu"), 7)
    assert lab.tolist() == list(range(7)) and mod.get_ground_truth(torch.device("cpu"), 7) is lab      # cached
    assert mod.get_ground_truth(torch.device("cpu"), 5).tolist() == list(range(5))
    sharded = ClipLoss(local_loss=True, rank=2, world_size=4)
    assert sharded.get_ground_truth(torch.device("cpu"), 3).tolist() == [6, 7, 8]
    assert ClipLoss(local_loss=False, rank=2, world_size=4).get_ground_truth(torch.device("cpu"), 3).tolist() == [0, 1, 2]


def test_no_grad_forward_forms_no_gradients(monkeypatch):
    """Under torch.no_grad() (a validation loss) the stored-exponential route must not run its backward inside forward():
    needs_input_grad stays True there, so the decision is taken from torch.is_grad_enabled()."""
    from tests import emulation
    emulation.install(monkeypatch)
    from cosmos_b200 import COSMOSLoss, infonce
    calls = {"bwd_e": 0, "keep_e": 0}
    monkeypatch.setattr(infonce, "_e_store_chunk", _force_e_route())
    real_bwd_e, real_fwd = infonce._k_bwd_e, infonce._k_fwd
    monkeypatch.setattr(infonce, "_k_bwd_e", lambda *a, **k: (calls.__setitem__("bwd_e", calls["bwd_e"] + 1), real_bwd_e(*a, **k))[1])
    monkeypatch.setattr(infonce, "_k_fwd", lambda *a, **k: (calls.__setitem__("keep_e", calls["keep_e"] + int(bool(k.get("keep_e") or (len(a) > 4 and a[4])))),
                                                            real_fwd(*a, **k))[1])
    case = torch.load(os.path.join(GOLDEN, "cosmos_w1_small.pt"), weights_only=False)[0]
    leaf = {k: [t.clone().requires_grad_(True) for t in v] for k, v in case["inputs"].items()}
    ls = torch.tensor(case["logit_scale"], requires_grad=True)
    args = (tuple(leaf["s_image"]), tuple(leaf["s_text"]), ls)
    kw = dict(t_image_features=leaf["t_image"], t_text_features=leaf["t_text"], output_dict=True,
              s_img_crossmodal_features=leaf["s_img_x"], s_txt_crossmodal_features=leaf["s_txt_x"])
    with torch.no_grad():
        out = COSMOSLoss()(*args, **kw)
    assert calls == {"bwd_e": 0, "keep_e": 0}
    assert not out["distill_loss"].requires_grad
    ref = case["out"]
    if case["distill_logit_scale"] is None:
        _close(out["distill_loss"], ref["distill_loss"], rtol=2e-5)
    _close(out["clip_loss"], ref["clip_loss"], rtol=2e-5)
    out = COSMOSLoss()(*args, **kw)                      # with gradients: the route runs
    assert calls["bwd_e"] > 0 and calls["keep_e"] > 0 and out["clip_loss"].requires_grad


REFERENCE = os.environ.get("COSMOS_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "src", "open_clip", "factory.py")),
                    reason="the reference checkout is only present in the build container")
def test_reference_create_loss_builds_the_drop_in():
    """The reference's own constructor site, create_loss(args) (src/open_clip/factory.py:372-415), compiled from the file
    where it lies with `open_clip.loss` swapped for `cosmos_b200.loss`: every branch the COSMOS recipes can take constructs
    the drop-in with the reference's keyword arguments; the out-of-scope losses raise loudly."""
    import ast
    import types
    import cosmos_b200.loss as ours
    path = os.path.join(REFERENCE, "src", "open_clip", "factory.py")
    tree = ast.parse(open(path).read(), filename=path)
    tree.body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "create_loss"]
    ns = {name: getattr(ours, name) for name in ("ClipLoss", "COSMOSLoss", "CoCaLoss", "DistillClipLoss", "SigLipLoss")}
    exec(compile(tree, path, "exec"), ns)
    base = dict(distill=False, model="ViT-B-16", siglip=False, cosmos=True, local_loss=False, gather_with_grad=False, rank=3,
                world_size=8, horovod=False, coca_caption_loss_weight=2.0, coca_contrastive_loss_weight=1.0)
    mod = ns["create_loss"](types.SimpleNamespace(**base))
    assert type(mod) is ours.COSMOSLoss and (mod.rank, mod.world_size, mod.cache_labels, mod.local_loss) == (3, 8, True, False)
    assert type(mod.clip_loss) is ours.ClipLoss and mod.clip_loss.world_size == 8
    mod = ns["create_loss"](types.SimpleNamespace(**{**base, "local_loss": True, "gather_with_grad": True}))
    assert mod.local_loss and mod.gather_with_grad
    assert type(ns["create_loss"](types.SimpleNamespace(**{**base, "cosmos": False}))) is ours.ClipLoss
    with pytest.raises(RuntimeError, match="Horovod"):
        ns["create_loss"](types.SimpleNamespace(**{**base, "cosmos": False, "horovod": True}))
    for other in ({"distill": True}, {"model": "coca_ViT-B-32"}, {"siglip": True}):
        with pytest.raises(NotImplementedError):
            ns["create_loss"](types.SimpleNamespace(**{**base, **other}))

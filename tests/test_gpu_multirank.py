"""GPU, world_size 2 over NCCL: the sharded loss head (rows local, columns all-gathered) against the
one-process oracle simulation of the reference's per-rank semantics, all four modes."""
import os
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir, keep_g=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    torch.set_num_threads(max(1, (os.cpu_count() or world) // world))      # the CPU oracle of every rank runs at the same time
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cosmos_b200 import COSMOSLoss, infonce
        from oracle import cosmos_oracle as O
        # (b, D, n_img, n_txt): dim 128 runs the CTA-pair backward, dim 512 the 4-CTA-cluster backward
        configs = ((160, 128, 3, 4), (192, 512, 4, 4))
        if keep_g == "e":   # stored-exponential route in the non-local modes (the local-loss modes keep the recompute kernels)
            infonce._e_store_chunk = lambda x_r, y_c, comm: 0 if comm.local_loss else min(3, x_r.shape[0])
            configs = ((192, 512, 4, 4),)
        elif keep_g:      # non-local modes at dim 512: image-side CLIP gradient = NCCL reduce-scatter of G^T x over the stored tiles
            infonce._G_STORE_MIN_BYTES = 0
            configs = ((192, 512, 4, 4),)
        cases = [(cfg, ll, gwg) for cfg in configs for ll, gwg in ((False, False), (False, True), (True, True), (True, False))]
        for (b, D, n_img, n_txt), ll, gwg in cases:
            shards_cpu = [O.make_features(b, D, seed=900 + r, n_img=n_img, n_txt=n_txt) for r in range(world)]
            # ours, on this rank's shard
            mine = {k: [t.bfloat16().cuda().requires_grad_(True) for t in v] for k, v in shards_cpu[rank].items()}
            ls = torch.tensor(14.2857, device="cuda", requires_grad=True)
            ds = torch.tensor(30.0, device="cuda", requires_grad=True)
            out = COSMOSLoss(local_loss=ll, gather_with_grad=gwg, cache_labels=True, rank=rank, world_size=world)(
                mine["s_image"], mine["s_text"], ls, t_image_features=mine["t_image"], t_text_features=mine["t_text"],
                output_dict=True, distill_logit_scale=ds, s_img_crossmodal_features=mine["s_img_x"],
                s_txt_crossmodal_features=mine["s_txt_x"])
            (out["distill_loss"] + out["clip_loss"]).backward()
            # oracle: every rank's scalar, gradients of their sum (= what autograd + all_gather-with-grad deliver)
            leafs = []
            for s in shards_cpu:
                d = {k: [t.bfloat16().float().requires_grad_(True) for t in v] for k, v in s.items()}
                d["logit_scale"] = torch.tensor(14.2857, requires_grad=True)
                d["distill_logit_scale"] = torch.tensor(30.0, requires_grad=True)
                leafs.append(d)
            outs = [O.cosmos_loss_rank(leafs, r, ll, gwg) for r in range(world)]
            sum(o["distill_loss"] + o["clip_loss"] for o in outs).backward()
            ref, rl = outs[rank], leafs[rank]
            for k in ("distill_loss", "clip_loss"):
                a, c = float(out[k]), float(ref[k])
                assert abs(a - c) <= 1e-4 * abs(c) + 3e-6, (ll, gwg, k, a, c)
            for k in ("s_image", "s_text", "s_img_x", "s_txt_x"):
                for t, r_ in zip(mine[k], rl[k]):
                    if r_.grad is None or float(r_.grad.abs().max()) == 0.0:
                        assert t.grad is None or float(t.grad.abs().max()) == 0.0
                        continue
                    ga, gb = t.grad.float().cpu().flatten().double(), r_.grad.flatten().double()
                    cos = float(ga @ gb / (ga.norm() * gb.norm()))
                    assert cos >= 0.9999, (ll, gwg, k, cos)
                    assert abs(float(ga.norm() / gb.norm()) - 1) < 5e-3, (ll, gwg, k)
            for name, t in (("logit_scale", ls), ("distill_logit_scale", ds)):
                c = float(rl[name].grad)
                assert abs(float(t.grad) - c) <= 3e-3 * abs(c) + 1e-6, (ll, gwg, name, float(t.grad), c)
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("keep_g,port", [(False, 29731), (True, 29732), ("e", 29733)])
def test_two_rank_nccl_all_modes(keep_g, port):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    with tempfile.TemporaryDirectory() as tmpdir:
        procs = [ctx.Process(target=_worker, args=(r, 2, port, tmpdir, keep_g)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(600)
        for r, p in enumerate(procs):
            assert p.exitcode == 0, f"rank {r} failed"
            assert os.path.exists(os.path.join(tmpdir, f"ok{r}"))

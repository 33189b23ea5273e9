#!/usr/bin/env python
"""Benchmark of the COSMOS loss head (BASELINE.json: loss-head fwd+bwd samples/s at global batch 32k).

    python bench.py --gpus N --steps K --warmup W            # ours, one rank per GPU (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU restatement of the reference

A step = COSMOSLoss.forward + backward to every student feature and both logit scales (+ the
collectives at N > 1) on one global batch of synthetic unit-norm bf16 embeddings: 8 + 8 student
image / text features, 8 + 8 cross-modal features, 2 + 2 teacher features per sample, dim 512.
The global batch is fixed (strong scaling): each rank holds global_batch / N rows.

Prints ONE JSON line (rank 0).  value = whole-job samples/s with inputs resident in HBM;
e2e = the same step driven through the public API from pinned HOST buffers (H2D copy of all
34 feature tensors and D2H read of the two loss values inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

DIM = 512
N_IMG, N_TXT = 8, 8
KEYS = (("s_image", N_IMG), ("s_text", N_TXT), ("s_img_x", N_IMG), ("s_txt_x", N_TXT), ("t_image", 2), ("t_text", 2))
LOGIT_SCALE = 14.2857
PROFILE_TAG = "r02m"    # profiles/ncu_<kernel>_<tag>.txt: the committed ncu summaries this round's roofline.traffic comes from


def algorithmic_flops(n_global: int, dim: int = DIM) -> float:
    """352 * N^2 * D: 64 distill pairs x (fwd 2 + bwd 2) + 16 clip pairs x (2 + 2 + 2)."""
    return 352.0 * n_global * n_global * dim


def ncu_traffic(kind, tag="r02"):
    """DRAM bytes of one launch of the dominant kernel, from the committed ncu --set full summary."""
    path = os.path.join(ROOT, "profiles", "ncu_%s_%s.txt" % (kind, tag))
    try:
        tot = 0.0
        for line in open(path):
            f = line.split()
            if f and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[2]]
        return tot or None
    except OSError:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return p.get("bf16_tflops", 1590.0), p.get("bf16_tflops_sustained", 1400.0), p.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------

def host_inputs(b_local: int, rank: int, pinned: bool):
    """One [n * b, D] bf16 host buffer per feature list (the layout train.py:171-182 chunks)."""
    g = torch.Generator().manual_seed(1234 + rank)
    z = torch.randn(b_local, DIM, generator=g)
    out = {}
    for key, n in KEYS:
        views = []
        for _ in range(n):
            x = z + 2.0 * torch.randn(b_local, DIM, generator=g)
            views.append(torch.nn.functional.normalize(x, dim=-1).to(torch.bfloat16))
        buf = torch.cat(views, dim=0)
        out[key] = buf.pin_memory() if pinned else buf
    return out


class Clocks:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc, self.path = None, f"/tmp/cosmos_clocks_{os.getpid()}.csv"
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(gpu_index)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons, power = [], [], set(), []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(s for s, p in zip(sm, power) if p >= 0.5 * max(power)) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}


# ------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------

class Instrument:
    """Counts our kernel launches and times the InfoNCE launches with CUDA events on the launching stream."""
    KINDS = ("fwd", "bwd", "bwd_e", "colgrad")
    NAMES = {"fwd": "infonce_fwd_kernel<pair> (+ column merge)", "bwd": "infonce_bwd_quad_kernel (+ dscale reduce)",
             "bwd_e": "infonce_bwd_e2_kernel (+ dscale reduce)", "colgrad": "infonce_bwd_e2t_kernel (column-side gradient)"}

    def __init__(self):
        from cosmos_b200 import infonce
        self.mod = infonce
        self.launches = 0
        self.events = {k: [] for k in self.KINDS}
        self.flops = {k: [] for k in self.KINDS}
        self.enabled = False
        self._fwd, self._loss, self._bwd, self._colgrad = infonce._k_fwd, infonce._k_loss_sums, infonce._k_bwd, infonce._k_colgrad
        self._bwd_e = infonce._k_bwd_e
        self._bwd_e_cols = infonce._k_bwd_e_cols
        infonce._k_bwd_e_cols = self.bwd_e_cols
        self._scale16, self._lse2_merge = infonce._k_scale16, infonce._k_lse2_merge
        infonce._k_scale16, infonce._k_lse2_merge = self.scale16, self.lse2_merge
        infonce._k_fwd, infonce._k_loss_sums, infonce._k_bwd, infonce._k_colgrad = self.fwd, self.loss, self.bwd, self.colgrad
        infonce._k_bwd_e = self.bwd_e

    def _timed(self, kind, flops, fn, *a, **kw):
        if not self.enabled:
            return fn(*a, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **kw)
        e1.record()
        self.events[kind].append((e0, e1))
        self.flops[kind].append(flops)
        return r

    def fwd(self, x, y, *a, **kw):
        self.launches += 2   # tile kernel + column-statistics merge
        fl = 2.0 * x.shape[0] * y.shape[0] * x.shape[1] * y.shape[1] * x.shape[2]
        return self._timed("fwd", fl, self._fwd, x, y, *a, **kw)

    def loss(self, *a):
        self.launches += 1
        return self._loss(*a)

    def bwd(self, x, y, *a):
        want_dx, want_ds = a[10], a[11]          # (..., weight, upstream, want_dx, want_dscale[, g_out])
        self.launches += 1 + (1 if want_ds else 0)
        fl = 2.0 * x.shape[0] * y.shape[0] * x.shape[1] * y.shape[1] * x.shape[2] if want_dx else 0.0
        return self._timed("bwd", fl, self._bwd, x, y, *a)

    def bwd_e(self, x, y, *a, **kw):
        want_ds = a[13]          # (label_offset, scale, e, off, diag, row, col, a_row, a_col, s_row, s_col, weight, upstream, want_dscale[, g_out])
        self.launches += 1 + (1 if want_ds else 0)
        fl = 2.0 * x.shape[0] * y.shape[0] * x.shape[1] * y.shape[1] * x.shape[2]
        return self._timed("bwd_e", fl, self._bwd_e, x, y, *a, **kw)

    def bwd_e_cols(self, x, y, *a, **kw):
        self.launches += 1                       # column-side gradient G^T x from the stored exponentials
        fl = 2.0 * x.shape[0] * y.shape[0] * x.shape[1] * y.shape[1] * x.shape[2]
        return self._timed("colgrad", fl, self._bwd_e_cols, x, y, *a, **kw)

    def scale16(self, *a, **kw):
        self.launches += 1                       # unit gradients x upstream / stand-in (backward)
        return self._scale16(*a, **kw)

    def lse2_merge(self, *a, **kw):
        self.launches += 1                       # merge of the all-gathered column statistics (sharded runs)
        return self._lse2_merge(*a, **kw)

    def colgrad(self, g, x2d, n_c, n_cols):
        self.launches += 1                       # the column-side gradient GEMM (G^T x) on the stored tiles
        return self._timed("colgrad", 2.0 * x2d.shape[0] * n_c * n_cols * x2d.shape[1], self._colgrad, g, x2d, n_c, n_cols)

    def reset(self):
        self.launches = 0
        self.events = {k: [] for k in self.KINDS}
        self.flops = {k: [] for k in self.KINDS}

    def summary(self):
        out = {}
        for kind in self.KINDS:
            ms = [a.elapsed_time(b) for a, b in self.events[kind]]
            if ms:
                out[kind] = {"launches": len(ms), "ms_total": sum(ms), "ms_avg": sum(ms) / len(ms),
                             "flops_avg": sum(self.flops[kind]) / len(ms)}
        return out


class LossHead:
    """The timed workload at one global batch: synthetic inputs (pinned host + device copies), the drop-in COSMOSLoss and the
    step / timing helpers shared by the headline line, the config-2 entry and the same-size CPU/GPU pair."""

    def __init__(self, n_global, dev, rank, world, inst, flush):
        from cosmos_b200 import COSMOSLoss
        assert n_global % world == 0
        self.n_global, self.dev, self.rank, self.world, self.inst, self.flush = n_global, dev, rank, world, inst, flush
        self.b = n_global // world
        self.host = host_inputs(self.b, rank, pinned=True)
        # two sets of device inputs: in the host-buffer (e2e) leg the copy of step k + 1's inputs runs on a side stream while
        # step k computes, as a data loader's prefetch would; every step's copy is inside the timed region, the first exposed
        self.dev_bufs = [{k: torch.empty_like(v, device=dev).requires_grad_(k not in ("t_image", "t_text")) for k, v in self.host.items()}
                         for _ in range(2)]
        self.dev_buf = self.dev_bufs[0]
        with torch.no_grad():
            for k in self.host:
                self.dev_buf[k].copy_(self.host[k])
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]      # copy into buffer i has landed
        self.free = [torch.cuda.Event(), torch.cuda.Event()]       # the step that read buffer i is done with it
        self.in_flight = [False, False]
        self.cur = 0
        self.logit_scale = torch.tensor(LOGIT_SCALE, device=dev, requires_grad=True)
        self.distill_scale = torch.tensor(LOGIT_SCALE, device=dev, requires_grad=True)
        self.loss_mod = COSMOSLoss(local_loss=False, gather_with_grad=False, cache_labels=True, rank=rank, world_size=world)
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host.values())

    def _issue_copy(self, i: int):
        with torch.no_grad(), torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[i])
            for k in self.host:
                self.dev_bufs[i][k].copy_(self.host[k], non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self.in_flight[i] = True

    def step(self, from_host: bool, prefetch_next: bool = False):
        host, dev_buf = self.host, self.dev_bufs[0]
        if from_host:
            cur = self.cur
            dev_buf = self.dev_bufs[cur]
            if not self.in_flight[cur]:
                self._issue_copy(cur)
            torch.cuda.current_stream().wait_event(self.ready[cur])
            if prefetch_next:
                self._issue_copy(1 - cur)
        for t in list(dev_buf.values()) + [self.logit_scale, self.distill_scale]:
            t.grad = None
        n = dict(KEYS)
        out = self.loss_mod(dev_buf["s_image"].chunk(n["s_image"]), dev_buf["s_text"].chunk(n["s_text"]), self.logit_scale,
                            t_image_features=dev_buf["t_image"].chunk(2), t_text_features=dev_buf["t_text"].chunk(2),
                            output_dict=True, distill_logit_scale=self.distill_scale,
                            s_img_crossmodal_features=dev_buf["s_img_x"].chunk(n["s_img_x"]),
                            s_txt_crossmodal_features=dev_buf["s_txt_x"].chunk(n["s_txt_x"]))
        total = out["distill_loss"] + out["clip_loss"]
        total.backward()
        if from_host:
            self.free[self.cur].record()
            self.in_flight[self.cur] = False
            self.cur ^= 1
            return torch.stack([out["distill_loss"].detach(), out["clip_loss"].detach()]).cpu()   # 8-byte D2H
        return out

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    L2_BYTES = 126 * 1024 * 1024

    def l2_policy(self) -> str:
        if self.h2d_bytes > self.L2_BYTES:
            return ("inputs larger than L2 (%.0f MB per rank, plus %.1f GB of stored exponentials written and read per step): "
                    "K steps timed back to back inside one barrier + synchronize bracket" %
                    (self.h2d_bytes / 1e6, 80.0 * 2.0 * self.b * self.n_global / 1e9))
        return "256 MB flush write before every timed step (inputs %.0f MB per rank fit L2), each step bracketed on its own" % (
            self.h2d_bytes / 1e6)

    def timed(self, from_host: bool, steps: int):
        """Total ms of `steps` steps, CUDA events on the launching stream, max over ranks.  The K steps run back to back inside
        ONE barrier + synchronize bracket when a rank's inputs alone exceed L2 (the contract's "inputs larger than L2" option:
        nothing of step k is still cached when step k + 1 reads it - each step also streams GBs of exponentials through L2);
        small configurations instead flush L2 with a 256 MB write before every step and bracket each step on its own."""
        total_ms = 0.0
        if self.h2d_bytes > self.L2_BYTES:
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(steps):
                self.step(from_host, prefetch_next=from_host and k + 1 < steps)
            e1.record()
            self.barrier()
            total_ms = e0.elapsed_time(e1)
        else:
            for _ in range(steps):
                self.flush.fill_(1)                  # evict L2 between timed iterations (256 MB > 126 MB L2)
                self.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                self.step(from_host)
                e1.record()
                torch.cuda.synchronize()
                total_ms += e0.elapsed_time(e1)
        t = torch.tensor([total_ms], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def graph_ms_per_step(self, steps: int):
        """The same device-resident step captured ONCE with torch.cuda.graph and replayed: what the kernels take when the host
        that launches them is out of the picture (small configurations are ~130 launches of 3 - 700 us, and a busy or slow
        host core shows up directly in their eager-launch time).  Same L2 policy as `timed`; one process only."""
        if self.world > 1:
            return None
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.step(False)
        torch.cuda.current_stream().wait_stream(side)
        for t in list(self.dev_bufs[0].values()) + [self.logit_scale, self.distill_scale]:
            t.grad = None
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.step(False)
        graph.replay()
        torch.cuda.synchronize()
        total_ms = 0.0
        if self.h2d_bytes > self.L2_BYTES:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            total_ms = e0.elapsed_time(e1)
        else:
            for _ in range(steps):
                self.flush.fill_(1)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                graph.replay()
                e1.record()
                torch.cuda.synchronize()
                total_ms += e0.elapsed_time(e1)
        loss = [float(out["distill_loss"].detach()), float(out["clip_loss"].detach())]
        del graph
        return {"ms_per_step": total_ms / steps, "loss": loss}

    def measure(self, steps, warmup, e2e=True):
        """-> dict(ms_per_step, e2e_ms_per_step, launches, kernels, loss): warm-up, device-resident leg, host-buffer leg."""
        inst = self.inst
        for _ in range(max(warmup, 3)):
            self.step(False)
        self.barrier()
        inst.reset()
        inst.enabled = True
        dev_ms = self.timed(False, steps)
        launches = inst.launches
        inst.enabled = False
        kernels = inst.summary()
        if e2e:
            for _ in range(2):
                self.step(True)
            e2e_ms = self.timed(True, steps)
            last = self.step(True)
            last = [float(x) for x in last]
        else:
            out = self.step(False)
            e2e_ms, last = float("nan"), [float(out["distill_loss"].detach()), float(out["clip_loss"].detach())]
        return {"ms_per_step": dev_ms / steps, "e2e_ms_per_step": e2e_ms / steps, "launches": launches, "kernels": kernels, "loss": last}


def parity_check(dev, rank, world):
    """Before the timed loop, at THIS world size and through the same route as the headline step (stored exponentials, forced
    at this small size): a global batch of 512 (rows sharded over the ranks) in the default mode and with gather_with_grad,
    rank 0 against the oracle's per-rank restatement of the reference (oracle.cosmos_loss_rank: src/open_clip/loss.py:21-65,
    103-142, 176-207).  The driver's scaling run so carries gradient parity under NCCL at every N, not only throughput."""
    from cosmos_b200 import COSMOSLoss, infonce
    from oracle import cosmos_oracle as O
    b = max(64, 512 // world)
    saved = infonce._E_STORE_MIN_BYTES
    infonce._E_STORE_MIN_BYTES = 0
    modes = {}
    try:
        for name, gwg in (("default", False), ("gather_with_grad", True)):
            mine = O.make_features(b, DIM, seed=900 + rank)
            x = {k: [t.bfloat16().to(dev).requires_grad_(k not in ("t_image", "t_text")) for t in v] for k, v in mine.items()}
            ls = torch.tensor(LOGIT_SCALE, device=dev, requires_grad=True)
            ds = torch.tensor(30.0, device=dev, requires_grad=True)
            out = COSMOSLoss(local_loss=False, gather_with_grad=gwg, cache_labels=True, rank=rank, world_size=world)(
                x["s_image"], x["s_text"], ls, t_image_features=x["t_image"], t_text_features=x["t_text"], output_dict=True,
                distill_logit_scale=ds, s_img_crossmodal_features=x["s_img_x"], s_txt_crossmodal_features=x["s_txt_x"])
            (out["distill_loss"] + out["clip_loss"]).backward()
            torch.cuda.synchronize()
            if rank != 0:
                continue
            torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
            leafs = []
            for r in range(world):
                d = {k: [t.bfloat16().float().requires_grad_(True) for t in v] for k, v in O.make_features(b, DIM, seed=900 + r).items()}
                d["logit_scale"] = torch.tensor(LOGIT_SCALE, requires_grad=True)
                d["distill_logit_scale"] = torch.tensor(30.0, requires_grad=True)
                leafs.append(d)
            # what rank 0 sees: its own scalar; with gather_with_grad the all_gather's backward sums every rank's scalar
            outs = [O.cosmos_loss_rank(leafs, r, False, gwg) for r in (range(world) if gwg else (0,))]
            sum(o["distill_loss"] + o["clip_loss"] for o in outs).backward()
            ref, rl = outs[0], leafs[0]
            rec = {"loss_rel": 0.0, "grad_cos_min": 1.0, "grad_norm_rel": 0.0, "dscale_rel": 0.0}
            for k in ("distill_loss", "clip_loss"):
                rec["loss_rel"] = max(rec["loss_rel"], abs(float(out[k]) - float(ref[k])) / abs(float(ref[k])))
            for k in ("s_image", "s_text", "s_img_x", "s_txt_x"):
                for t, r_ in zip(x[k], rl[k]):
                    if r_.grad is None or float(r_.grad.abs().max()) == 0.0:
                        if t.grad is not None and float(t.grad.abs().max()) != 0.0:
                            rec["grad_cos_min"] = 0.0            # a gradient where the reference has none
                        continue
                    ga, gb = t.grad.float().cpu().flatten().double(), r_.grad.flatten().double()
                    rec["grad_cos_min"] = min(rec["grad_cos_min"], float(ga @ gb / (ga.norm() * gb.norm())))
                    rec["grad_norm_rel"] = max(rec["grad_norm_rel"], abs(float(ga.norm() / gb.norm()) - 1.0))
            for name_s, t in (("logit_scale", ls), ("distill_logit_scale", ds)):
                c = float(rl[name_s].grad)
                rec["dscale_rel"] = max(rec["dscale_rel"], abs(float(t.grad) - c) / abs(c))
            rec["ok"] = bool(rec["loss_rel"] <= 1e-4 and rec["grad_cos_min"] >= 0.9999 and rec["grad_norm_rel"] <= 5e-3
                             and rec["dscale_rel"] <= 3e-3)
            modes[name] = rec
    finally:
        infonce._E_STORE_MIN_BYTES = saved
    if rank != 0:
        return None
    return {"ok": all(m["ok"] for m in modes.values()), "world_size": world, "batch_per_rank": b, "global_batch": b * world,
            "route": "stored exponentials (the headline route), bf16 features, dim 512, 80 pairs",
            "checker": "oracle.cosmos_loss_rank on host cores (fp32 on the same bf16-valued inputs)",
            "tolerance": {"loss_rel": 1e-4, "grad_cos": 0.9999, "grad_norm_rel": 5e-3, "dscale_rel": 3e-3},
            "max_loss_rel": max(m["loss_rel"] for m in modes.values()),
            "min_grad_cos": min(m["grad_cos_min"] for m in modes.values()), "modes": modes}


def run_ours(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    n_global = args.global_batch
    parity = None if args.no_parity_check else parity_check(dev, rank, world)
    inst = Instrument()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    head = LossHead(n_global, dev, rank, world, inst, flush)
    b, h2d_bytes = head.b, head.h2d_bytes

    clocks = Clocks(local_rank) if rank == 0 else None     # sampled from the warm-up on: short timed regions still get samples
    # Warm-up: the W steps asked for (at least 3), and then on until a second of steps has run - a rank of an 8-GPU job steps
    # in 23 ms, and three of those after the seconds-long host-side parity check are over before the clocks have ramped up.
    # Every rank runs the same count (rank 0 decides), the count is what the line reports as "warmup".
    import time as _time
    n_warm = 0
    t_w = _time.perf_counter()
    while True:
        head.step(False)
        n_warm += 1
        if n_warm < max(args.warmup, 3):
            continue
        torch.cuda.synchronize()
        go_on = torch.tensor([1 if (_time.perf_counter() - t_w < 1.0 and n_warm < 100) else 0], device=dev)
        if world > 1:
            dist.broadcast(go_on, src=0)
        if not int(go_on.item()):
            break
    head.barrier()
    inst.reset()
    inst.enabled = True
    dev_ms = head.timed(False, args.steps)
    launches = inst.launches
    inst.enabled = False
    ksum = inst.summary()
    clk = clocks.stop() if clocks else None
    if args.no_e2e:
        out = head.step(False)
        e2e_ms, last = float("nan"), [float(out["distill_loss"].detach()), float(out["clip_loss"].detach())]
    else:
        for _ in range(2):
            head.step(True)
        e2e_ms = head.timed(True, args.steps)
        last = head.step(True)

    if rank == 0:
        burst, sustained, hbm, src = peaks()
        ms_per_step = dev_ms / args.steps
        value = n_global / (ms_per_step * 1e-3)
        e2e_value = n_global / (e2e_ms / args.steps * 1e-3)
        dom = max(ksum, key=lambda kind: ksum[kind]["ms_total"])
        k = ksum[dom]
        achieved = k["flops_avg"] / (k["ms_avg"] * 1e-3) / 1e12
        step_tflops = algorithmic_flops(n_global) / world / (ms_per_step * 1e-3) / 1e12
        kernel_ms = sum(v["ms_total"] for v in ksum.values()) / args.steps
        line = {
            "metric": "loss-head fwd+bwd samples/s at global batch %d" % n_global,
            "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "COSMOS ViT-B/16 loss head fwd+bwd, global batch %d, dim 512, 8+8 student / 8+8 "
                                   "cross-modal / 2+2 teacher features per sample, 80 InfoNCE pairs" % n_global,
                       "global_batch": n_global, "per_gpu_batch": b, "parallelism": "dp%d (rows sharded, columns all-gathered)" % world,
                       "l2": head.l2_policy(),
                       "loss": [float(x) for x in last]},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8,
                    "ms_per_step": e2e_ms / args.steps,
                    "input_pipeline": "pinned host -> one of two device input buffers on a copy stream; step k + 1's copy overlaps "
                                      "step k's kernels (the first copy of the timed region is exposed); the loss is read back "
                                      "to the host after every step"},
            "gpu_launches": launches,
            "clocks": clk,
            "parity_check": parity,
            "roofline": {"bound": "tensor",
                         "kernel": Instrument.NAMES[dom],
                         "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained,
                         "traffic": ncu_traffic(dom, PROFILE_TAG) if (n_global == 32768 and world == 1) else None,
                         "traffic_note": "dram__bytes_read+write of one launch of this kernel from the committed ncu --set full "
                                         "summary (profiles/ncu_<kernel>_%s.txt: global batch 32768, one GPU, 16 pairs per "
                                         "launch); null for other configurations" % PROFILE_TAG,
                         "executed_tflops": achieved * (2.0 if dom == "bwd" else 1.0),   # the recompute backward runs S and dX
                         "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step); burst %.1f" % (src, burst),
                         "launch_ms_avg": k["ms_avg"], "launches_timed": k["launches"],
                         "step_algorithmic_tflops_per_gpu": step_tflops, "step_frac": step_tflops / sustained,
                         "step_minus_kernels_ms": ms_per_step - kernel_ms,
                         "kernels": ksum},
        }
    del head
    torch.cuda.empty_cache()
    if rank == 0:
        def extra(fn, *a, **kw):
            """The other rows of the line: a failure in one of them is reported in its place, the headline still prints."""
            try:
                return fn(*a, **kw)
            except Exception as e:      # noqa: BLE001
                sys.stderr.write("bench.py: %s failed: %s: %s\n" % (fn.__name__, type(e).__name__, e))
                return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

        if world == 1 and not args.no_extras:
            line["config2_global_batch_4096"] = extra(bench_other_batch, 4096, dev, inst, flush, steps=10)
            line["same_size_cpu_vs_gpu"] = extra(bench_same_size, dev, inst, flush, args.cpu_seconds)
            line["ema"] = extra(bench_ema, dev, flush)
            line["xattn"] = extra(bench_xattn, dev, flush)
            line["retrieval"] = extra(bench_retrieval, dev, flush)
            line["eager_gpu_baseline"] = extra(bench_eager_gpu, dev)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = extra(cpu_baseline, args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _graph_leg(head, steps):
    """LossHead.graph_ms_per_step, never fatal for the bench line: -> dict with ms_per_step (None + error text on failure)."""
    try:
        return head.graph_ms_per_step(steps)
    except Exception as e:      # noqa: BLE001 - an extra leg must not take the headline down with it
        torch.cuda.synchronize()
        return {"ms_per_step": None, "loss": None, "error": "%s: %s" % (type(e).__name__, str(e)[:200])}


def bench_other_batch(n_global, dev, inst, flush, steps=10):
    """BASELINE config 2 (global batch 4096, dim 512, bf16, one B200) through the same LossHead as the headline line."""
    head = LossHead(n_global, dev, 0, 1, inst, flush)
    m = head.measure(steps, 3)
    g = _graph_leg(head, steps)
    burst, sustained, _, src = peaks()
    tf = algorithmic_flops(n_global) / (m["ms_per_step"] * 1e-3) / 1e12
    return {"workload": "COSMOS ViT-B/16 loss head fwd+bwd, global batch %d, dim 512, bf16, 1 B200" % n_global,
            "ms_per_step": m["ms_per_step"], "value": n_global / (m["ms_per_step"] * 1e-3), "unit": "samples/s",
            "cuda_graph": {"error": g["error"]} if "error" in g else {
                "ms_per_step": g["ms_per_step"], "value": n_global / (g["ms_per_step"] * 1e-3), "unit": "samples/s",
                "frac_of_sustained_peak": algorithmic_flops(n_global) / (g["ms_per_step"] * 1e-3) / 1e12 / sustained,
                "loss": g["loss"],
                "note": "the same public-API step (COSMOSLoss forward + backward) captured once with torch.cuda.graph "
                        "and replayed: kernel time without the host's launch path"},
            "e2e": {"value": n_global / (m["e2e_ms_per_step"] * 1e-3), "unit": "samples/s", "ms_per_step": m["e2e_ms_per_step"],
                    "h2d_bytes_per_step": head.h2d_bytes, "d2h_bytes_per_step": 8},
            "algorithmic_tflops": tf, "frac_of_sustained_peak": tf / sustained, "frac_of_burst_peak": tf / burst, "peak_source": src,
            "gpu_launches": m["launches"], "kernels": m["kernels"], "loss": m["loss"], "steps": steps}


def bench_same_size(dev, inst, flush, seconds, n_global=1024):
    """One ratio on the SAME configuration for both arms: global batch 1024 (what the reference arm's bounded sample is) on the
    B200 through the public API (device-resident and from pinned host buffers) and on the host cores through the oracle port."""
    head = LossHead(n_global, dev, 0, 1, inst, flush)
    m = head.measure(20, 3)
    g = _graph_leg(head, 20)
    del head
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_step_fn(n_global)
    step()
    t0 = time.perf_counter()
    n = 0
    while n < 2 or (time.perf_counter() - t0 < min(seconds, 8.0) and n < 50):
        step()
        n += 1
    cpu_ms = (time.perf_counter() - t0) / n * 1e3
    gpu, e2e, cpu = n_global / (m["ms_per_step"] * 1e-3), n_global / (m["e2e_ms_per_step"] * 1e-3), n_global / (cpu_ms * 1e-3)
    return {"workload": "COSMOS loss head fwd+bwd, global batch %d, dim 512, 80 pairs: the same configuration on both arms "
                        "(GPU: bf16 kernels; CPU: fp32 oracle port, %d torch threads)" % (n_global, cores),
            "gpu_ms_per_step": m["ms_per_step"], "gpu_samples_per_s": gpu, "gpu_e2e_ms_per_step": m["e2e_ms_per_step"],
            "gpu_cuda_graph_ms_per_step": g["ms_per_step"],
            "gpu_e2e_samples_per_s": e2e, "cpu_port_ms_per_step": cpu_ms, "cpu_port_samples_per_s": cpu, "cpu_cores": cores,
            "cpu_steps": n, "same_config": True, "ratio": gpu / cpu, "e2e_ratio": e2e / cpu}


# ------------------------------------------------------------------------------------------------
# the other rows of the hot path, timed as separate entries of the same JSON line (N = 1 only)
# ------------------------------------------------------------------------------------------------

def vitb16_cosmos_param_shapes():
    """Parameter table of ViT-B/16 COSMOS (student == teacher): 323 tensors, 152,382,466 fp32 parameters
    (SURVEY.md §8(a) A5), rebuilt from the architecture so that no checkpoint is needed."""
    def block(w, mlp):
        return [(w,), (w,), (3 * w, w), (3 * w,), (w, w), (w,), (w,), (w,), (mlp, w), (mlp,), (w, mlp), (w,)]
    shapes = [(768, 3, 16, 16), (768,), (197, 768), (768,), (768,)]
    for _ in range(12):
        shapes += block(768, 3072)
    shapes += [(768,), (768,), (768, 512), (49408, 512), (77, 512)]
    for _ in range(12):
        shapes += block(512, 2048)
    shapes += [(512,), (512,), (512, 512), (), ()]
    pool = [(1536, 512), (1536,), (512, 512), (512,), (512,), (512,), (512,), (512,)]
    shapes += pool + pool + [(512, 768), (512,), (512, 512), (512,)]
    return shapes


def _event_ms(fn, reps, flush=None):
    ms = []
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2], ms[0]


def _graph_event_ms(fn, reps, flush=None):
    """`fn` (launches only, no host reads) captured once with torch.cuda.graph, then the replay event-timed like _event_ms:
    -> (median, best) ms, or (None, None) when the capture fails (an extra number must not take the line down)."""
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        graph.replay()
        torch.cuda.synchronize()
        return _event_ms(graph.replay, reps, flush)
    except Exception:      # noqa: BLE001
        torch.cuda.synchronize()
        return None, None


def bench_ema(dev, flush):
    from cosmos_b200 import EmaPlan, ema_update_
    from oracle import cosmos_oracle as O
    shapes = vitb16_cosmos_param_shapes()
    g = torch.Generator(device=dev).manual_seed(5)
    student = [torch.randn(s, generator=g, device=dev) * 0.02 for s in shapes]
    teacher = [torch.randn(s, generator=g, device=dev) * 0.02 for s in shapes]
    n_params = sum(t.numel() for t in teacher)
    plan = EmaPlan(student, teacher)
    for _ in range(3):
        plan.apply(0.99)
        ema_update_(student, teacher, 0.99)
    med, best = _event_ms(lambda: plan.apply(0.99), 20, flush)
    api_med, _ = _event_ms(lambda: ema_update_(student, teacher, 0.99), 20, flush)
    ref_med, _ = _event_ms(lambda: O.ema_update_(teacher, student, 0.99), 5, flush)      # the reference loop, eager, same GPU
    _, _, hbm, src = peaks()
    gbs = 12.0 * n_params / (med * 1e-3) / 1e9
    return {"workload": "EMA teacher update, ViT-B/16 COSMOS: %d fp32 parameters in %d tensors, one launch" % (n_params, len(shapes)),
            "ms": med, "ms_best": best, "algorithmic_bytes": 12 * n_params,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "peak_source": src},
            "ema_update_api_ms": api_med, "eager_reference_loop_ms": ref_med}


def bench_xattn(dev, flush):
    """BASELINE config 4 at the reference's shapes: 8 queries per sample, 77 text / 196 image tokens, d = 512, batch 1024."""
    from cosmos_b200.pooler import AttentionalCrossPooler, crossmodal_features
    from oracle import cosmos_oracle as O
    out = {}
    B, n, d, heads = 1024, 8, 512, 8
    for name, L in (("text_tokens_77", 77), ("image_tokens_196", 196)):
        params, _, _, _ = O.make_pooler_case(d, 4, 1, 1, seed=3)
        mod = AttentionalCrossPooler(d, d, heads).to(dev)
        mod.load_state_dict(params)
        g = torch.Generator(device=dev).manual_seed(L)
        tokens = torch.randn(B, L, d, generator=g, device=dev).bfloat16().requires_grad_(True)
        feats = torch.randn(n * B, d, generator=g, device=dev).bfloat16().requires_grad_(True)
        w = torch.randn(n * B, d, generator=g, device=dev).bfloat16()

        def step():
            for t in [tokens, feats] + list(mod.parameters()):
                t.grad = None
            xm = crossmodal_features(mod, tokens, feats, B)
            xm.backward(w)

        for _ in range(3):
            step()
        med, best = _event_ms(step, 10, flush)
        g_med, g_best = _graph_event_ms(step, 10, flush)       # the same step replayed from a CUDA graph (no host launch path)
        # algorithmic flops: K/V projection once per unique sample + 8 queries, x3 for fwd+bwd
        fwd = 2.0 * B * L * d * 2 * d + 2.0 * n * B * d * d * 2 + 4.0 * n * B * L * d
        p16 = {k: v.to(dev) for k, v in params.items()}

        def eager():
            t32, f32 = tokens.detach().requires_grad_(True), feats.detach().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                xm = O.cosmos_crossmodal(f32, t32, {k: v.requires_grad_(True) for k, v in p16.items()}, heads, B)
            xm.backward(w.to(xm.dtype))

        eager(); eager()
        ref_med, _ = _event_ms(eager, 3, flush)
        burst, sustained, hbm, src = peaks()
        tf = 3.0 * fwd / (med * 1e-3) / 1e12
        # With ONE query per (sample, crop) the honest roofline of this shape is HBM bytes, not flops: every token is read in
        # the forward, read again and its gradient written in the backward; features, their gradient and the output are 8 rows
        # per sample.  The K/V projection the kernels still materialise is why they sit far below it.
        alg_bytes = 2.0 * (3 * B * L * d + 5 * n * B * d)
        gbs = alg_bytes / (med * 1e-3) / 1e9
        out[name] = {"ms": med, "ms_best": best, "cuda_graph_ms": g_med, "cuda_graph_ms_best": g_best,
                     "algorithmic_tflops": tf, "frac_of_bf16_peak": tf / burst,
                     "hbm_roofline": {"bound": "hbm", "algorithmic_bytes": alg_bytes, "achieved": gbs, "peak": hbm, "unit": "GB/s",
                                      "frac": gbs / hbm, "peak_source": src},
                     "eager_restatement_ms": ref_med, "batch": B, "queries_per_sample": n, "tokens": L, "dim": d}
    # BASELINE config 4 as literally written (token sequences as queries: 77 x 197 and 197 x 77, width 768, 12 heads),
    # through the module's forward(x, q); SURVEY §0 D2 explains why the reference never runs this shape.
    for name, (Lq, Lk) in (("literal_77q_x_197kv_w768", (77, 197)), ("literal_197q_x_77kv_w768", (197, 77))):
        d2, h2, B2 = 768, 12, 1024      # BASELINE config 4: batch 1024
        params, _, _, _ = O.make_pooler_case(d2, 4, 1, 1, seed=4)
        mod = AttentionalCrossPooler(d2, d2, h2).to(dev)
        mod.load_state_dict(params)
        g = torch.Generator(device=dev).manual_seed(Lq)
        x = torch.randn(B2, Lk, d2, generator=g, device=dev).bfloat16().requires_grad_(True)
        q = torch.randn(B2, Lq, d2, generator=g, device=dev).bfloat16().requires_grad_(True)
        w = torch.randn(B2, Lq, d2, generator=g, device=dev).bfloat16()

        def step2():
            for t in [x, q] + list(mod.parameters()):
                t.grad = None
            mod(x, q).backward(w)

        for _ in range(3):
            step2()
        med, best = _event_ms(step2, 5, flush)
        g_med, g_best = _graph_event_ms(step2, 5, flush)
        fwd = 2.0 * B2 * (Lq + 2 * Lk) * d2 * d2 + 4.0 * B2 * Lq * Lk * d2 + 2.0 * B2 * Lq * d2 * d2
        tf2 = 3.0 * fwd / (med * 1e-3) / 1e12
        out[name] = {"ms": med, "ms_best": best, "cuda_graph_ms": g_med, "cuda_graph_ms_best": g_best, "algorithmic_tflops": tf2,
                     "tensor_roofline": {"bound": "tensor", "achieved": tf2, "peak": peaks()[0], "unit": "TFLOP/s", "frac": tf2 / peaks()[0]},
                     "batch": B2, "queries": Lq, "tokens": Lk, "dim": d2, "heads": h2}
    return out


def bench_retrieval(dev, flush, n=32768, n_cpu=2048):
    """Eval metrics: paired retrieval ranks at the headline batch, fp32 on the CUDA cores, next to the
    reference's CPU path (similarity matrix + argsort + position search, the oracle restatement) on a bounded sample."""
    from cosmos_b200.retrieval import get_clip_metrics, retrieval_ranks
    from oracle import cosmos_oracle as O
    g = torch.Generator(device=dev).manual_seed(21)
    z = torch.randn(n, DIM, generator=g, device=dev)
    img = torch.nn.functional.normalize(z + 2.0 * torch.randn(n, DIM, generator=g, device=dev), dim=-1)
    txt = torch.nn.functional.normalize(z + 2.0 * torch.randn(n, DIM, generator=g, device=dev), dim=-1)
    for _ in range(2):
        retrieval_ranks(img, txt)
    med, best = _event_ms(lambda: retrieval_ranks(img, txt), 5, flush)
    t0 = time.perf_counter()
    m = get_clip_metrics(img, txt, LOGIT_SCALE)                      # both directions + D2H of the ranks + numpy metrics
    api_ms = (time.perf_counter() - t0) * 1e3
    img_c, txt_c = img[:n_cpu].cpu(), txt[:n_cpu].cpu()
    t0 = time.perf_counter()
    O.get_clip_metrics(img_c, txt_c, torch.tensor(LOGIT_SCALE))
    cpu_ms = (time.perf_counter() - t0) * 1e3
    tf = 2.0 * n * n * DIM / (med * 1e-3) / 1e12
    return {"workload": "paired retrieval ranks, %d x %d items, dim %d, fp32 (one direction per launch)" % (n, n, DIM),
            "ms": med, "ms_best": best, "fp32_tflops": tf, "frac_of_fp32_fma_peak": tf / (148 * 128 * 2 * 1.965e9 / 1e12),
            "get_clip_metrics_api_ms": api_ms, "R@1": float(m["image_to_text_R@1"]),
            "cpu_reference_path": {"items": n_cpu, "ms": cpu_ms, "note": "oracle restatement of train.py:766-785 on host cores"}}


def bench_eager_gpu(dev, n_global=4096, steps=3):
    """The oracle (eager PyTorch restatement of the reference loss) on the same B200 under bf16 autocast:
    the like-for-like bar, since the reference has no native kernel."""
    from oracle import cosmos_oracle as O
    inp = O.make_features(n_global, DIM, seed=1234)
    leaf = {k: [t.bfloat16().to(dev).requires_grad_(k not in ("t_image", "t_text")) for t in v] for k, v in inp.items()}
    ls = torch.tensor(LOGIT_SCALE, device=dev, requires_grad=True)
    ds = torch.tensor(LOGIT_SCALE, device=dev, requires_grad=True)

    def step():
        for v in leaf.values():
            for t in v:
                t.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = O.cosmos_loss_single(leaf["s_image"], leaf["s_text"], ls, leaf["t_image"], leaf["t_text"], ds,
                                       leaf["s_img_x"], leaf["s_txt_x"])
        (out["distill_loss"] + out["clip_loss"]).backward()

    step(); step()
    med, _ = _event_ms(step, steps)
    return {"workload": "oracle (eager PyTorch) loss head fwd+bwd on the same B200, bf16 autocast, global batch %d" % n_global,
            "ms_per_step": med, "value": n_global / (med * 1e-3), "unit": "samples/s"}


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle (a restatement of the reference's PyTorch loss) on host cores
# ------------------------------------------------------------------------------------------------

def cpu_step_fn(batch: int, dtype=torch.float32):
    from oracle import cosmos_oracle as O
    inp = O.make_features(batch, DIM, seed=1234, dtype=dtype)
    leaf = {k: [t.requires_grad_(k not in ("t_image", "t_text")) for t in v] for k, v in inp.items()}
    ls = torch.tensor(LOGIT_SCALE, requires_grad=True)
    ds = torch.tensor(LOGIT_SCALE, requires_grad=True)

    def step():
        for v in leaf.values():
            for t in v:
                t.grad = None
        out = O.cosmos_loss_single(leaf["s_image"], leaf["s_text"], ls, leaf["t_image"], leaf["t_text"], ds,
                                   leaf["s_img_x"], leaf["s_txt_x"])
        (out["distill_loss"] + out["clip_loss"]).backward()
        return out
    return step


def cpu_baseline(seconds: float, batch: int = 256):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_step_fn(batch)
    step(); step()
    t0 = time.perf_counter()
    n = 0
    while n < 3 or (time.perf_counter() - t0 < seconds and n < 2000):
        step()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": batch * n / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "BASELINE config 1: %d steps of batch %d, dim 512, fp32, oracle (torch CPU restatement of the reference "
                      "loss; the reference is Python and is not present on the GPU box), %.1f s" % (n, batch, dt),
            "ms_per_step": dt / n * 1e3, "torch_threads": cores}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = args.ref_batch
    step = cpu_step_fn(batch)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = batch * args.steps / dt
    line = {
        "impl": "reference", "metric": "loss-head fwd+bwd samples/s at global batch %d" % args.global_batch,
        "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "COSMOS loss head fwd+bwd on host cores; each step is a bounded sample of the global-batch-%d "
                               "workload: one batch-%d step (dim 512, 80 pairs); the full batch needs %d x %d fp32 logits "
                               "per pair and does not fit the time budget" % (args.global_batch, batch, args.global_batch,
                                                                              args.global_batch),
                   "global_batch": args.global_batch, "sample_batch": batch},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": "%d steps of batch %d on %d torch threads" % (args.steps, batch, cores)},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--global-batch", type=int, default=32768)
    ap.add_argument("--ref-batch", type=int, default=1024, help="batch of one reference-arm step (bounded sample)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the EMA / cross-attention / eager-GPU entries")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the small oracle check before the timed loop")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

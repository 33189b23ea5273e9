"""Host side of the block-structured InfoNCE kernels.

`pairs_infonce(rows, cols, scale, ...)` is the mean, over the cartesian product of two feature
lists, of the symmetric InfoNCE loss the reference computes pair by pair in
`ClipLoss.forward` (src/open_clip/loss.py:121-142), including the multi-rank gather semantics of
`gather_features` / `get_logits` (loss.py:21-65, 103-119).  All pairs of one call go through a
few grouped launches (one forward + one backward per chunk of row tensors, see below).

Partitioning: this rank owns `b` rows of every tensor.  The forward computes the
row block  S[local rows of R_i, all N rows of C_j]  for every pair; that yields
  * the complete row log-sum-exp of the local rows,
  * this rank's partial column log-sum-exp for all N columns  -> combined with two all-reduces,
  * the positives (diagonal).
Gradients, two routes:
  * stored exponentials (dim 512, every mode but local_loss): the forward keeps 2^(s2 - max) of every
    logit (bf16) and cosmos_infonce_bwd_e forms dX = G Y from them - no logit is recomputed.  The
    gradients for a unit upstream gradient are formed inside forward(), chunk of row tensors by chunk,
    so only one chunk of exponentials is alive; backward() scales them.  The column side (CLIP term) is
    one GEMM G^T X over the G tiles the same kernel writes out (+ reduce-scatter across ranks).
  * recompute: the backward recomputes the same block and needs nothing else from other ranks for
    the row side; the column side is either the G^T X GEMM over tiles stored by the row pass or the
    same kernel run on the transposed block  S^T[local rows of C_j, all N rows of R_i].

Mode table (probed against the reference on gloo ranks, tests/golden/multirank_w*.pt), N = W*b,
P = number of pairs, R/C = row/column softmax, I = positives:

  local_loss gather_with_grad   returned loss            d features (local)          d scale
  False      False              global  /(2NP)           (R+C-2I)      /(2NP)        global sum (all-reduce)
  False      True               global  /(2NP)           (R+C-2I) * W  /(2NP)        global sum (all-reduce)
  True       True               local   /(2bP)           (R+C-2I)      /(2bP)        local rows of S and of S^T
  True       False              local   /(2bP)           (R - I) only  /(2bP)        local rows of S and of S^T
"""
from __future__ import annotations

import os

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib

LN2 = math.log(2.0)


@dataclass(frozen=True)
class Comm:
    """Which ranks share the loss; mirrors the ctor arguments of the reference losses."""
    rank: int = 0
    world_size: int = 1
    local_loss: bool = False
    gather_with_grad: bool = False
    group: Optional[object] = None

    @property
    def distributed(self) -> bool:
        return self.world_size > 1


# ------------------------------------------------------------------------------------------------
# kernel entry points (ctypes).  Tests that exercise the multi-rank host logic on CPU replace these
# three functions with an oracle-backed emulation; the product has no other implementation.
# ------------------------------------------------------------------------------------------------

def _problem(x: torch.Tensor, y: torch.Tensor, label_offset: int, scale: torch.Tensor) -> _lib.InfoNceProblem:
    gx, n_rows, dim = x.shape
    gy, n_cols, dim_y = y.shape
    if dim != dim_y or x.dtype != y.dtype:
        raise RuntimeError("cosmos_b200: row and column stacks must share dim and dtype")
    return _lib.InfoNceProblem(x=x.data_ptr(), y=y.data_ptr(), gx=gx, gy=gy, n_rows=n_rows, n_cols=n_cols, dim=dim,
                               label_offset=label_offset, dtype=_lib.torch_dtype_code(x.dtype), reserved=0,
                               scale=scale.data_ptr())


def _workspace(prob: _lib.InfoNceProblem, device) -> torch.Tensor:
    nbytes = _lib.lib().cosmos_infonce_workspace_bytes(C.byref(prob))
    if nbytes < 0:
        raise RuntimeError("cosmos_b200: unsupported InfoNCE problem "
                           f"(dim={prob.dim} must be a multiple of 64 and <= 512, dtype bf16/fp16, n_cols >= label_offset + n_rows)")
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _k_fwd(x, y, label_offset, scale, keep_e=False):
    """-> row_lse2 [P, b], diag_raw [P, b], col_lse2 (this rank's rows only) [P, N]; fp32.
    keep_e: also -> (e, off [P, ceil(N / 32), b] fp32): the exponentials 2^(s2 - off) of every logit as bf16, in the tiled
    layout include/cosmos_b200.h describes (stored-exponential route, csrc/infonce_bwd_e.cu)."""
    dev = x.device
    prob = _problem(x, y, label_offset, scale)
    P = prob.gx * prob.gy
    row_lse2 = torch.empty(P, prob.n_rows, dtype=torch.float32, device=dev)
    diag_raw = torch.empty(P, prob.n_rows, dtype=torch.float32, device=dev)
    col_lse2 = torch.empty(P, prob.n_cols, dtype=torch.float32, device=dev)
    ws = _workspace(prob, dev)
    if not keep_e:
        st = _lib.lib().cosmos_infonce_fwd(C.byref(prob), row_lse2.data_ptr(), diag_raw.data_ptr(), col_lse2.data_ptr(),
                                           ws.data_ptr(), ws.numel(), dev.index, torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(st, "infonce_fwd")
        return row_lse2, diag_raw, col_lse2
    e = torch.empty(_lib.lib().cosmos_infonce_e_bytes(C.byref(prob)) // 2, dtype=torch.bfloat16, device=dev)
    off = torch.empty(P, (prob.n_cols + 31) // 32, prob.n_rows, dtype=torch.float32, device=dev)
    st = _lib.lib().cosmos_infonce_fwd_e(C.byref(prob), row_lse2.data_ptr(), diag_raw.data_ptr(), col_lse2.data_ptr(),
                                         e.data_ptr(), off.data_ptr(), ws.data_ptr(), ws.numel(), dev.index,
                                         torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(st, "infonce_fwd_e")
    return row_lse2, diag_raw, col_lse2, e, off


def _k_bwd_e(x, y, label_offset, scale, e, off, row_lse2, col_lse2, a_row, a_col, s_row, s_col, weight, upstream,
             want_dscale, g_out=None):
    """Backward from the stored exponentials: -> dx [gx, b, 512] in x.dtype, dscale fp32 [1] (or None).  No logit is
    recomputed; x is only read for d(scale) = sum_r <x_r, (G y)_r>."""
    dev = x.device
    prob = _problem(x, y, label_offset, scale)
    dx = torch.empty_like(x)
    dscale = torch.empty(1, dtype=torch.float32, device=dev) if want_dscale else None
    ws = torch.empty(max(4 * prob.gx * ((prob.n_rows + 127) // 128) * 8, 256) + 256, dtype=torch.uint8, device=dev)
    st = _lib.lib().cosmos_infonce_bwd_e(C.byref(prob), e.data_ptr(), off.data_ptr(), row_lse2.data_ptr(),
                                         col_lse2.data_ptr(), a_row, a_col, s_row, s_col, weight,
                                         upstream.data_ptr(), dx.data_ptr(), dscale.data_ptr() if want_dscale else None,
                                         g_out.data_ptr() if g_out is not None else None,
                                         g_out.stride(0) if g_out is not None else 0, ws.data_ptr(), ws.numel(), dev.index,
                                         torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(st, "infonce_bwd_e")
    return dx, dscale


def _k_loss_sums(x, y, label_offset, scale, row_lse2, diag_raw, col_lse2):
    """-> [P, 2] fp32: natural-log row-CE sum over local rows, column-CE sum over this rank's diagonal columns."""
    dev = x.device
    prob = _problem(x, y, label_offset, scale)
    out = torch.empty(prob.gx * prob.gy, 2, dtype=torch.float32, device=dev)
    st = _lib.lib().cosmos_infonce_loss_sums(C.byref(prob), row_lse2.data_ptr(), diag_raw.data_ptr(), col_lse2.data_ptr(),
                                             1, 1, out.data_ptr(), None, dev.index,
                                             torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(st, "infonce_loss_sums")
    return out


def _k_bwd(x, y, label_offset, scale, row_lse2, col_lse2, a_row, a_col, s_row, s_col, weight, upstream, want_dx, want_dscale,
           g_out=None):
    """-> dx [gx, b, D] in x.dtype (or None), dscale fp32 [1] (or None).
    g_out [gx * b, >= gy * N] (x.dtype): also receives G itself, block (i, j) at rows i * b, columns j * N."""
    dev = x.device
    prob = _problem(x, y, label_offset, scale)
    dx = torch.empty_like(x) if want_dx else None
    dscale = torch.empty(1, dtype=torch.float32, device=dev) if want_dscale else None
    ws = _workspace(prob, dev)
    common = (C.byref(prob), row_lse2.data_ptr(), col_lse2.data_ptr(), a_row, a_col, s_row, s_col, weight, upstream.data_ptr(),
              dx.data_ptr() if want_dx else None, dscale.data_ptr() if want_dscale else None)
    tail = (ws.data_ptr(), ws.numel(), dev.index, torch.cuda.current_stream(dev).cuda_stream)
    if g_out is None:
        st = _lib.lib().cosmos_infonce_bwd(*common, *tail)
    else:
        st = _lib.lib().cosmos_infonce_bwd_g(*common, g_out.data_ptr(), g_out.stride(0), *tail)
    _lib.check(st, "infonce_bwd")
    return dx, dscale


def _k_colgrad(g: torch.Tensor, x2d: torch.Tensor, n_c: int, n_cols: int) -> torch.Tensor:
    """sum over rows of G^T x: g [R, >= n_c * n_cols] (block j at columns j * n_cols), x2d [R, D] -> fp32 [n_c, n_cols, D].
    A tcgen05 GEMM with both operands read in their stored layout (contraction over the rows, like a weight gradient)."""
    from .pooler import _gemm
    R, D = x2d.shape
    out = torch.empty(n_c, n_cols, D, dtype=torch.float32, device=g.device)
    # the column blocks are adjacent in g, so [n_c * n_cols, D] = g^T x2d is ONE GEMM (more tiles per wave than n_c small ones)
    _gemm(g, x2d, out.view(n_c * n_cols, D), n_c * n_cols, D, R, g.stride(0), x2d.stride(0), False, False)
    return out


# ------------------------------------------------------------------------------------------------
# layout helpers
# ------------------------------------------------------------------------------------------------

def compute_dtype(dtype: torch.dtype) -> torch.dtype:
    """16-bit inputs are consumed as they are; fp32 features are rounded to bf16 for the tensor cores
    (the arithmetic type of this path), accumulation and softmax stay fp32."""
    return dtype if dtype in (torch.bfloat16, torch.float16) else torch.bfloat16


def stack_views(tensors: Sequence[torch.Tensor], dtype: torch.dtype) -> torch.Tensor:
    """[n, b, D] stack of same-shape 2-D tensors.  Zero-copy when they are consecutive views of one
    buffer (the reference's `.chunk()` outputs, src/training/train.py:171-182)."""
    t0 = tensors[0]
    if t0.dim() != 2:
        raise RuntimeError(f"cosmos_b200: features must be [batch, dim] matrices, got shape {tuple(t0.shape)}")
    b, d = t0.shape
    for t in tensors:
        if t.shape != t0.shape or t.dtype != t0.dtype or t.device != t0.device:
            raise RuntimeError("cosmos_b200: every feature tensor of a list must have the same shape, dtype and device")
    step = b * d * t0.element_size()
    if (t0.dtype == dtype and all(t.is_contiguous() for t in tensors) and t0.data_ptr() % 16 == 0
            and all(t.data_ptr() == t0.data_ptr() + k * step for k, t in enumerate(tensors))
            and t0.untyped_storage().data_ptr() == tensors[-1].untyped_storage().data_ptr()):
        return torch.as_strided(t0.detach(), (len(tensors), b, d), (b * d, d, 1))
    return torch.stack([t.detach() for t in tensors]).to(dtype)


def gather_stack(local: torch.Tensor, comm: Comm) -> torch.Tensor:
    """[n, b, D] on every rank -> [n, W*b, D], rows ordered rank-major as torch.cat(all_gather) does."""
    if not comm.distributed:
        return local
    n, b, d = local.shape
    out = torch.empty(comm.world_size * n, b, d, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=comm.group)
    return out.view(comm.world_size, n, b, d).permute(1, 0, 2, 3).reshape(n, comm.world_size * b, d)


class GatherHandle:
    """An all-gather of one [n, b, D] stack started with async_op=True: the transfer runs on NCCL's stream while the
    caller keeps launching kernels; `get()` makes the current stream wait for it and returns the [n, W*b, D] layout."""

    def __init__(self, local: torch.Tensor, comm: Comm):
        self.comm, self.shape = comm, tuple(local.shape)
        if comm.distributed:
            n, b, d = local.shape
            self.out = torch.empty(comm.world_size * n, b, d, dtype=local.dtype, device=local.device)
            self.work = dist.all_gather_into_tensor(self.out, local.contiguous(), group=comm.group, async_op=True)
        else:
            self.out, self.work = local, None
        self._result = None

    def get(self) -> torch.Tensor:
        if self._result is None:
            if self.work is not None:
                self.work.wait()
                n, b, d = self.shape
                W = self.comm.world_size
                self._result = self.out.view(W, n, b, d).permute(1, 0, 2, 3).reshape(n, W * b, d)
            else:
                self._result = self.out
        return self._result


class Prefetch:
    """Gathers started ahead of time by the caller (COSMOSLoss starts every list's all-gather before the first kernel,
    so only the first one is exposed); keyed by the data pointer of the local stack."""

    def __init__(self):
        self.handles = {}

    def start(self, tensors: Sequence[torch.Tensor], comm: Comm) -> None:
        if not comm.distributed:
            return
        tensors = list(tensors)
        st = stack_views(tensors, compute_dtype(tensors[0].dtype))
        self.handles[tuple(id(t) for t in tensors)] = (st, GatherHandle(st, comm), tensors)   # keeps the ids alive


def _allreduce_lse2(lse2: torch.Tensor, comm: Comm) -> torch.Tensor:
    """log2-sum-exp2 over ranks of per-rank partial column statistics (max all-reduce + sum all-reduce)."""
    if not comm.distributed:
        return lse2
    m = lse2.clone()
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=comm.group)
    s = torch.exp((lse2 - m) * LN2)          # exp / log instead of exp2 / log2: the latter are NVRTC-jitted by torch
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=comm.group)
    return m + torch.log(s) / LN2


def _gather_rows(t: torch.Tensor, comm: Comm) -> torch.Tensor:
    """[P, b] per rank -> [P, W*b]."""
    if not comm.distributed:
        return t
    P, b = t.shape
    out = torch.empty(comm.world_size * P, b, dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=comm.group)
    return out.view(comm.world_size, P, b).permute(1, 0, 2).reshape(P, comm.world_size * b)


def _swap_pairs(t: torch.Tensor, n_r: int, n_c: int) -> torch.Tensor:
    """[n_r * n_c, L] with pair = i * n_c + j  ->  [n_c * n_r, L] with pair = j * n_r + i."""
    return t.view(n_r, n_c, -1).transpose(0, 1).reshape(n_c * n_r, -1).contiguous()


# ------------------------------------------------------------------------------------------------
# autograd node
# ------------------------------------------------------------------------------------------------

_G_STORE_MAX_BYTES = int(float(os.environ.get("COSMOS_B200_GSTORE_MAX_GB", "48")) * (1 << 30))
# below ~1 GB of tiles the second sweep is short and the extra launches cost what they save (measured at N = 4096 on one
# GPU: 4.67 ms with the GEMM route, 4.54 ms with the second sweep; at N = 32768: 265 vs 283 ms)
_G_STORE_MIN_BYTES = int(float(os.environ.get("COSMOS_B200_GSTORE_MIN_GB", "1")) * (1 << 30))


def _g_store_ok(x_r: torch.Tensor, y_c: torch.Tensor) -> bool:
    """Can the row pass keep its G tiles for the column-side GEMM?  dim 512 (the cluster kernel), 16-byte pieces that do
    not straddle the last column, and a buffer that fits comfortably (n_r * b x n_c * N 16-bit values)."""
    n_r, b, dim = x_r.shape
    n_c, n_all, _ = y_c.shape
    if dim != 512 or (n_all & 7) != 0 or _G_STORE_MAX_BYTES <= 0:
        return False
    nbytes = n_r * b * n_c * n_all * x_r.element_size()
    if nbytes > _G_STORE_MAX_BYTES or nbytes < _G_STORE_MIN_BYTES:
        return False
    if x_r.is_cuda:
        free, _total = torch.cuda.mem_get_info(x_r.device)
        free += torch.cuda.memory_reserved(x_r.device) - torch.cuda.memory_allocated(x_r.device)   # cached blocks are reusable
        if 2 * nbytes > free:
            return False
    return True


def _reduce_scatter_rows(d_all: torch.Tensor, comm: Comm, b: int) -> torch.Tensor:
    """[n_c, W * b, D] per-rank partial sums -> [n_c, b, D]: the sum over ranks of this rank's rows."""
    if not comm.distributed:
        return d_all
    W, rank = comm.world_size, comm.rank
    n_c, _, D = d_all.shape
    if dist.get_backend(comm.group) == "nccl":
        send = d_all.view(n_c, W, b, D).transpose(0, 1).contiguous()          # rank-major chunks
        out = torch.empty(n_c, b, D, dtype=d_all.dtype, device=d_all.device)
        dist.reduce_scatter_tensor(out, send, op=dist.ReduceOp.SUM, group=comm.group)
        return out
    dist.all_reduce(d_all, op=dist.ReduceOp.SUM, group=comm.group)           # backends without reduce-scatter (gloo, CPU tests)
    return d_all[:, rank * b:(rank + 1) * b].contiguous()


_E_STORE_MAX_BYTES = int(float(os.environ.get("COSMOS_B200_ESTORE_MAX_GB", "40")) * (1 << 30))
# small problems: the recompute kernels finish in microseconds and one launch per group beats the chunk loop
_E_STORE_MIN_BYTES = int(float(os.environ.get("COSMOS_B200_ESTORE_MIN_GB", "0.25")) * (1 << 30))


_e_chunk_cache: dict = {}


def _e_store_chunk(x_r: torch.Tensor, y_c: torch.Tensor, comm: Comm) -> int:
    """Row tensors per pass of the stored-exponential route (csrc/infonce_bwd_e.cu), 0 = use the recompute kernels.

    The route needs dim 512 (a [128 x 512] fp32 dX accumulator is exactly the tensor memory of an SM) and gradient mixes
    that weigh d(scale) like dX (every mode but local_loss).  The forward keeps 2 bytes per logit; the gradients are formed
    right away, chunk of row tensors by chunk, so only one chunk of exponentials is alive at a time."""
    n_r, b, dim = x_r.shape
    n_c, n_all, _ = y_c.shape
    if dim != 512 or _E_STORE_MAX_BYTES <= 0 or (comm.distributed and comm.local_loss):
        return 0
    per_tensor = n_c * (-(-b // 128) * 128) * (-(-n_all // 128) * 128) * 2
    if per_tensor * n_r < _E_STORE_MIN_BYTES:
        return 0
    key = (n_r, b, n_c, n_all, x_r.device, comm.world_size, id(comm.group))
    chunk = _e_chunk_cache.get(key)
    if chunk is None:
        # decided once per problem shape: cudaMemGetInfo stalls the launch queue, which a per-step call would pay every step
        budget = _E_STORE_MAX_BYTES
        if x_r.is_cuda:
            free, _total = torch.cuda.mem_get_info(x_r.device)
            free += torch.cuda.memory_reserved(x_r.device) - torch.cuda.memory_allocated(x_r.device)   # cached blocks are reusable
            budget = min(budget, free // 3)      # exponentials + (CLIP group) the G tiles of the same chunk + headroom
        most = min(n_r, budget // per_tensor)
        if most <= 0:
            chunk = 0
        else:
            passes = -(-n_r // most)
            chunk = -(-n_r // passes)            # equal passes: 16 tensors at most 5 at a time -> 4 + 4 + 4 + 4, not 5 + 5 + 5 + 1
        if comm.distributed:
            # every rank must make the same number of passes (each pass all-reduces its column statistics): agree on the
            # smallest chunk once per shape; ranks see the same shapes in the same order, so they all arrive here together
            agreed = torch.tensor([chunk], dtype=torch.int32, device=x_r.device if x_r.is_cuda else "cpu")
            dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=comm.group)
            chunk = int(agreed.item())
        if len(_e_chunk_cache) > 64:
            _e_chunk_cache.clear()
        _e_chunk_cache[key] = chunk
    return chunk


class _PairsInfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scale: torch.Tensor, comm: Comm, n_r: int, prefetch, *feats: torch.Tensor):
        rows, cols = feats[:n_r], feats[n_r:]
        n_c = len(cols)
        dev = rows[0].device
        dt = compute_dtype(rows[0].dtype)
        x_r = stack_views(rows, dt)                       # [n_r, b, D]
        pre_c = pre_r = None
        if prefetch is not None and comm.distributed:
            pre_c = prefetch.handles.get(tuple(id(t) for t in cols))
            pre_r = prefetch.handles.get(tuple(id(t) for t in rows))
        x_c = pre_c[0] if pre_c is not None else stack_views(cols, dt)   # [n_c, b, D]
        if pre_r is not None:
            x_r = pre_r[0]
        b = x_r.shape[1]
        if x_c.shape[1] != b:
            raise RuntimeError("cosmos_b200: both feature lists must have the same batch size")
        W = comm.world_size
        N = W * b
        off = comm.rank * b if comm.distributed else 0
        scale_f = scale.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()

        y_c = pre_c[1].get() if pre_c is not None else gather_stack(x_c, comm)     # [n_c, N, D]
        P = n_r * n_c
        need_scale = ctx.needs_input_grad[0]
        need_rows = any(ctx.needs_input_grad[4:4 + n_r])
        need_cols = any(ctx.needs_input_grad[4 + n_r:])
        # (the column-side gradient of this route is a GEMM over stored G tiles, whose 16-byte pieces need N % 8 == 0)
        chunk = _e_store_chunk(x_r, y_c, comm) if (need_rows and not (need_cols and N % 8 != 0)) else 0
        ctx.eager = chunk > 0
        if ctx.eager:
            # Stored-exponential route: per chunk of row tensors, forward (statistics + 2^(s2 - max) of every logit, bf16) and,
            # as soon as the chunk's column statistics are complete, the gradients for a unit upstream gradient - dX = G Y is
            # the only contraction left, no logit is recomputed.  backward() only scales by the upstream gradient.
            boost = float(W) if (comm.distributed and comm.gather_with_grad) else 1.0
            weight = boost / (2.0 * N * P)
            # The unit gradients are held in the feature dtype.  Under GradScaler (fp16 features) the upstream factor exists
            # precisely because |dX| ~ weight * scale would underflow fp16, so they are formed with a power-of-two stand-in
            # for it (exact in every dtype; |dX * pre| <= 2 * n_c * scale / 64) and backward() multiplies by upstream / pre.
            pre = 2.0 ** (math.floor(math.log2(1.0 / weight)) - 6)
            one = torch.full((1,), pre, dtype=torch.float32, device=dev)
            row_parts, diag_parts, col_parts, dx_parts = [], [], [], []
            ds_unit = None
            d_all = None
            for i0 in range(0, n_r, chunk):
                xs = x_r[i0:i0 + chunk]
                r_, d_, c_part, e_, o_ = _k_fwd(xs, y_c, off, scale_f, True)
                c_ = _allreduce_lse2(c_part, comm)
                g_tiles = torch.empty(xs.shape[0] * b, n_c * N, dtype=x_r.dtype, device=dev) if need_cols else None
                dx_, ds_ = _k_bwd_e(xs, y_c, off, scale_f, e_, o_, r_, c_, 1.0, 1.0, 1.0 / boost, 1.0 / boost, weight, one,
                                    need_scale, g_tiles)
                del e_, o_
                if g_tiles is not None:
                    part = _k_colgrad(g_tiles, xs.reshape(xs.shape[0] * b, xs.shape[2]), n_c, N)    # [n_c, N, D] fp32
                    d_all = part if d_all is None else d_all.add_(part)
                    del g_tiles
                if ds_ is not None:
                    ds_unit = ds_ if ds_unit is None else ds_unit + ds_
                row_parts.append(r_); diag_parts.append(d_); col_parts.append(c_); dx_parts.append(dx_)
            row_lse2, diag_raw, col_lse2 = (torch.cat(t) if len(t) > 1 else t[0] for t in (row_parts, diag_parts, col_parts))
            ctx.unit = (torch.cat(dx_parts) if len(dx_parts) > 1 else dx_parts[0], ds_unit, d_all, weight, pre)
        else:
            row_lse2, diag_raw, col_lse2_part = _k_fwd(x_r, y_c, off, scale_f)
            col_lse2 = _allreduce_lse2(col_lse2_part, comm)   # global over all rows
        sums = _k_loss_sums(x_r, y_c, off, scale_f, row_lse2, diag_raw, col_lse2)   # [P, 2]
        total = sums.sum()
        if comm.distributed and not comm.local_loss:
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=comm.group)
            loss = total / (2.0 * N * P)
        else:
            loss = total / (2.0 * b * P)

        ctx.comm, ctx.n_r, ctx.n_c, ctx.b, ctx.off = comm, n_r, n_c, b, off
        ctx.pre_r = pre_r[1] if pre_r is not None else None
        ctx.in_dtypes = [t.dtype for t in feats]
        ctx.scale_dtype = scale.dtype
        if ctx.eager:
            ctx.save_for_backward(scale_f)
        else:
            ctx.save_for_backward(x_r, x_c, y_c, scale_f, row_lse2, col_lse2)
        return loss

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        if ctx.eager:
            return _PairsInfoNCE._backward_eager(ctx, g)
        x_r, x_c, y_c, scale_f, row_lse2, col_lse2 = ctx.saved_tensors
        comm, n_r, n_c, b, off = ctx.comm, ctx.n_r, ctx.n_c, ctx.b, ctx.off
        W = comm.world_size
        N, P = W * b, n_r * n_c
        need_scale = ctx.needs_input_grad[0]
        need_rows = any(ctx.needs_input_grad[4:4 + n_r])
        need_cols = any(ctx.needs_input_grad[4 + n_r:])
        up = g.detach().to(torch.float32).reshape(1).contiguous()

        local = comm.distributed and comm.local_loss
        if local:
            weight = 1.0 / (2.0 * b * P)
            a = (1.0, 1.0) if comm.gather_with_grad else (1.0, 0.0)
            s_mix = (1.0, 0.0)                      # each side reports its own rows' part of dscale
        else:
            boost = float(W) if (comm.distributed and comm.gather_with_grad) else 1.0
            weight = boost / (2.0 * N * P)
            a = (1.0, 1.0)
            s_mix = (1.0 / boost, 1.0 / boost)

        d_rows = d_scale = d_cols = None
        g_tiles = None
        if need_rows and need_cols and not local and _g_store_ok(x_r, y_c):
            # non-local modes weigh the row and the column softmax equally, so the gradient matrix of the transposed block is
            # G^T: keep G from the row pass (bf16, the operand precision of both passes anyway) and turn the column side
            # into one GEMM instead of a second sweep that recomputes every logit
            g_tiles = torch.empty(n_r * b, n_c * N, dtype=x_r.dtype, device=x_r.device)
        if need_rows or need_scale:
            d_rows, d_scale = _k_bwd(x_r, y_c, off, scale_f, row_lse2, col_lse2, a[0], a[1], s_mix[0], s_mix[1], weight, up,
                                     need_rows, need_scale, g_tiles)
        if g_tiles is not None:
            d_all = _k_colgrad(g_tiles, x_r.reshape(n_r * b, x_r.shape[2]), n_c, N)        # [n_c, N, D] fp32, this rank's rows
            del g_tiles
            d_loc = _reduce_scatter_rows(d_all, comm, b)                                     # [n_c, b, D], all ranks' rows
            d_cols = (d_loc * (up * scale_f.reshape(1) * weight)).to(x_c.dtype)
        elif need_cols or (local and need_scale):
            # transposed block: rows = local column-side tensors, columns = all rows of the row side
            y_r = ctx.pre_r.get() if ctx.pre_r is not None else gather_stack(x_r, comm)     # [n_r, N, D]
            row_lse2_all = _gather_rows(row_lse2, comm)                     # [P, N]
            t_row = _swap_pairs(col_lse2[:, off:off + b], n_r, n_c)         # row LSE of S^T for my rows
            t_col = _swap_pairs(row_lse2_all, n_r, n_c)                     # column LSE of S^T (global)
            d_cols, d_scale_t = _k_bwd(x_c, y_r, off, scale_f, t_row, t_col, a[0], a[1], s_mix[0], s_mix[1], weight, up,
                                       need_cols, local and need_scale)
            if d_scale_t is not None:
                d_scale = d_scale_t if d_scale is None else d_scale + d_scale_t
        if need_scale and comm.distributed and not comm.local_loss:
            dist.all_reduce(d_scale, op=dist.ReduceOp.SUM, group=comm.group)

        grads: List[Optional[torch.Tensor]] = []
        for k in range(n_r):
            grads.append(d_rows[k].to(ctx.in_dtypes[k]) if (d_rows is not None and ctx.needs_input_grad[4 + k]) else None)
        for k in range(n_c):
            grads.append(d_cols[k].to(ctx.in_dtypes[n_r + k])
                         if (d_cols is not None and ctx.needs_input_grad[4 + n_r + k]) else None)
        g_scale = d_scale.reshape(()).to(ctx.scale_dtype) if need_scale else None
        return (g_scale, None, None, None, *grads)


    @staticmethod
    def _backward_eager(ctx, g: torch.Tensor):
        """The gradients for a unit upstream gradient were formed in forward(): scale them (and finish the collectives)."""
        (scale_f,) = ctx.saved_tensors
        comm, n_r, n_c, b = ctx.comm, ctx.n_r, ctx.n_c, ctx.b
        dx_unit, ds_unit, d_all, weight, pre = ctx.unit
        up = g.detach().to(torch.float32).reshape(1)
        grads: List[Optional[torch.Tensor]] = []
        d_rows = (dx_unit.float() * (up / pre)).to(dx_unit.dtype) if any(ctx.needs_input_grad[4:4 + n_r]) else None
        for k in range(n_r):
            grads.append(d_rows[k].to(ctx.in_dtypes[k]) if (d_rows is not None and ctx.needs_input_grad[4 + k]) else None)
        d_cols = None
        if d_all is not None:
            d_loc = _reduce_scatter_rows(d_all, comm, b)                                  # [n_c, b, D], all ranks' rows
            d_cols = (d_loc * (up * scale_f.reshape(1) * weight)).to(dx_unit.dtype)
        for k in range(n_c):
            grads.append(d_cols[k].to(ctx.in_dtypes[n_r + k])
                         if (d_cols is not None and ctx.needs_input_grad[4 + n_r + k]) else None)
        g_scale = None
        if ctx.needs_input_grad[0]:
            d_scale = ds_unit * (up / pre)
            if comm.distributed:
                dist.all_reduce(d_scale, op=dist.ReduceOp.SUM, group=comm.group)
            g_scale = d_scale.reshape(()).to(ctx.scale_dtype)
        return (g_scale, None, None, None, *grads)


def pairs_infonce(rows: Sequence[torch.Tensor], cols: Sequence[torch.Tensor], scale, comm: Comm = Comm(),
                  prefetch: Optional["Prefetch"] = None) -> torch.Tensor:
    """Mean symmetric InfoNCE over all (row tensor, column tensor) pairs; a 0-dim fp32 tensor.

    The loss is symmetric in its two lists, so callers pass the longer / gradient-carrying list as
    `rows` (it is never communicated) and the shorter one as `cols` (it is all-gathered)."""
    rows, cols = list(rows), list(cols)
    if not rows or not cols:
        raise RuntimeError("cosmos_b200: empty feature list")
    for t in rows + cols:
        _lib.require_cuda(t, "feature tensor")
    if not isinstance(scale, torch.Tensor):
        scale = torch.tensor(float(scale), dtype=torch.float32, device=rows[0].device)
    return _PairsInfoNCE.apply(scale, comm, len(rows), prefetch, *rows, *cols)

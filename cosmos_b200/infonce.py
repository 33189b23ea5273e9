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


def _k_fwd(x, y, label_offset, scale, keep_e=False, out=None):
    """-> row_lse2 [P, b], diag_raw [P, b], col_lse2 (this rank's rows only) [P, N]; fp32.
    keep_e: also -> (e, off [P, ceil(N / 32), b] fp32): the exponentials 2^(s2 - off) of every logit as bf16, in the tiled
    layout include/cosmos_b200.h describes (stored-exponential route, csrc/infonce_bwd_e2.cu).
    out: (row_lse2, diag_raw, col_lse2) to write into (contiguous slices of per-group buffers) instead of new tensors."""
    dev = x.device
    prob = _problem(x, y, label_offset, scale)
    P = prob.gx * prob.gy
    if out is not None:
        row_lse2, diag_raw, col_lse2 = out
    else:
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


def _k_bwd_e(x, y, label_offset, scale, e, off, diag_raw, row_lse2, col_lse2, a_row, a_col, s_row, s_col, weight, upstream,
             want_dscale, g_out=None, dx_out=None):
    """Backward from the stored exponentials: -> dx [gx, b, 512] in x.dtype, dscale fp32 [1] (or None).  No logit is
    recomputed; x is only read for d(scale) = sum_r <x_r, (G y)_r>.  dx_out: where to write dx (a slice of the group's buffer)."""
    dev = x.device
    prob = _problem(x, y, label_offset, scale)
    dx = dx_out if dx_out is not None else torch.empty_like(x)
    dscale = torch.empty(1, dtype=torch.float32, device=dev) if want_dscale else None
    ws = torch.empty(max(int(_lib.lib().cosmos_infonce_bwd_e_workspace_bytes(C.byref(prob), dev.index)), 256), dtype=torch.uint8,
                     device=dev)
    st = _lib.lib().cosmos_infonce_bwd_e(C.byref(prob), e.data_ptr(), off.data_ptr(), diag_raw.data_ptr(), row_lse2.data_ptr(),
                                         col_lse2.data_ptr(), a_row, a_col, s_row, s_col, weight,
                                         upstream.data_ptr(), dx.data_ptr(), dscale.data_ptr() if want_dscale else None,
                                         g_out.data_ptr() if g_out is not None else None,
                                         g_out.stride(0) if g_out is not None else 0, ws.data_ptr(), ws.numel(), dev.index,
                                         torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(st, "infonce_bwd_e")
    return dx, dscale


def _k_bwd_e_cols(x, y, label_offset, scale, e, off, diag_raw, row_lse2, col_lse2, a_row, a_col, rank_major=0):
    """Column side of the stored-exponential route: -> fp32 [gy, N, 512], sum over the row tensors and THIS rank's rows of
    G^T x at unit scale (csrc/infonce_bwd_e2t.cu) - no G tile is written to memory.
    rank_major = W > 0: -> [W, gy, N / W, 512] instead, the layout a reduce-scatter sends (the sum over the kernel's row-sweep
    slices and the reordering are one pass)."""
    dev = x.device
    prob = _problem(x, y, label_offset, scale)
    splits = _lib.lib().cosmos_infonce_bwd_e_cols_splits(C.byref(prob), dev.index)
    if splits < 1:
        raise RuntimeError("cosmos_b200: unsupported problem for the column-side stored-exponential backward")
    dy = torch.empty(splits, prob.gy, prob.n_cols, prob.dim, dtype=torch.float32, device=dev)
    st = _lib.lib().cosmos_infonce_bwd_e_cols(C.byref(prob), e.data_ptr(), off.data_ptr(), diag_raw.data_ptr(), row_lse2.data_ptr(),
                                              col_lse2.data_ptr(),
                                              a_row, a_col, dy.data_ptr(), splits, dev.index,
                                              torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(st, "infonce_bwd_e_cols")
    if rank_major:
        W = rank_major
        src = dy.view(splits, prob.gy, W, prob.n_cols // W, prob.dim).permute(0, 2, 1, 3, 4)
        if splits == 1:
            return src[0].contiguous()
        return torch.sum(src, dim=0, out=torch.empty(src.shape[1:], dtype=torch.float32, device=dev))
    return dy[0] if splits == 1 else dy.sum(0)


def _k_scale16(src: torch.Tensor, num: torch.Tensor, den: float) -> torch.Tensor:
    """src * (num / den) for a 16-bit tensor and a device scalar num (fp32 [1]); a new tensor of src's dtype."""
    dev = src.device
    dst = torch.empty_like(src)
    st = _lib.lib().cosmos_scale16(src.data_ptr(), dst.data_ptr(), num.data_ptr(), float(den), _lib.torch_dtype_code(src.dtype),
                                   src.numel(), dev.index, torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(st, "scale16")
    return dst


def _k_lse2_merge(parts: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """parts [W, ...] fp32 (every rank's partial column log2-sum-exp, all-gathered) -> out [...]: their log2-sum-exp2."""
    dev = parts.device
    st = _lib.lib().cosmos_lse2_merge(parts.data_ptr(), out.data_ptr(), parts.shape[0], out.numel(), dev.index,
                                      torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(st, "lse2_merge")
    return out


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
    local = local.contiguous()
    # one collective per tensor, straight into its [W*b, D] slice: the rows arrive rank-major, no reordering copy afterwards
    out = torch.empty(n, comm.world_size * b, d, dtype=local.dtype, device=local.device)
    for k in range(n):
        dist.all_gather_into_tensor(out[k], local[k], group=comm.group)
    return out


class GatherHandle:
    """An all-gather of one [n, b, D] stack started with async_op=True: the transfer runs on NCCL's stream while the
    caller keeps launching kernels; `get()` makes the current stream wait for it and returns the [n, W*b, D] layout."""

    def __init__(self, local: torch.Tensor, comm: Comm):
        self.comm, self.shape = comm, tuple(local.shape)
        if comm.distributed:
            n, b, d = local.shape
            local = local.contiguous()
            # one collective per tensor, straight into its [W*b, D] slice of the [n, W*b, D] result: the rows arrive rank-major
            # (a single gather of the whole stack lands as [W, n, b, D] and needs a reordering copy on the compute stream)
            self.out = torch.empty(n, comm.world_size * b, d, dtype=local.dtype, device=local.device)
            self.work = [dist.all_gather_into_tensor(self.out[k], local[k], group=comm.group, async_op=True) for k in range(n)]
        else:
            self.out, self.work = local, None
        self._waited = False

    def get(self) -> torch.Tensor:
        if self.work is not None and not self._waited:
            for w in self.work:
                w.wait()
            self._waited = True
        return self.out


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


class _LseGather:
    """Per-rank partial column statistics [P, N] -> the global ones: ONE all-gather (started with async_op=True, so the
    caller keeps launching kernels behind it) and one merge kernel, instead of a MAX all-reduce, an eager exp, a SUM
    all-reduce and an eager log."""

    def __init__(self, part: torch.Tensor, comm: Comm):
        self.part, self.comm = part, comm
        # (flat leading dimension: the gloo backend of the CPU tests accepts no other output shape)
        self.gathered = torch.empty(comm.world_size * part.shape[0], *part.shape[1:], dtype=part.dtype, device=part.device)
        self.work = dist.all_gather_into_tensor(self.gathered, part, group=comm.group, async_op=True)

    def finish(self, out: torch.Tensor) -> torch.Tensor:
        self.work.wait()                       # the current stream waits; the host does not (NCCL)
        return _k_lse2_merge(self.gathered.view(self.comm.world_size, *self.part.shape), out)


def _allreduce_lse2(lse2: torch.Tensor, comm: Comm) -> torch.Tensor:
    """log2-sum-exp2 over ranks of per-rank partial column statistics."""
    if not comm.distributed:
        return lse2
    return _LseGather(lse2.contiguous(), comm).finish(torch.empty_like(lse2))


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


def _rank_major_ok(comm: Comm) -> bool:
    """Column-side partials can be produced in the reduce-scatter's send layout right away (NCCL groups only)."""
    return comm.distributed and dist.get_backend(comm.group) == "nccl"


def _reduce_scatter_rows(d_all: torch.Tensor, comm: Comm, b: int, async_op: bool = False, rank_major: bool = False):
    """[n_c, W * b, D] per-rank partial sums (rank_major: already [W, n_c, b, D]) -> [n_c, b, D]: the sum over ranks of this
    rank's rows.  async_op: -> (result, work) with the collective still running (work is None when nothing was started)."""
    if not comm.distributed:
        return (d_all, None) if async_op else d_all
    W, rank = comm.world_size, comm.rank
    if rank_major:
        _, n_c, _, D = d_all.shape
    else:
        n_c, _, D = d_all.shape
    if dist.get_backend(comm.group) == "nccl":
        send = d_all if rank_major else d_all.view(n_c, W, b, D).transpose(0, 1).contiguous()          # rank-major chunks
        out = torch.empty(n_c, b, D, dtype=d_all.dtype, device=d_all.device)
        work = dist.reduce_scatter_tensor(out, send, op=dist.ReduceOp.SUM, group=comm.group, async_op=async_op)
        return (out, work) if async_op else out
    dist.all_reduce(d_all, op=dist.ReduceOp.SUM, group=comm.group)           # backends without reduce-scatter (gloo, CPU tests)
    out = d_all[:, rank * b:(rank + 1) * b].contiguous()
    return (out, None) if async_op else out


_E_STORE_MAX_BYTES = int(float(os.environ.get("COSMOS_B200_ESTORE_MAX_GB", "96")) * (1 << 30))
# small problems: the recompute kernels finish in microseconds and one launch per group beats the chunk loop
_E_STORE_MIN_BYTES = int(float(os.environ.get("COSMOS_B200_ESTORE_MIN_GB", "0.25")) * (1 << 30))


# COSMOS_B200_COLS=gemm: column-side gradient of the stored-exponential route through stored G tiles + a GEMM (diagnostics)
_COLS_VIA_GEMM = os.environ.get("COSMOS_B200_COLS", "") == "gemm"

_e_chunk_cache: dict = {}


def _reusable_bytes(device) -> int:
    """Free device memory as this process can use it: cudaMemGetInfo plus the allocator's cached blocks."""
    free, _total = torch.cuda.mem_get_info(device)
    return free + torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)


_SM_PAIRS = 74      # CTA pairs resident at a time (148 SMs): the backward kernel runs one pair per 256 rows of one row tensor


def _wave_aware_chunk(n_r: int, most: int, row_tiles: int) -> int:
    """Row tensors per pass: at most `most`, equal passes (16 tensors at most 5 at a time -> 4 + 4 + 4 + 4, not 5 + 5 + 5 + 1),
    and among those the split whose launches fill whole waves of CTA pairs best - 4 tensors of 128 row tiles are 256 pairs =
    3.46 waves (the fourth 46 % full), 8 of them 6.92; ties go to the smaller chunk (less memory alive)."""
    if most <= 0:
        return 0
    pairs_per_tensor = (row_tiles + 1) // 2
    best, best_eff = 0, -1.0
    for passes in range(-(-n_r // most), n_r + 1):
        chunk = -(-n_r // passes)
        sizes = [min(chunk, n_r - k) for k in range(0, n_r, chunk)]
        waves = sum(-(-(c * pairs_per_tensor) // _SM_PAIRS) for c in sizes)
        eff = n_r * pairs_per_tensor / (waves * _SM_PAIRS) - 0.002 * len(sizes)      # a pass costs a few launches and allocations
        if eff > best_eff + 1e-9:
            best, best_eff = chunk, eff
    return best


def _e_store_chunk(x_r: torch.Tensor, y_c: torch.Tensor, comm: Comm) -> int:
    """Row tensors per pass of the stored-exponential route (csrc/infonce_bwd_e2.cu), 0 = use the recompute kernels.

    The route needs dim 512 (a [128 x 512] fp32 dX accumulator is exactly the tensor memory of an SM) and gradient mixes
    that weigh d(scale) like dX (every mode but local_loss).  The forward keeps 2 bytes per logit; the gradients are formed
    right away, chunk of row tensors by chunk, so only one chunk of exponentials is alive at a time - two in a distributed
    run, where the forward of the next chunk is launched behind the all-gather of this chunk's column statistics."""
    n_r, b, dim = x_r.shape
    n_c, n_all, _ = y_c.shape
    if dim != 512 or _E_STORE_MAX_BYTES <= 0 or (comm.distributed and comm.local_loss):
        return 0
    per_tensor = n_c * (-(-b // 128) * 128) * (-(-n_all // 128) * 128) * 2
    if per_tensor * n_r < _E_STORE_MIN_BYTES:
        return 0
    in_flight = 2 if comm.distributed else 1
    key = (n_r, b, n_c, n_all, x_r.device, comm.world_size, id(comm.group))
    hit = _e_chunk_cache.get(key)
    if hit is not None:
        chunk, allocated_then = hit
        # Decided once per problem shape (cudaMemGetInfo stalls the launch queue, which a per-step call would pay every step)
        # and re-checked with the allocator's own host-side counter: if this process holds much more memory than when the
        # decision was taken (optimizer state, activations of a later phase), decide again instead of running out of memory.
        if not x_r.is_cuda or torch.cuda.memory_allocated(x_r.device) <= allocated_then + in_flight * chunk * per_tensor:
            return chunk
    budget = _E_STORE_MAX_BYTES // in_flight
    if x_r.is_cuda:
        # ~70 % of what is free for everything alive at once: the chunks in flight plus one more chunk's worth for the G tiles
        # of a group with column-side gradients (as large as its exponentials); the rest is headroom for the caller
        budget = min(budget, (_reusable_bytes(x_r.device) * 7 // 10) // (in_flight + 1))
    most = min(n_r, budget // per_tensor)
    chunk = _wave_aware_chunk(n_r, most, -(-b // 128))
    if comm.distributed:
        # every rank must make the same number of passes (each pass gathers its column statistics): agree on the smallest
        # chunk; ranks see the same shapes in the same order, so they all arrive here together.  (A re-decision on one rank
        # only would desynchronise the collectives: the re-check above uses the same counter on every rank of a symmetric
        # job; asymmetric jobs should pin COSMOS_B200_ESTORE_MAX_GB.)
        agreed = torch.tensor([chunk], dtype=torch.int32, device=x_r.device if x_r.is_cuda else "cpu")
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=comm.group)
        chunk = int(agreed.item())
    if len(_e_chunk_cache) > 64:
        _e_chunk_cache.clear()
    _e_chunk_cache[key] = (chunk, torch.cuda.memory_allocated(x_r.device) if x_r.is_cuda else 0)
    return chunk


# ------------------------------------------------------------------------------------------------
# stored-exponential route: the schedule shared by one group (ClipLoss) and two (COSMOSLoss)
# ------------------------------------------------------------------------------------------------

class _Group:
    """One cartesian product rows x cols with its logit scale: stacked inputs, outputs, and the unit gradients."""

    def __init__(self, rows, cols, scale, comm: Comm, prefetch, need_scale: bool, need_rows: bool, need_cols: bool):
        self.n_r, self.n_c = len(rows), len(cols)
        dev = rows[0].device
        dt = compute_dtype(rows[0].dtype)
        x_r = stack_views(rows, dt)                       # [n_r, b, D]
        pre_c = pre_r = None
        if prefetch is not None and comm.distributed:
            pre_c = prefetch.handles.get(tuple(id(t) for t in cols))
            pre_r = prefetch.handles.get(tuple(id(t) for t in rows))
        self.x_c = pre_c[0] if pre_c is not None else stack_views(cols, dt)   # [n_c, b, D]
        if pre_r is not None:
            x_r = pre_r[0]
        self.x_r = x_r
        self.pre_c, self.pre_r = pre_c, pre_r
        self.b = x_r.shape[1]
        if self.x_c.shape[1] != self.b:
            raise RuntimeError("cosmos_b200: both feature lists must have the same batch size")
        self.comm = comm
        self.N = comm.world_size * self.b
        self.P = self.n_r * self.n_c
        self.off = comm.rank * self.b if comm.distributed else 0
        self.scale_f = scale.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        self.need_scale, self.need_rows, self.need_cols = need_scale, need_rows, need_cols
        self._y_c = None

    @property
    def y_c(self) -> torch.Tensor:                        # [n_c, N, D]; waits for the prefetched all-gather on first use
        if self._y_c is None:
            self._y_c = self.pre_c[1].get() if self.pre_c is not None else gather_stack(self.x_c, self.comm)
        return self._y_c

    def eager_chunk(self) -> int:
        # (the column-side gradient of this route is a GEMM over stored G tiles, whose 16-byte pieces need N % 8 == 0)
        if not self.need_rows or (self.need_cols and self.N % 8 != 0):
            return 0
        # shapes only: the gathered column stack need not have arrived for this decision
        y_shape = torch.empty(self.n_c, self.N, self.x_r.shape[2], dtype=self.x_r.dtype, device="meta")
        return _e_store_chunk(self.x_r, y_shape, self.comm)


def _eager_schedule(groups: Sequence[_Group], comm: Comm):
    """Stored-exponential route for one or more groups: per chunk of row tensors, forward (statistics + 2^(s2 - max) of every
    logit, bf16) and, once the chunk's column statistics are complete, the gradients for a unit upstream gradient - dX = G Y
    is the only contraction left, no logit is recomputed.  backward() only scales by the upstream gradient.

    Distributed runs are software-pipelined so that no collective is waited for right after it was started: the forward of
    chunk k + 1 is launched behind the all-gather of chunk k's column statistics, the reduce-scatter of a group's column-side
    gradient runs behind the next group's kernels, and the loss sums and d(scale) of ALL groups share one small all-reduce
    at the end.  Returns the local loss totals; fills g.dx_unit, g.ds_unit, g.d_cols_unit, g.weight, g.pre per group."""
    W = comm.world_size
    items = []
    for g in groups:
        dev = g.x_r.device
        boost = float(W) if (comm.distributed and comm.gather_with_grad) else 1.0
        g.boost = boost
        g.weight = boost / (2.0 * g.N * g.P)
        # The unit gradients are held in the feature dtype.  Under GradScaler (fp16 features) the upstream factor exists
        # precisely because |dX| ~ weight * scale would underflow fp16, so they are formed with a power-of-two stand-in
        # for it (exact in every dtype; |dX * pre| <= 2 * n_c * scale / 64) and backward() multiplies by upstream / pre.
        g.pre = 2.0 ** (math.floor(math.log2(1.0 / g.weight)) - 6)
        g.one = torch.full((1,), g.pre, dtype=torch.float32, device=dev)
        g.row_lse2 = torch.empty(g.P, g.b, dtype=torch.float32, device=dev)
        g.diag_raw = torch.empty(g.P, g.b, dtype=torch.float32, device=dev)
        g.col_lse2 = torch.empty(g.P, g.N, dtype=torch.float32, device=dev)
        g.dx_unit = torch.empty_like(g.x_r)
        g.ds_unit = None
        g.d_all = None
        g.d_cols_unit = None
        g.rs_work = None
        g.rank_major = g.need_cols and not _COLS_VIA_GEMM and _rank_major_ok(comm)
        starts = list(range(0, g.n_r, g.chunk))
        for i0 in starts:
            items.append((g, i0, i0 == starts[-1]))

    def forward_of(item):
        g, i0, last = item
        xs = g.x_r[i0:i0 + g.chunk]
        sl = slice(i0 * g.n_c, (i0 + xs.shape[0]) * g.n_c)
        col_out = g.col_lse2[sl]
        part = torch.empty_like(col_out) if comm.distributed else col_out
        _r, _d, _c, e_, o_ = _k_fwd(xs, g.y_c, g.off, g.scale_f, True, out=(g.row_lse2[sl], g.diag_raw[sl], part))
        gather = _LseGather(part, comm) if comm.distributed else None
        return (g, i0, last, xs, sl, e_, o_, gather)

    def backward_of(state):
        g, i0, last, xs, sl, e_, o_, gather = state
        if gather is not None:
            gather.finish(g.col_lse2[sl])
        via_gemm = g.need_cols and _COLS_VIA_GEMM
        g_tiles = torch.empty(xs.shape[0] * g.b, g.n_c * g.N, dtype=g.x_r.dtype, device=xs.device) if via_gemm else None
        _dx, ds_ = _k_bwd_e(xs, g.y_c, g.off, g.scale_f, e_, o_, g.diag_raw[sl], g.row_lse2[sl], g.col_lse2[sl], 1.0, 1.0, 1.0 / g.boost,
                            1.0 / g.boost, g.weight, g.one, g.need_scale, g_tiles, dx_out=g.dx_unit[i0:i0 + xs.shape[0]])
        if g.need_cols:
            if via_gemm:      # diagnostics: G tiles through HBM + one GEMM (the first version of this route)
                part = _k_colgrad(g_tiles, xs.reshape(xs.shape[0] * g.b, xs.shape[2]), g.n_c, g.N)    # [n_c, N, D] fp32
            else:             # the same exponentials once more, read as G^T: nothing but dY goes to memory
                part = _k_bwd_e_cols(xs, g.y_c, g.off, g.scale_f, e_, o_, g.diag_raw[sl], g.row_lse2[sl], g.col_lse2[sl], 1.0, 1.0,
                                     rank_major=comm.world_size if g.rank_major else 0)
            g.d_all = part if g.d_all is None else g.d_all.add_(part)
        if ds_ is not None:
            g.ds_unit = ds_ if g.ds_unit is None else g.ds_unit + ds_
        if last and g.d_all is not None:
            # column-side gradient of this rank's rows -> sum over ranks of every rank's own rows; started now, needed by
            # backward(): it runs behind whatever is launched next
            g.d_cols_unit, g.rs_work = _reduce_scatter_rows(g.d_all, comm, g.b, async_op=True, rank_major=g.rank_major)
            g.d_all = None

    if comm.distributed:
        pending = None
        for item in items:
            state = forward_of(item)
            if pending is not None:
                backward_of(pending)
            pending = state
        backward_of(pending)
    else:
        for item in items:
            backward_of(forward_of(item))

    totals = []
    for g in groups:
        sums = _k_loss_sums(g.x_r, g.y_c, g.off, g.scale_f, g.row_lse2, g.diag_raw, g.col_lse2)   # [P, 2]
        totals.append(sums.sum().reshape(1))
    if comm.distributed:
        # one small all-reduce for everything that is a global scalar: the loss sums and the unit d(scale) of every group
        zero = torch.zeros(1, dtype=torch.float32, device=totals[0].device)
        vec = torch.cat(totals + [g.ds_unit.reshape(1) if g.ds_unit is not None else zero for g in groups])
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=comm.group)
        n = len(groups)
        totals = [vec[k:k + 1] for k in range(n)]
        for k, g in enumerate(groups):
            if g.ds_unit is not None:
                g.ds_unit = vec[n + k:n + k + 1]
        for g in groups:
            if g.rs_work is not None:
                g.rs_work.wait()
                g.rs_work = None
    return totals


def _eager_grads(g: _Group, up: torch.Tensor, in_dtypes, needs_rows, needs_cols, scale_dtype):
    """The gradients of one group for the upstream gradient `up` (fp32 [1]) from its unit gradients."""
    factor = up / g.pre
    grads: List[Optional[torch.Tensor]] = []
    # one vectorised pass, products formed in fp32 and rounded once to the stack dtype
    d_rows = _k_scale16(g.dx_unit, up.contiguous(), g.pre) if any(needs_rows) else None
    for k in range(g.n_r):
        grads.append(d_rows[k].to(in_dtypes[k]) if (d_rows is not None and needs_rows[k]) else None)
    d_cols = None
    if g.d_cols_unit is not None and any(needs_cols):
        d_cols = (g.d_cols_unit * (up * g.scale_f.reshape(1) * g.weight)).to(g.dx_unit.dtype)
    for k in range(g.n_c):
        grads.append(d_cols[k].to(in_dtypes[g.n_r + k]) if (d_cols is not None and needs_cols[k]) else None)
    g_scale = (g.ds_unit.reshape(()) * factor.reshape(())).to(scale_dtype) if (g.need_scale and g.ds_unit is not None) else None
    return g_scale, grads


class _PairsInfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scale: torch.Tensor, comm: Comm, n_r: int, prefetch, want_grad: bool, *feats: torch.Tensor):
        rows, cols = feats[:n_r], feats[n_r:]
        n_c = len(cols)
        # needs_input_grad stays True under torch.no_grad(): a validation loss must not pay for gradients nobody will ask for
        need_scale = want_grad and ctx.needs_input_grad[0]
        need_rows = want_grad and any(ctx.needs_input_grad[5:5 + n_r])
        need_cols = want_grad and any(ctx.needs_input_grad[5 + n_r:])
        g = _Group(rows, cols, scale, comm, prefetch, need_scale, need_rows, need_cols)
        b, N, P, off = g.b, g.N, g.P, g.off
        g.chunk = g.eager_chunk()
        ctx.eager = g.chunk > 0
        if ctx.eager:
            (total,) = _eager_schedule([g], comm)         # local_loss never takes this route: the total is global already
            total = total.reshape(())
            loss = total / (2.0 * N * P)
            ctx.group = g
        else:
            row_lse2, diag_raw, col_lse2_part = _k_fwd(g.x_r, g.y_c, off, g.scale_f)
            col_lse2 = _allreduce_lse2(col_lse2_part, comm)   # global over all rows
            sums = _k_loss_sums(g.x_r, g.y_c, off, g.scale_f, row_lse2, diag_raw, col_lse2)   # [P, 2]
            total = sums.sum()
            if comm.distributed and not comm.local_loss:
                dist.all_reduce(total, op=dist.ReduceOp.SUM, group=comm.group)
                loss = total / (2.0 * N * P)
            else:
                loss = total / (2.0 * b * P)

        ctx.comm, ctx.n_r, ctx.n_c, ctx.b, ctx.off = comm, n_r, n_c, b, off
        ctx.pre_r = g.pre_r[1] if g.pre_r is not None else None
        ctx.in_dtypes = [t.dtype for t in feats]
        ctx.scale_dtype = scale.dtype
        ctx.want_grad = want_grad
        if not ctx.eager and want_grad:
            ctx.save_for_backward(g.x_r, g.x_c, g.y_c, g.scale_f, row_lse2, col_lse2)
        return loss

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        if not ctx.want_grad:
            raise RuntimeError("cosmos_b200: this loss was computed under torch.no_grad(); it has no gradients")
        if ctx.eager:
            up = g.detach().to(torch.float32).reshape(1)
            n_r = ctx.n_r
            g_scale, grads = _eager_grads(ctx.group, up, ctx.in_dtypes, ctx.needs_input_grad[5:5 + n_r],
                                          ctx.needs_input_grad[5 + n_r:], ctx.scale_dtype)
            if not ctx.needs_input_grad[0]:
                g_scale = None
            return (g_scale, None, None, None, None, *grads)
        x_r, x_c, y_c, scale_f, row_lse2, col_lse2 = ctx.saved_tensors
        comm, n_r, n_c, b, off = ctx.comm, ctx.n_r, ctx.n_c, ctx.b, ctx.off
        W = comm.world_size
        N, P = W * b, n_r * n_c
        need_scale = ctx.needs_input_grad[0]
        need_rows = any(ctx.needs_input_grad[5:5 + n_r])
        need_cols = any(ctx.needs_input_grad[5 + n_r:])
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
            grads.append(d_rows[k].to(ctx.in_dtypes[k]) if (d_rows is not None and ctx.needs_input_grad[5 + k]) else None)
        for k in range(n_c):
            grads.append(d_cols[k].to(ctx.in_dtypes[n_r + k])
                         if (d_cols is not None and ctx.needs_input_grad[5 + n_r + k]) else None)
        g_scale = d_scale.reshape(()).to(ctx.scale_dtype) if need_scale else None
        return (g_scale, None, None, None, None, *grads)


class _CosmosHead(torch.autograd.Function):
    """Both InfoNCE groups of the COSMOS loss (src/open_clip/loss.py:193-206) on the stored-exponential route in ONE
    schedule: the CLIP group (student captions x the two global student crops) first, then the distillation group (all
    cross-modal student features x the four teacher features), so that in a distributed run the reduce-scatter of the CLIP
    group's image-side gradient and every all-gather of column statistics run behind kernels of the other group.
    Returns (distill_loss, clip_loss)."""

    @staticmethod
    def forward(ctx, logit_scale, distill_scale, comm: Comm, counts, groups, *feats):
        clip, dist_g = groups
        t_clip, t_dist = _eager_schedule([clip, dist_g], comm)
        ctx.groups = (clip, dist_g)
        ctx.counts = counts
        ctx.in_dtypes = [t.dtype for t in feats]
        ctx.scale_dtypes = (logit_scale.dtype, distill_scale.dtype)
        return (t_dist.reshape(()) / (2.0 * dist_g.N * dist_g.P), t_clip.reshape(()) / (2.0 * clip.N * clip.P))

    @staticmethod
    def backward(ctx, g_distill, g_clip):
        clip, dist_g = ctx.groups
        n_txt, n_img, n_x, n_t = ctx.counts
        need = ctx.needs_input_grad
        o = 5
        dt = ctx.in_dtypes
        up_c = g_clip.detach().to(torch.float32).reshape(1)
        up_d = g_distill.detach().to(torch.float32).reshape(1)
        gs_c, grads_c = _eager_grads(clip, up_c, dt[:n_txt + n_img], need[o:o + n_txt], need[o + n_txt:o + n_txt + n_img],
                                     ctx.scale_dtypes[0])
        k0 = n_txt + n_img
        gs_d, grads_d = _eager_grads(dist_g, up_d, dt[k0:], need[o + k0:o + k0 + n_x], need[o + k0 + n_x:], ctx.scale_dtypes[1])
        return (gs_c if need[0] else None, gs_d if need[1] else None, None, None, None, *grads_c, *grads_d)


def cosmos_head(texts, images2, xrows, teacher, logit_scale, distill_scale, comm: Comm, prefetch=None):
    """(distill_loss, clip_loss) of the COSMOS loss when BOTH groups can take the stored-exponential route (dim 512, not a
    local-loss mode, large enough, gradients wanted); None otherwise - the caller then evaluates the groups one by one."""
    if not torch.is_grad_enabled() or (comm.distributed and comm.local_loss):
        return None
    texts, images2, xrows, teacher = list(texts), list(images2), list(xrows), list(teacher)
    for t in texts + images2 + xrows + teacher:
        _lib.require_cuda(t, "feature tensor")
    dev = texts[0].device
    scales = []
    for sc in (logit_scale, distill_scale):
        scales.append(sc if isinstance(sc, torch.Tensor) else torch.tensor(float(sc), dtype=torch.float32, device=dev))
    rg = lambda ts: any(t.requires_grad for t in ts)
    clip = _Group(texts, images2, scales[0], comm, prefetch, scales[0].requires_grad, rg(texts), rg(images2))
    dist_g = _Group(xrows, teacher, scales[1], comm, prefetch, scales[1].requires_grad, rg(xrows), False)
    clip.chunk, dist_g.chunk = clip.eager_chunk(), dist_g.eager_chunk()
    if clip.chunk <= 0 or dist_g.chunk <= 0:
        return None
    counts = (len(texts), len(images2), len(xrows), len(teacher))
    return _CosmosHead.apply(scales[0], scales[1], comm, counts, (clip, dist_g), *texts, *images2, *xrows, *teacher)


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
    return _PairsInfoNCE.apply(scale, comm, len(rows), prefetch, torch.is_grad_enabled(), *rows, *cols)

"""EMA teacher update, one kernel launch for the whole parameter set.

Drop-in for the inline loop of the reference's train_one_epoch
(src/training/train.py:195-203):

    with torch.no_grad():
        for param_q, param_k in zip(student.parameters(), teacher.parameters()):
            param_k.data.mul_(momentum).add_((1 - momentum) * param_q.detach().data)

`ema_update_(student, teacher, momentum)` has the same semantics (same rounding: fp32 results
are bit-identical), updates the SAME teacher storage in place (checkpointing keeps working,
SURVEY.md §5) and runs after backward / before optimizer.step exactly where the loop sat.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, List, Sequence, Union

import torch

from . import _lib


def _params(x) -> List[torch.Tensor]:
    if isinstance(x, torch.nn.Module):
        return list(x.parameters())
    return list(x)


class EmaPlan:
    """Chunk table for one (student, teacher) parameter set; rebuilt only when a pointer moves."""

    def __init__(self, student: Sequence[torch.Tensor], teacher: Sequence[torch.Tensor]):
        if len(student) != len(teacher):
            raise RuntimeError("cosmos_b200.ema: student and teacher have different numbers of parameters")
        self.groups = []   # one (dtype_code, device_index, table_tensor, n_entries) per (dtype, device)
        buckets = {}
        for q, k in zip(student, teacher):
            _lib.require_cuda(k, "teacher parameter")
            _lib.require_cuda(q, "student parameter")
            if q.shape != k.shape or q.dtype != k.dtype or q.device != k.device:
                raise RuntimeError("cosmos_b200.ema: student/teacher parameter mismatch "
                                   f"({tuple(q.shape)} {q.dtype} {q.device} vs {tuple(k.shape)} {k.dtype} {k.device})")
            if not (q.is_contiguous() and k.is_contiguous()):
                raise RuntimeError("cosmos_b200.ema: parameters must be contiguous")
            if k.numel() == 0:
                continue
            buckets.setdefault((k.dtype, k.device.index), []).append((k, q))
        self.key = tuple((k.data_ptr(), q.data_ptr(), k.numel()) for q, k in zip(student, teacher))
        self._keep = (list(student), list(teacher))   # keeps the ids used as cache key alive
        lib = _lib.lib()
        for (dtype, dev), pairs in buckets.items():
            n = len(pairs)
            numel = (C.c_int64 * n)(*[k.numel() for k, _ in pairs])
            kp = (C.c_uint64 * n)(*[k.data_ptr() for k, _ in pairs])
            qp = (C.c_uint64 * n)(*[q.data_ptr() for _, q in pairs])
            entries = lib.cosmos_ema_table_entries(n, numel)
            if entries < 0:
                raise RuntimeError("cosmos_b200.ema: bad parameter sizes")
            host = torch.empty(entries * C.sizeof(_lib.EmaChunk), dtype=torch.uint8, pin_memory=True)
            # the element size of THIS bucket (mixed-precision models keep fp32 norms / embeddings next to 16-bit weights)
            elem_size = torch.empty((), dtype=dtype).element_size()
            _lib.check(lib.cosmos_ema_table_fill(n, kp, qp, numel, elem_size, host.data_ptr()), "ema_table_fill")
            table = host.to(torch.device("cuda", dev), non_blocking=False)
            self.groups.append((_lib.torch_dtype_code(dtype), dev, table, entries))

    def apply(self, momentum: float) -> None:
        lib = _lib.lib()
        for code, dev, table, entries in self.groups:
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.cosmos_ema_apply(table.data_ptr(), entries, float(momentum), code, dev, stream), "ema_apply")


_plans = {}


def _full_key(sp, tp):
    return tuple((k.data_ptr(), q.data_ptr(), k.numel()) for q, k in zip(sp, tp))


@torch.no_grad()
def ema_update_(student: Union[torch.nn.Module, Iterable[torch.Tensor]],
                teacher: Union[torch.nn.Module, Iterable[torch.Tensor]], momentum: float) -> None:
    """teacher <- teacher * momentum + (1 - momentum) * student, in place, on the current stream.

    The chunk table is cached per parameter set (keyed by the identity of the parameter objects) and holds raw
    device pointers, so every (pointer, pointer, numel) triple is re-checked on every call - a few hundred Python
    ints, against the 969 launches this replaces: a parameter whose storage moved (`p.data = ...`, `.to()`,
    `load_state_dict(assign=True)`, offload) rebuilds the table instead of updating freed memory.  Hold an
    `EmaPlan` yourself (and call `plan.apply(m)`) to skip even that bookkeeping; then pointer stability is yours."""
    sp, tp = _params(student), _params(teacher)
    ident = (tuple(map(id, tp)), tuple(map(id, sp)))
    plan = _plans.get(ident)
    if plan is not None and plan.key != _full_key(sp, tp):
        plan = None
    if plan is None:
        if len(_plans) > 16:
            _plans.clear()
        plan = _plans[ident] = EmaPlan(sp, tp)
    plan.apply(momentum)


def _unwrap(model):
    """DistributedDataParallel / DataParallel wrappers keep the model in `.module` (train.py: unwrap_model)."""
    return model.module if hasattr(model, "module") else model


@torch.no_grad()
def clamp_logit_scales_(student, teacher=None, lo: float = 0.0, hi: float = math.log(100)) -> None:
    """Drop-in for the clamps at the end of the reference's training step (src/training/train.py:237-243):

        unwrap_model(student).logit_scale.clamp_(0, math.log(100))
        unwrap_model(teacher).logit_scale.clamp_(0, math.log(100))
        # and, when the model has one, the same for distill_logit_scale

    One launch for all (up to four) scalars instead of four; same results bit for bit (torch.clamp_ semantics).
    `student` / `teacher` may be modules (wrapped or not) or iterables of 1-element tensors."""
    scalars: List[torch.Tensor] = []
    for m in (student, teacher):
        if m is None:
            continue
        if isinstance(m, torch.nn.Module):
            m = _unwrap(m)
            for name in ("logit_scale", "distill_logit_scale"):
                t = getattr(m, name, None)
                if isinstance(t, torch.Tensor):
                    scalars.append(t)
        else:
            scalars.extend(m)
    groups = {}
    for t in scalars:
        _lib.require_cuda(t, "logit scale")
        if t.numel() != 1:
            raise RuntimeError(f"cosmos_b200.ema: clamp_logit_scales_ expects 1-element tensors, got {tuple(t.shape)}")
        groups.setdefault((t.dtype, t.device.index), []).append(t)
    lib = _lib.lib()
    for (dtype, dev), ts in groups.items():
        code = _lib.torch_dtype_code(dtype)
        stream = torch.cuda.current_stream(dev).cuda_stream
        for first in range(0, len(ts), _lib.CLAMP_MAX):
            part = ts[first:first + _lib.CLAMP_MAX]
            ptrs = (C.c_uint64 * len(part))(*[t.data_ptr() for t in part])
            _lib.check(lib.cosmos_clamp_scalars(ptrs, len(part), float(lo), float(hi), code, dev, stream), "clamp_scalars")

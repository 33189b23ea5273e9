"""Drop-in replacement for the loss API of the reference's `open_clip.loss`
(src/open_clip/loss.py:21-207): `gather_features`, `ClipLoss`, `COSMOSLoss` keep their constructor
and forward signatures, argument meaning, return conventions (0-dim tensors, the
{"distill_loss", "clip_loss"} dict) and error behaviour, but the pairwise
logits + cross-entropy work runs in the sm_100a kernels of libcosmos_b200.so
(cosmos_b200/infonce.py) instead of an N x N matmul + F.cross_entropy per pair.

`create_loss(args)` (src/open_clip/factory.py:372-415) constructs `COSMOSLoss(local_loss=...,
gather_with_grad=..., cache_labels=True, rank=..., world_size=..., use_horovod=...)`; the same call
works here.  Horovod is rejected (north_star: no multi-backend dispatch).  CoCaLoss /
DistillClipLoss / SigLipLoss are not reachable with --cosmos (factory.py:372-407) and are out of
scope: the names exist so `from open_clip.loss import ...` keeps importing, and raise on use.
"""
from __future__ import annotations

import torch
import torch.nn as nn

try:
    import torch.distributed.nn
    from torch import distributed as dist

    has_distributed = True
except ImportError:  # pragma: no cover
    has_distributed = False

from .infonce import Comm, Prefetch, _reduce_scatter_rows, cosmos_head, gather_stack, pairs_infonce

__all__ = ["gather_features", "ClipLoss", "COSMOSLoss", "CoCaLoss", "DistillClipLoss", "SigLipLoss"]


def _as_list(features):
    return list(features) if isinstance(features, (list, tuple)) else [features]


class _GatherRows(torch.autograd.Function):
    """[k, b, D] on every rank -> [k, W * b, D] (rank-major rows) with the gradient an all-gather has: the rows of the incoming
    gradient that belong to this rank, summed over ranks (reduce-scatter)."""

    @staticmethod
    def forward(ctx, local: torch.Tensor, comm: Comm):
        ctx.comm, ctx.b = comm, local.shape[1]
        return gather_stack(local.detach(), comm)

    @staticmethod
    def backward(ctx, g):
        return _reduce_scatter_rows(g.contiguous(), ctx.comm, ctx.b), None


def gather_features(image_features, text_features, local_loss=False, gather_with_grad=False, rank=0, world_size=1,
                    use_horovod=False):
    """Both feature matrices of every rank, rows in rank-major order (the reference's gather_features,
    src/open_clip/loss.py:21-65): ONE all_gather_into_tensor for the two of them instead of two list all-gathers.

      gather_with_grad                  the gathered rows carry the all-gather's gradient (reduce-scatter-sum) to every rank
      not gather_with_grad, local_loss  remote and local rows are constants
      neither                           remote rows are constants, this rank's own rows are its inputs (they keep their graph)

    Pure communication on the default process group the training script created (src/training/distributed.py:89-102);
    `ClipLoss.forward` does not use it - the kernels gather whole feature lists at once (cosmos_b200/infonce.py)."""
    assert has_distributed, 'torch.distributed did not import correctly, please use a PyTorch version with support.'
    if use_horovod:
        raise RuntimeError("cosmos_b200: Horovod is not supported (NCCL via torch.distributed only)")
    comm = Comm(rank=rank, world_size=world_size, local_loss=bool(local_loss), gather_with_grad=bool(gather_with_grad))
    if image_features.shape == text_features.shape and image_features.dtype == text_features.dtype:
        stacks = [torch.stack([image_features, text_features])]
    else:
        stacks = [image_features[None], text_features[None]]
    out = []
    for local in stacks:
        b = local.shape[1]
        if gather_with_grad:
            both = _GatherRows.apply(local, comm)
        else:
            with torch.no_grad():
                both = gather_stack(local, comm)
            if not local_loss:
                lo = rank * b
                both = torch.cat([both[:, :lo], local, both[:, lo + b:]], dim=1)
        out.extend(both.unbind(0))
    return out[0], out[1]


class ClipLoss(nn.Module):
    """Symmetric InfoNCE, averaged over every (image tensor, text tensor) pair (loss.py:68-142)."""

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1,
                 use_horovod=False):
        super().__init__()
        if use_horovod:
            raise RuntimeError("cosmos_b200: Horovod is not supported (NCCL via torch.distributed only)")
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod

        # cache state (kept for attribute compatibility; the kernels derive labels from rank * batch)
        self.prev_num_logits = 0
        self.labels = {}

    def _comm(self) -> Comm:
        return Comm(rank=self.rank, world_size=self.world_size, local_loss=bool(self.local_loss),
                    gather_with_grad=bool(self.gather_with_grad))

    def get_ground_truth(self, device, num_logits) -> torch.Tensor:
        """Column of the positive for every local row (loss.py:90-101): row r pairs with column r, shifted by this rank's
        first row in the sharded local-loss modes; cached per device when `cache_labels`."""
        cached = self.labels.get(device) if self.prev_num_logits == num_logits else None
        if cached is not None:
            return cached
        first = num_logits * self.rank if (self.world_size > 1 and self.local_loss) else 0
        labels = torch.arange(first, first + num_logits, device=device, dtype=torch.long)
        if self.cache_labels:
            self.labels[device] = labels
            self.prev_num_logits = num_logits
        return labels

    def get_logits(self, image_features, text_features, logit_scale):
        """The materialised logit matrices (loss.py:103-119), for callers outside the COSMOS path that want them; `forward`
        never builds them - the kernels keep the logits in tensor memory.  Sharded runs: local rows against everybody's
        columns with local_loss, the full N x N matrix (and its transpose) otherwise."""
        scaled_image, scaled_text = logit_scale * image_features, logit_scale * text_features
        if self.world_size <= 1:
            return scaled_image @ text_features.T, scaled_text @ image_features.T
        all_image, all_text = gather_features(image_features, text_features, self.local_loss, self.gather_with_grad,
                                              self.rank, self.world_size, self.use_horovod)
        if self.local_loss:
            return scaled_image @ all_text.T, scaled_text @ all_image.T
        per_image = (logit_scale * all_image) @ all_text.T
        return per_image, per_image.T

    def forward(self, image_features, text_features, logit_scale, output_dict=False):
        image_features, text_features = _as_list(image_features), _as_list(text_features)   # one tensor or one per crop / caption
        # The loss is symmetric in the two lists; the shorter one becomes the all-gathered column side.
        if len(text_features) >= len(image_features):
            total_loss = pairs_infonce(text_features, image_features, logit_scale, self._comm())
        else:
            total_loss = pairs_infonce(image_features, text_features, logit_scale, self._comm())
        return {"contrastive_loss": total_loss} if output_dict else total_loss


class COSMOSLoss(nn.Module):
    """COSMOS loss head (loss.py:145-207): cross-modality self-distillation (student cross-modal
    features against detached EMA-teacher features, 4 x 16 InfoNCE pairs) + CLIP loss between the two
    global student image crops and all student captions."""

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1,
                 use_horovod=False):
        super().__init__()
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod

        # cache state
        self.prev_num_logits = 0
        self.labels = {}

        self.clip_loss = ClipLoss(local_loss=self.local_loss, gather_with_grad=self.gather_with_grad,
                                  cache_labels=self.cache_labels, rank=self.rank, world_size=self.world_size,
                                  use_horovod=self.use_horovod)

    def forward(self, s_image_features, s_text_features, logit_scale, t_image_features=None, t_text_features=None,
                output_dict=False, distill_logit_scale=None, s_img_crossmodal_features=None,
                s_txt_crossmodal_features=None):
        s_image_features, s_text_features = _as_list(s_image_features), _as_list(s_text_features)
        if t_image_features is None or t_text_features is None:
            raise RuntimeError("COSMOSLoss needs teacher image and text features")
        if s_img_crossmodal_features is None or s_txt_crossmodal_features is None:
            raise RuntimeError("COSMOSLoss needs the student cross-modal features")
        assert len(t_image_features) == 2
        assert len(t_text_features) == 2
        # no gradient flows to teacher
        teacher = [f.detach() for f in t_image_features] + [f.detach() for f in t_text_features]

        comm = self.clip_loss._comm()
        scale = distill_logit_scale if distill_logit_scale is not None else logit_scale
        # Start every all-gather now (NCCL stream), in the order of their first use: the student image stack (CLIP group,
        # evaluated first), the teacher stack (distillation group), the student text stack (only the local-loss modes sweep
        # the transposed block and need it) - all but the first hide behind kernels.
        images2, texts = list(s_image_features[:2]), list(s_text_features)
        prefetch = Prefetch()
        prefetch.start(images2, comm)
        prefetch.start(teacher, comm)
        if any(t.requires_grad for t in images2) and comm.local_loss:
            prefetch.start(texts, comm)
        # mean over {img-x, txt-x} x {t_img, t_txt} of ClipLoss(n x 2 pairs)  ==  mean of the two n x 4 groups; with equally
        # long lists (the COSMOS recipes: 8 + 8) that is the plain mean over all 16 x 4 pairs -> ONE grouped launch, which
        # fills whole waves of SM clusters where two half-sized launches would each leave a partial wave.
        img_x, txt_x = list(s_img_crossmodal_features), list(s_txt_crossmodal_features)
        if len(img_x) == len(txt_x):
            # the headline path: both groups in one schedule on the stored-exponential route (cosmos_b200/infonce.py)
            both = cosmos_head(texts, images2, img_x + txt_x, teacher, logit_scale, scale, comm, prefetch)
            if both is not None:
                cosmos_loss, clip_loss = both
                return {"distill_loss": cosmos_loss, "clip_loss": clip_loss} if output_dict else cosmos_loss + clip_loss
            cosmos_loss = pairs_infonce(img_x + txt_x, teacher, scale, comm, prefetch)
        else:
            cosmos_loss = (pairs_infonce(img_x, teacher, scale, comm, prefetch)
                           + pairs_infonce(txt_x, teacher, scale, comm, prefetch)) / 2

        # CLIP loss: only the two global crops on the image side (loss.py:205-206)
        clip_loss = pairs_infonce(texts, images2, logit_scale, comm, prefetch)
        return {"distill_loss": cosmos_loss, "clip_loss": clip_loss} if output_dict else cosmos_loss + clip_loss


class _OutOfScope(nn.Module):
    _what = ""

    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError(
            f"cosmos_b200: {self._what} is not part of the COSMOS loss-head path (unreachable with --cosmos, "
            "src/open_clip/factory.py:372-407); use the reference implementation for it")


class CoCaLoss(_OutOfScope):
    _what = "CoCaLoss"


class DistillClipLoss(_OutOfScope):
    _what = "DistillClipLoss"


class SigLipLoss(_OutOfScope):
    _what = "SigLipLoss"

"""Retrieval metrics of the evaluation without a similarity matrix, an argsort or a Python loop over rows.

Replaces, outside the training step:

  * `get_clip_metrics(image_features, text_features, logit_scale)`      src/training/train.py:766-785
  * `compute_retrieval(similarity_scores, txt2img, img2txt)`            src/training/train.py:712-763
    together with the similarity matrix its callers build on the CPU     src/training/train.py:665-684

The position of the ground-truth item in a descending sort of a row is the number of items with a larger score,
so `retrieval_ranks` counts instead of sorting (csrc/retrieval.cu: fp32 dot products on the CUDA cores, nothing of
size queries x gallery is ever stored).  Only the integer ranks come back to the host; the metric dictionaries are then
formed with the reference's own expressions and key names.  A positive `logit_scale` does not change any rank and is
not applied.  Ties and NaN are ordered as torch.sort orders them when it is stable (NaN above every number, equal
scores by index): a collapsed or diverged model gets chance-level ranks like in the reference, never "rank 0 for every
query".  Ranks can differ from the reference's only where two scores tie to within fp32 summation-order noise (the
reference's own unstable argsort is arbitrary there too).
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Sequence

import numpy as np
import torch

from . import _lib


def retrieval_ranks(queries: torch.Tensor, gallery: torch.Tensor, gt_offsets: Optional[torch.Tensor] = None,
                    gt_index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int32 [M]: for every query row, the 0-based position of its best ground-truth row in a stable descending sort
    of its scores (rows scoring higher + tied rows with a lower index; NaN sorts above every number).

    Ground truth in CSR form (int32 CUDA tensors): rows gt_index[gt_offsets[r]:gt_offsets[r+1]] of the gallery; both
    None = row r; gt_index None = the contiguous range gt_offsets[r]:gt_offsets[r+1]."""
    _lib.require_cuda(queries, "queries")
    _lib.require_cuda(gallery, "gallery")
    if queries.dim() != 2 or gallery.dim() != 2 or queries.shape[1] != gallery.shape[1]:
        raise RuntimeError(f"cosmos_b200.retrieval: expected [M, D] and [N, D], got {tuple(queries.shape)} and {tuple(gallery.shape)}")
    if queries.dtype != gallery.dtype or queries.device != gallery.device:
        raise RuntimeError("cosmos_b200.retrieval: queries and gallery must share dtype and device")
    M, D = queries.shape
    N = gallery.shape[0]
    if M == 0:
        return torch.empty(0, dtype=torch.int32, device=queries.device)
    if N == 0:
        raise RuntimeError("cosmos_b200.retrieval: empty gallery")
    if queries.stride(1) != 1:
        queries = queries.contiguous()
    if gallery.stride(1) != 1:
        gallery = gallery.contiguous()
    if gt_offsets is None:
        if gt_index is not None:
            raise RuntimeError("cosmos_b200.retrieval: gt_index needs gt_offsets")
        if N < M:
            raise RuntimeError("cosmos_b200.retrieval: paired metrics need one gallery row per query")
    else:
        for name, t in (("gt_offsets", gt_offsets), ("gt_index", gt_index)):
            if t is None:
                continue
            _lib.require_cuda(t, name)
            if t.dtype != torch.int32 or not t.is_contiguous() or t.device != queries.device:
                raise RuntimeError(f"cosmos_b200.retrieval: {name} must be a contiguous int32 tensor on the queries' device")
        if gt_offsets.numel() != M + 1:
            raise RuntimeError("cosmos_b200.retrieval: gt_offsets needs M + 1 entries")
    dev = queries.device
    best = torch.empty(M, dtype=torch.float32, device=dev)
    best_col = torch.empty(M, dtype=torch.int32, device=dev)
    ranks = torch.empty(M, dtype=torch.int32, device=dev)
    st = _lib.lib().cosmos_retrieval_ranks(
        queries.data_ptr(), gallery.data_ptr(), _lib.torch_dtype_code(queries.dtype), M, N, D, queries.stride(0), gallery.stride(0),
        gt_offsets.data_ptr() if gt_offsets is not None else None, gt_index.data_ptr() if gt_index is not None else None,
        best.data_ptr(), best_col.data_ptr(), ranks.data_ptr(), dev.index, torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(st, "retrieval_ranks")
    return ranks


def _csr(lists: Sequence[Sequence[int]], n_gallery: int, device) -> tuple:
    offsets, index = [0], []
    for items in lists:
        for t in items:
            t = int(t)
            if not 0 <= t < n_gallery:
                raise RuntimeError(f"cosmos_b200.retrieval: ground-truth index {t} outside the gallery (size {n_gallery})")
            index.append(t)
        offsets.append(len(index))
    return (torch.tensor(offsets, dtype=torch.int32, device=device), torch.tensor(index, dtype=torch.int32, device=device))


def get_clip_metrics(image_features: torch.Tensor, text_features: torch.Tensor, logit_scale=None) -> Dict[str, float]:
    """Same keys and values as the reference's get_clip_metrics (train.py:766-785) for paired features."""
    del logit_scale     # order-preserving (the reference multiplies by it before sorting)
    metrics = {}
    sides = {"image_to_text": (image_features, text_features), "text_to_image": (text_features, image_features)}
    for name, (q, g) in sides.items():
        preds = retrieval_ranks(q.detach(), g.detach()).cpu().numpy().astype(np.int64)
        metrics[f"{name}_mean_rank"] = preds.mean() + 1
        metrics[f"{name}_median_rank"] = np.floor(np.median(preds)) + 1
        for k in [1, 5, 10]:
            metrics[f"{name}_R@{k}"] = np.mean(preds < k)
    return metrics


def compute_retrieval(image_features: torch.Tensor, text_features: torch.Tensor, txt2img: Mapping[int, int],
                      img2txt: Mapping[int, Sequence[int]]) -> Dict[str, float]:
    """The reference's compute_retrieval (train.py:712-763) from the FEATURES instead of a CPU similarity matrix
    (`compute_similarity_scores_original_clip`, train.py:665-684): image i matches captions img2txt[i] (the best one
    counts), caption c matches image txt2img[c]."""
    dev = image_features.device
    n_img, n_txt = image_features.shape[0], text_features.shape[0]
    i_off, i_idx = _csr([img2txt[i] for i in range(n_img)], n_txt, dev)
    t_off, t_idx = _csr([[txt2img[c]] for c in range(n_txt)], n_img, dev)
    i2t = retrieval_ranks(image_features.detach(), text_features.detach(), i_off, i_idx).cpu().to(torch.float32)
    t2i = retrieval_ranks(text_features.detach(), image_features.detach(), t_off, t_idx).cpu().to(torch.float32)

    def report(ranks: torch.Tensor, name: str) -> Dict[str, float]:      # the reference holds ranks in a float32 tensor
        n = len(ranks)
        return {f"{name}_R@1": len(torch.where(ranks < 1)[0]) / n, f"{name}_R@5": len(torch.where(ranks < 5)[0]) / n,
                f"{name}_R@10": len(torch.where(ranks < 10)[0]) / n, f"{name}_mean_rank": ranks.mean().item() + 1,
                f"{name}_median_rank": np.floor(np.median(ranks.numpy())) + 1}

    return {**report(t2i, "text_to_image"), **report(i2t, "image_to_text")}

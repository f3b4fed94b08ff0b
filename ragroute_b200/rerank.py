"""Boundary B3: cross-source merge.

Drop-in signatures for /root/reference/ragroute/rerank.py: ``rerank_medrag(docs,
scores, k)`` (:3-9, called at http_server.py:289) and ``rerank_wikipedia`` (:28-34,
ascending as written, called at :293), plus the batched tensor form ``merge_topk``
used by the federation driver and by the per-shard merge after the all-gather.
The selection runs in the warp-level k-way select kernel of libragroute_b200.so.

Tie order: the reference's ``np.argsort`` is not stable, so the order of equal
scores is undefined there; here equal scores keep their input order.
``rerank_feb4rag`` (:12-25) orders by ground-truth qrels (string keys, no
arithmetic) and is out of scope.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def merge_topk(scores: torch.Tensor, ids: torch.Tensor, k: int, descending: bool = True,
               want_pos: bool = True) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """scores fp32 CUDA [B, L], ids int64 CUDA [B, L] (id < 0 = empty slot) ->
    (scores [B, k], ids [B, k], pos int32 [B, k]); unused slots (-/+FLT_MAX, -1, -1)."""
    _lib.require_cuda()
    assert scores.is_cuda and ids.is_cuda and scores.dtype == torch.float32 and ids.dtype == torch.int64
    assert scores.shape == ids.shape and scores.dim() == 2 and scores.is_contiguous() and ids.is_contiguous()
    B, L = scores.shape
    with torch.cuda.device(scores.device):
        os_ = torch.empty((B, k), dtype=torch.float32, device=scores.device)
        oi = torch.empty((B, k), dtype=torch.int64, device=scores.device)
        op = torch.empty((B, k), dtype=torch.int32, device=scores.device) if want_pos else None
        _lib.check(_lib.load().rr_merge_topk(scores.data_ptr(), ids.data_ptr(), B, L, int(k), 1 if descending else 0,
                                             os_.data_ptr(), oi.data_ptr(), op.data_ptr() if want_pos else None,
                                             _stream_ptr()))
    return os_, oi, op


def _rerank(docs: Sequence, scores: Sequence[float], k: int, descending: bool) -> Tuple[List, List]:
    n = len(scores)
    if n == 0:
        return [], []
    s = torch.tensor(np.asarray(scores, dtype=np.float32)).reshape(1, n).cuda()
    ids = torch.arange(n, dtype=torch.int64, device=s.device).reshape(1, n)
    _, _, pos = merge_topk(s, ids, min(k, n), descending=descending)
    order = pos[0].cpu().tolist()
    return [docs[i] for i in order], [scores[i] for i in order]


def rerank_medrag(docs, scores, k):
    """Global top-k of the concatenated per-source lists, best (largest) first."""
    return _rerank(docs, scores, k, descending=True)


def rerank_wikipedia(docs, scores, k):
    """As the reference writes it: ascending (smallest score first)."""
    return _rerank(docs, scores, k, descending=False)

"""oracle/rerank_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's cross-source merge:

* ``rerank_medrag``    <- /root/reference/ragroute/rerank.py:3-9
  (``np.argsort(scores)[::-1]``: global top-k, descending), called at
  /root/reference/ragroute/http_server.py:289.
* ``rerank_wikipedia`` <- rerank.py:28-34 (``np.argsort(scores)[::]``: ASCENDING,
  behaviour as written), called at http_server.py:293.
* ``merge_topk``       -- the same selection on [B, L] arrays with ids carried
  along; tie order made explicit (score, then lower input position first; for
  lists with source structure: score, source, id, position), which the reference
  leaves undefined (default ``argsort`` is not stable).

Pinned against the reference itself: tests/golden/rerank.npz was produced by
calling /root/reference/ragroute/rerank.py (script: tests/golden/gen_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs
may import this module.
"""
from __future__ import annotations

import numpy as np


def rerank_medrag(docs, scores, k):
    order = np.argsort(scores)[::-1]
    return [docs[i] for i in order][:k], [scores[i] for i in order][:k]


def rerank_wikipedia(docs, scores, k):
    order = np.argsort(scores)
    return [docs[i] for i in order][:k], [scores[i] for i in order][:k]


def merge_topk(scores, ids, k, descending=True, seg_len=0, list_len=0):
    """scores [B, L] fp32, ids [B, L] int64 (id < 0 = empty slot, never selected).
    Returns (scores [B,k], ids [B,k], pos [B,k] int32); unused slots are
    (-FLT_MAX, -1, -1) when descending, (+FLT_MAX, -1, -1) when ascending.

    Equal scores: lower input position first.  With ``list_len > 0`` the input is
    ``L / seg_len`` blocks (one per rank) of ``seg_len / list_len`` per-source lists and equal
    scores are ordered by (source index, id, position) -- the order that does not depend on how
    the rows were sharded over the blocks (rr_merge_topk_segments)."""
    scores = np.asarray(scores, dtype=np.float32)
    ids = np.asarray(ids, dtype=np.int64)
    B, L = scores.shape
    fmax = np.finfo(np.float32).max
    pad = np.float32(-fmax if descending else fmax)
    os_ = np.full((B, k), pad, dtype=np.float32)
    oi = np.full((B, k), -1, dtype=np.int64)
    op = np.full((B, k), -1, dtype=np.int32)
    for b in range(B):
        valid = np.nonzero((ids[b] >= 0) & ~np.isnan(scores[b]))[0]
        key = scores[b, valid].astype(np.float64)
        if list_len > 0:
            src = (valid % (seg_len if seg_len > 0 else L)) // list_len
            order = np.lexsort((valid, ids[b, valid], src, -key if descending else key))[:k]
        else:
            order = np.lexsort((valid, -key if descending else key))[:k]
        sel = valid[order]
        n = len(sel)
        os_[b, :n] = scores[b, sel]
        oi[b, :n] = ids[b, sel]
        op[b, :n] = sel
    return os_, oi, op

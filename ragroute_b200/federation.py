"""Routed federated search: router -> per-source top-k -> cross-source merge.

This is the hot path of one query batch as the reference runs it across processes:
``Router.select_relevant_sources`` (/root/reference/ragroute/router.py:221-239) ->
one ``index.search`` per selected source (data_source.py:186) -> concatenate the
per-source lists and ``rerank_medrag`` (http_server.py:280-289).  Here the three
pieces run back to back on one CUDA stream with no host synchronisation; the
router's [B, S] mask gates each source's scan on the device.

Row-sharded corpora: every rank holds rows
[rank*ceil(N/W), ...) of each source, searches its shard, and the per-shard top-k
lists (12*k bytes per query and source) are combined by ONE all-gather over
NCCL/NVLink followed by the same on-device merge kernel.  Top-k is decomposable,
so the result equals the unsharded search.
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .rerank import merge_topk
from .router import DEFAULT_THRESHOLD, FusedCorpusRouter

ROUTING_STRATEGIES = ("ragroute", "all", "random", "none")  # main.py:16, router.py:221-239
RANDOM_PICK = {"medrag": 2, "feb4rag": 9, "wikipedia": 2}  # router.py:229-234


def shard_rows(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of `rank` when n_rows are split over `world` shards."""
    per = -(-n_rows // world)
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def gather_lists(D_loc: torch.Tensor, I_loc: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather per-rank candidate lists [B, L] -> [B, W*L] (rank-major along L).

    Scores and ids travel in ONE collective: they are packed into a single byte
    buffer (4*B*L + 8*B*L bytes per rank).  Works on NCCL (CUDA tensors) and gloo
    (CPU tensors, used by the world_size-2 tests)."""
    import torch.distributed as dist
    W = dist.get_world_size(group)
    B, L = D_loc.shape
    if W == 1:
        return D_loc, I_loc
    nS, nI = B * L * 4, B * L * 8
    send = torch.empty(nS + nI, dtype=torch.uint8, device=D_loc.device)
    send[:nS].view(torch.float32).copy_(D_loc.reshape(-1))
    send[nS:].view(torch.int64).copy_(I_loc.reshape(-1))
    recv = torch.empty(W * (nS + nI), dtype=torch.uint8, device=D_loc.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(W, nS + nI)
    D_all = recv[:, :nS].contiguous().view(torch.float32).view(W, B, L)
    I_all = recv[:, nS:].contiguous().view(torch.int64).view(W, B, L)
    return (D_all.permute(1, 0, 2).reshape(B, W * L).contiguous(),
            I_all.permute(1, 0, 2).reshape(B, W * L).contiguous())


def decode_pos(pos: torch.Tensor, n_sources: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Position inside the gathered [W * S * k] list -> (rank, source index); -1 stays -1."""
    L = n_sources * k
    neg = pos < 0
    rank = torch.where(neg, pos, pos // L)
    src = torch.where(neg, pos, (pos % L) // k)
    return rank, src


class Federation:
    """A set of named sources (IndexFlatIP-shaped objects on one GPU, each possibly a
    row shard), an optional router, and the merge."""

    def __init__(self, sources: Dict[str, object], k: int, router: Optional[FusedCorpusRouter] = None,
                 routing: str = "ragroute", threshold: float = DEFAULT_THRESHOLD, descending: bool = True,
                 dataset: str = "medrag", group=None, distributed: bool = False, seed: int = 0):
        if routing not in ROUTING_STRATEGIES:
            raise ValueError(f"Unknown routing strategy: {routing}")  # router.py:239
        if routing == "ragroute" and router is None:
            raise ValueError("routing='ragroute' needs a router")
        self.names: List[str] = list(sources.keys())
        self.indexes = [sources[n] for n in self.names]
        self.k = int(k)
        self.router = router
        self.routing = routing
        self.threshold = float(threshold)
        self.descending = descending
        self.dataset = dataset
        self.group = group
        self.distributed = distributed
        self._rng = random.Random(seed)
        if router is not None and routing == "ragroute":
            assert list(router.data_sources) == self.names, "router and federation must list sources in the same order"

    # -- routing (router.py:221-239) ------------------------------------------------
    def route(self, q_router: torch.Tensor) -> torch.Tensor:
        """uint8 mask [B, S] on the device of q_router."""
        B, S = q_router.shape[0], len(self.names)
        if self.routing == "ragroute":
            return self.router.select(q_router, self.threshold)
        if self.routing == "all":
            return torch.ones((B, S), dtype=torch.uint8, device=q_router.device)
        if self.routing == "none":
            return torch.zeros((B, S), dtype=torch.uint8, device=q_router.device)
        m = np.zeros((B, S), dtype=np.uint8)
        pick = min(S, RANDOM_PICK.get(self.dataset, 2))
        for b in range(B):
            m[b, self._rng.sample(range(S), pick)] = 1
        return torch.from_numpy(m).to(q_router.device)

    # -- the batch step ---------------------------------------------------------------
    def search(self, q: torch.Tensor, q_router: Optional[torch.Tensor] = None):
        """q fp32 CUDA [B, d] (the embedding every source consumes; sources with their
        own encoder take q from `q_by_source`-style callers via search_sources).
        Returns dict(scores [B,k], ids [B,k] source-local row ids, source int32 [B,k],
        rank int32 [B,k], mask uint8 [B,S])."""
        mask = self.route(q if q_router is None else q_router)
        D_loc, I_loc = self.search_sources(q, mask)
        return self.merge(D_loc, I_loc, mask)

    def search_sources(self, q, mask: torch.Tensor):
        """Per-source (shard-local) top-k, lists side by side in [B, S*k] buffers.
        q: one CUDA tensor [B, d] for all sources, or a list with one tensor per source."""
        S, k = len(self.names), self.k
        q0 = q[0] if isinstance(q, (list, tuple)) else q
        B = q0.shape[0]
        D_loc = torch.empty((B, S * k), dtype=torch.float32, device=q0.device)
        I_loc = torch.empty((B, S * k), dtype=torch.int64, device=q0.device)
        for s, index in enumerate(self.indexes):
            qs = q[s] if isinstance(q, (list, tuple)) else q
            index.search_tensor(qs, k, mask=mask, mask_col=s, out=(D_loc, I_loc), out_col=s * k)
        return D_loc, I_loc

    def merge(self, D_loc: torch.Tensor, I_loc: torch.Tensor, mask: Optional[torch.Tensor] = None):
        S, k = len(self.names), self.k
        if self.distributed:
            D_cat, I_cat = gather_lists(D_loc, I_loc, self.group)
        else:
            D_cat, I_cat = D_loc, I_loc
        scores, ids, pos = merge_topk(D_cat, I_cat, k, descending=self.descending)
        rank, src = decode_pos(pos, S, k)
        return {"scores": scores, "ids": ids, "source": src, "rank": rank, "mask": mask}

    # -- host-buffer entry point (what a server process calls) --------------------------
    def search_host(self, q_host: torch.Tensor):
        """q_host: pinned CPU tensor [B, d].  Copies in, runs the step, copies the
        top-k (scores, ids, source) out and synchronises."""
        q = q_host.to("cuda", non_blocking=True)
        out = self.search(q)
        res = (out["scores"].to("cpu", non_blocking=True), out["ids"].to("cpu", non_blocking=True),
               out["source"].to("cpu", non_blocking=True))
        torch.cuda.current_stream().synchronize()
        return res

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

from . import _lib
from .faiss_compat import IndexFlatIP
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


class PackedLists:
    """Per-source lists of one rank as ONE byte block: [scores fp32 [B, L] | ids int64 [B, L]], each part
    padded to 16 bytes.  The search writes straight into the views, the block travels in one all-gather,
    and the merge kernel reads the gathered blocks in place (rr_merge_topk_segments) -- no packing or
    unpacking copies on the step."""

    def __init__(self, B: int, L: int, device, n_blocks: int = 1):
        self.B, self.L, self.n_blocks = B, L, n_blocks
        self.n_scores = -(-(B * L * 4) // 16) * 16
        self.block_bytes = self.n_scores + -(-(B * L * 8) // 16) * 16
        self.buf = torch.empty(n_blocks * self.block_bytes, dtype=torch.uint8, device=device)

    def scores(self, block: int = 0) -> torch.Tensor:
        o = block * self.block_bytes
        return self.buf[o:o + self.B * self.L * 4].view(torch.float32).view(self.B, self.L)

    def ids(self, block: int = 0) -> torch.Tensor:
        o = block * self.block_bytes + self.n_scores
        return self.buf[o:o + self.B * self.L * 8].view(torch.int64).view(self.B, self.L)

    def gather(self, group=None) -> "PackedLists":
        """All ranks' blocks, rank-major (one collective)."""
        import torch.distributed as dist
        W = dist.get_world_size(group)
        out = PackedLists(self.B, self.L, self.buf.device, n_blocks=W)
        dist.all_gather_into_tensor(out.buf, self.buf, group=group)
        return out


class MergeOut:
    """Result of one merge as ONE device byte block: [ids int64 | scores fp32 | source int32 | rank int32 |
    pos int32], each [B, k] -- so a host caller fetches (ids, scores, source) with a single copy."""

    def __init__(self, B: int, k: int, device):
        self.B, self.k = B, k
        n = B * k
        self.buf = torch.empty(n * 24, dtype=torch.uint8, device=device)
        self.ids = self.buf[:8 * n].view(torch.int64).view(B, k)
        self.scores = self.buf[8 * n:12 * n].view(torch.float32).view(B, k)
        self.source = self.buf[12 * n:16 * n].view(torch.int32).view(B, k)
        self.rank = self.buf[16 * n:20 * n].view(torch.int32).view(B, k)
        self.pos = self.buf[20 * n:24 * n].view(torch.int32).view(B, k)
        self.host_bytes = 16 * n  # ids, scores, source


def merge_packed(lists: PackedLists, k: int, list_len: int, descending: bool = True,
                 out: Optional[MergeOut] = None):
    """Merge every block of `lists` ([B, n_blocks * L] in block-major order) on the GPU.
    Returns (scores [B,k], ids [B,k], pos, source, rank int32 [B,k])."""
    B, dev = lists.B, lists.buf.device
    with torch.cuda.device(dev):
        o = out if out is not None else MergeOut(B, k, dev)
        assert o.B == B and o.k == k
        base = lists.buf.data_ptr()
        _lib.check(_lib.load().rr_merge_topk_segments(
            base, base + lists.n_scores, B, lists.n_blocks, lists.L, lists.block_bytes // 4, lists.block_bytes // 8,
            int(k), 1 if descending else 0, int(list_len), o.scores.data_ptr(), o.ids.data_ptr(), o.pos.data_ptr(),
            o.source.data_ptr(), o.rank.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return o.scores, o.ids, o.pos, o.source, o.rank


class Federation:
    """A set of named sources (IndexFlatIP-shaped objects on one GPU, each possibly a
    row shard), an optional router, and the merge."""

    def __init__(self, sources: Dict[str, object], k: int, router: Optional[FusedCorpusRouter] = None,
                 routing: str = "ragroute", threshold: float = DEFAULT_THRESHOLD, descending: bool = True,
                 dataset: str = "medrag", group=None, distributed: bool = False, seed: int = 0,
                 fused: bool = True, slot_of_source: Optional[Sequence[int]] = None, d_max: int = 0):
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
        self.fused = fused  # False: one rr_index_search_ex call per source (testing)
        # several query encoders (FeB4RAG, config.py:44-58): which encoder's embedding each source consumes
        # and the router's padded width (EMBEDDING_MAX_LENGTH, config.py:92-95); None = one shared embedding
        self.slot_of_source = None if slot_of_source is None else [int(x) for x in slot_of_source]
        self.d_max = int(d_max)
        self._packed: Optional[PackedLists] = None
        self._host_out: Dict[int, torch.Tensor] = {}  # pinned result buffers of search_host, per batch size
        self._all_mask: Dict[tuple, torch.Tensor] = {}  # routing "all": the all-ones mask per (batch size, device)
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
            # (kept per batch size: a constant -- callers treat the returned mask as read-only -- and no fill
            # kernel in front of every step)
            key = (B, q_router.device)
            m = self._all_mask.get(key)
            if m is None:
                m = self._all_mask[key] = torch.ones((B, S), dtype=torch.uint8, device=q_router.device)
            return m
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
        self._packed = PackedLists(B, S * k, q0.device)
        D_loc, I_loc = self._packed.scores(), self._packed.ids()
        qs_list = [(q[s] if isinstance(q, (list, tuple)) else q) for s in range(S)]
        if self.fused and all(isinstance(ix, IndexFlatIP) for ix in self.indexes):
            # one native call for the whole fan-out (http_server.py:200-209): one scan launch per source,
            # every other stage batched over the sources
            for ix, qs in zip(self.indexes, qs_list):
                assert qs.is_cuda and qs.dtype == torch.float32 and qs.is_contiguous() and qs.shape == (B, ix.d)
                assert ix.device == qs.device.index, "source and queries on different devices"
            if mask is not None:
                assert mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous() and mask.shape == (B, S)
            handles = (_lib.c_void_p * S)(*[ix._h for ix in self.indexes])
            qptrs = (_lib.c_void_p * S)(*[qs.data_ptr() for qs in qs_list])
            with torch.cuda.device(q0.device):
                _lib.check(_lib.load().rr_federation_search(
                    handles, S, qptrs, B, int(k), None if mask is None else mask.data_ptr(), S,
                    D_loc.data_ptr(), I_loc.data_ptr(), S * k, torch.cuda.current_stream().cuda_stream))
            return D_loc, I_loc
        for s, index in enumerate(self.indexes):
            index.search_tensor(qs_list[s], k, mask=mask, mask_col=s, out=(D_loc, I_loc), out_col=s * k)
        return D_loc, I_loc

    def merge(self, D_loc: torch.Tensor, I_loc: torch.Tensor, mask: Optional[torch.Tensor] = None):
        S, k = len(self.names), self.k
        packed = getattr(self, "_packed", None)
        if packed is None or packed.scores().data_ptr() != D_loc.data_ptr() or packed.ids().data_ptr() != I_loc.data_ptr():
            packed = PackedLists(D_loc.shape[0], S * k, D_loc.device)   # lists that did not come from search_sources
            packed.scores().copy_(D_loc)
            packed.ids().copy_(I_loc)
        if self.distributed:
            packed = packed.gather(self.group)
        out = MergeOut(packed.B, k, packed.buf.device)
        scores, ids, _, src, rank = merge_packed(packed, k, k, descending=self.descending, out=out)
        return {"scores": scores, "ids": ids, "source": src, "rank": rank, "mask": mask, "_block": out}

    def search_slots(self, q_slots: Sequence[torch.Tensor]):
        """One embedding per query ENCODER (q_slots[e] fp32 CUDA [B, d_e]; FeB4RAG-shaped federations):
        the router sees them zero padded side by side, every source searches with its encoder's embedding."""
        assert self.slot_of_source is not None, "this federation has a single encoder: use search(q)"
        q_list = [q_slots[e] for e in self.slot_of_source]
        q_router = None
        if self.routing == "ragroute":
            B = q_slots[0].shape[0]
            q_router = torch.zeros((B, len(q_slots), self.d_max), dtype=torch.float32, device=q_slots[0].device)
            for e, q in enumerate(q_slots):
                q_router[:, e, :q.shape[1]] = q
        mask = self.route(q_list[0] if q_router is None else q_router)
        D_loc, I_loc = self.search_sources(q_list, mask)
        return self.merge(D_loc, I_loc, mask)

    # -- host-buffer entry point (what a server process calls) --------------------------
    def search_host(self, q_host):
        """q_host: pinned CPU tensor [B, d] (or one per encoder, see search_slots).  Copies in, runs the
        step, copies the top-k (scores, ids, source) out -- one device-to-host copy of 16 k bytes per query
        into a pinned buffer kept per batch size -- and synchronises.  The returned arrays are views of that
        buffer: valid until the next call with the same batch size."""
        if isinstance(q_host, (list, tuple)):
            out = self.search_slots([q.to("cuda", non_blocking=True) for q in q_host])
        else:
            out = self.search(q_host.to("cuda", non_blocking=True))
        blk: MergeOut = out["_block"]
        B, k, n = blk.B, blk.k, blk.B * blk.k
        host = self._host_out.get(B)
        if host is None:
            host = self._host_out[B] = torch.empty(blk.host_bytes, dtype=torch.uint8).pin_memory()
        host.copy_(blk.buf[:blk.host_bytes], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return (host[8 * n:12 * n].view(torch.float32).view(B, k), host[:8 * n].view(torch.int64).view(B, k),
                host[12 * n:16 * n].view(torch.int32).view(B, k))


def place_sources(names: Sequence[str], world: int) -> Dict[str, int]:
    """Whole-source placement for federations of many small corpora (BASELINE config 5: 16 sources on
    8 GPUs, no row sharding): source i lives on rank i % world.  A rank passes an EMPTY index for the
    sources it does not hold -- an empty source returns padding rows, so the all-gather + merge of the
    row-sharded path combines the ranks unchanged."""
    return {n: i % world for i, n in enumerate(names)}

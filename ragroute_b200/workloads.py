"""Synthetic MedRAG-shaped federations built directly on the device.

The corpus embeddings of the reference live on the authors' NFS
(/root/reference/ragroute/config.py:16-24) and cannot ship; benchmarks therefore
use synthetic embeddings of the real corpus shapes:
X_s = mu_s + N(0,1)/sqrt(D) with a per-source mean of norm 0.5.  Rows are generated
in fixed global chunks so that the same global corpus results whatever the number
of row shards, and never touch host memory (the fp32 federation is 167 GB).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import synthetic as syn
from .faiss_compat import IndexFlatIP
from .federation import Federation, shard_rows
from .router import MEDRAG_THRESHOLD, FusedCorpusRouter

GEN_CHUNK_ROWS = 1 << 19  # global generation granularity (rows)


def device_rows(source_id: int, row0: int, row1: int, d: int, device: torch.device,
                buf: Optional[torch.Tensor] = None):
    """Yield (global_row_start, tensor [n, d] fp32 on device) covering [row0, row1)."""
    mu = torch.from_numpy(syn.source_mean(source_id, d)).to(device)
    scale = float(1.0 / np.sqrt(d))
    c0, c1 = row0 // GEN_CHUNK_ROWS, (max(row1, 1) - 1) // GEN_CHUNK_ROWS
    if buf is None:
        buf = torch.empty((GEN_CHUNK_ROWS, d), dtype=torch.float32, device=device)
    gen = torch.Generator(device=device)
    for c in range(c0, c1 + 1):
        g0 = c * GEN_CHUNK_ROWS
        gen.manual_seed(1_000_003 * (source_id + 1) + c)
        buf.normal_(0.0, 1.0, generator=gen)       # the whole global chunk, so shards line up
        buf.mul_(scale).add_(mu)
        lo, hi = max(row0, g0), min(row1, g0 + GEN_CHUNK_ROWS)
        if hi > lo:
            yield lo, buf[lo - g0:hi - g0]


def build_medrag_federation(rows: Dict[str, int], d: int = syn.MEDRAG_DIM, k: int = syn.MEDRAG_K,
                            rank: int = 0, world: int = 1, routing: str = "ragroute",
                            router_seed: int = 0, device: Optional[int] = None,
                            centroid_sample: int = 65536, group=None) -> Tuple[Federation, Dict[str, int]]:
    """Build this rank's row shards of every source, the router and the Federation.
    `rows`: global row count per source, in data_sources order."""
    dev_index = torch.cuda.current_device() if device is None else device
    dev = torch.device("cuda", dev_index)
    names = list(rows.keys())
    buf = torch.empty((GEN_CHUNK_ROWS, d), dtype=torch.float32, device=dev)
    sources, local_rows, centroids = {}, {}, {}
    for name in names:
        sid = syn.MEDRAG_SOURCE_TO_ID[name]
        lo, hi = shard_rows(rows[name], world, rank)
        ix = IndexFlatIP(d, device=dev_index, reserve=max(hi - lo, 1))
        for _, block in device_rows(sid, lo, hi, d, dev, buf):
            ix.add(block)
        torch.cuda.synchronize(dev)
        ix.set_id_offset(lo)
        sources[name] = ix
        local_rows[name] = hi - lo
        # router centroid = column mean of the source (from its first rows; the mean is
        # mu_s up to sampling noise, identical on every rank)
        n_c = min(rows[name], centroid_sample)
        acc = torch.zeros(d, dtype=torch.float64, device=dev)
        for _, block in device_rows(sid, 0, n_c, d, dev, buf):
            acc += block.double().sum(dim=0)
        centroids[name] = (acc / max(n_c, 1)).float().cpu().numpy()
    del buf
    router = None
    if routing == "ragroute":
        router = make_synthetic_router(names, d, centroids, router_seed, dev_index)
    fed = Federation(sources, k, router=router, routing=routing, threshold=MEDRAG_THRESHOLD,
                     dataset="medrag", distributed=world > 1, group=group)
    return fed, local_rows


def make_synthetic_router(names: Sequence[str], d: int, centroids: Dict[str, np.ndarray], seed: int,
                          device: int, calib_queries: int = 512) -> FusedCorpusRouter:
    """CorpusRoutingNN(2*d + 4) with seeded weights and scaler; fc3.bias is shifted so
    that the median logit over a calibration stream sits at the medrag threshold
    (maximally sensitive routing, about half the pairs routed)."""
    in_dim = 2 * d + len(syn.MEDRAG_SOURCE_TO_ID)
    sd = syn.router_state_dict(in_dim, seed)
    scaler = syn.router_scaler(in_dim, seed)
    tables = dict(data_sources=list(names), d_max=d,
                  model_of_source={n: "ncbi/MedCPT-Query-Encoder" for n in names},  # config.py:38-43
                  centroids=centroids, onehot_id={n: syn.MEDRAG_SOURCE_TO_ID[n] for n in names},
                  n_onehot=len(syn.MEDRAG_SOURCE_TO_ID), scaler=scaler)
    r = FusedCorpusRouter(in_dim, device=device)
    r.load_state_dict(sd)
    r.set_tables(**tables)
    sids = [syn.MEDRAG_SOURCE_TO_ID[n] for n in names]
    q = torch.from_numpy(syn.queries(calib_queries, d, sids, seed=4242)).cuda(device)
    _, logits = r.select(q, MEDRAG_THRESHOLD, return_logits=True)
    shift = float(np.log(MEDRAG_THRESHOLD / (1 - MEDRAG_THRESHOLD)) - logits.median().item())
    sd = dict(sd)
    sd["fc3.bias"] = (sd["fc3.bias"] + np.float32(shift)).astype(np.float32)
    r.load_state_dict(sd)
    r.set_tables(**tables)
    r.synthetic_state = {"sd": sd, "scaler": scaler, "centroids": centroids}
    return r

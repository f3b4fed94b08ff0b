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


CLUSTER_RUN = 128          # rows per near-duplicate run of the "clustered" corpus flavour
CLUSTER_SPREAD = 0.03      # row = run centre + CLUSTER_SPREAD * N(0,1)/sqrt(d)
CLUSTER_QUERY_SPREAD = 0.05


def _chunk_centres(source_id: int, chunk: int, d: int, device: torch.device, mu: torch.Tensor) -> torch.Tensor:
    """Run centres of one global generation chunk of the clustered flavour: [GEN_CHUNK_ROWS / CLUSTER_RUN, d]."""
    gen = torch.Generator(device=device)
    gen.manual_seed(7_000_003 * (source_id + 1) + chunk)
    c = torch.empty((GEN_CHUNK_ROWS // CLUSTER_RUN, d), dtype=torch.float32, device=device)
    c.normal_(0.0, 1.0, generator=gen)
    return c.mul_(float(1.0 / np.sqrt(d))).add_(mu)


def device_rows(source_id: int, row0: int, row1: int, d: int, device: torch.device,
                buf: Optional[torch.Tensor] = None, corpus: str = "iid"):
    """Yield (global_row_start, tensor [n, d] fp32 on device) covering [row0, row1).

    corpus = "iid": X = mu_s + N(0,1)/sqrt(d).
    corpus = "clustered": documents stored as runs of CLUSTER_RUN near-duplicate chunks (what splitting
    articles into similar passages produces): row = centre_run + 0.03 N(0,1)/sqrt(d), centre_run as iid
    above.  The whole top-k of a query near a centre lies inside one run, i.e. inside one candidate list's
    territory, so the first-pass certificate fails and the second tensor-core pass settles the query."""
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
        if corpus == "clustered":
            centres = _chunk_centres(source_id, c, d, device, mu)
            buf.mul_(CLUSTER_SPREAD * scale)
            buf.view(-1, CLUSTER_RUN, d).add_(centres.unsqueeze(1))
        else:
            buf.mul_(scale).add_(mu)
        lo, hi = max(row0, g0), min(row1, g0 + GEN_CHUNK_ROWS)
        if hi > lo:
            yield lo, buf[lo - g0:hi - g0]


def clustered_queries(n: int, d: int, rows: Dict[str, int], source_ids: Dict[str, int], seed: int,
                      device: torch.device) -> torch.Tensor:
    """Queries for the clustered flavour: each sits next to the centre of a random run of a random source
    (runs of the first generation chunk of each source, which every shard layout contains)."""
    rng = np.random.default_rng(seed)
    names = list(rows)
    pick = rng.integers(0, len(names), size=n)
    out = torch.empty((n, d), dtype=torch.float32, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(99_991 + seed)
    noise = torch.empty((n, d), dtype=torch.float32, device=device).normal_(0.0, 1.0, generator=gen)
    noise.mul_(CLUSTER_QUERY_SPREAD / float(np.sqrt(d)))
    for si, name in enumerate(names):
        sel = np.nonzero(pick == si)[0]
        if not len(sel):
            continue
        sid = source_ids[name]
        mu = torch.from_numpy(syn.source_mean(sid, d)).to(device)
        centres = _chunk_centres(sid, 0, d, device, mu)
        n_runs = max(1, min(centres.shape[0], rows[name] // CLUSTER_RUN))
        runs = torch.from_numpy(rng.integers(0, n_runs, size=len(sel))).to(device)
        out[torch.from_numpy(sel).to(device)] = centres[runs]
    return out.add_(noise)


def build_medrag_federation(rows: Dict[str, int], d: int = syn.MEDRAG_DIM, k: int = syn.MEDRAG_K,
                            rank: int = 0, world: int = 1, routing: str = "ragroute",
                            router_seed: int = 0, device: Optional[int] = None,
                            centroid_sample: int = 65536, group=None,
                            corpus: str = "iid") -> Tuple[Federation, Dict[str, int]]:
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
        for _, block in device_rows(sid, lo, hi, d, dev, buf, corpus):
            ix.add(block)
        torch.cuda.synchronize(dev)
        ix.set_id_offset(lo)
        sources[name] = ix
        local_rows[name] = hi - lo
        # router centroid = column mean of the source (from its first rows; the mean is
        # mu_s up to sampling noise, identical on every rank)
        n_c = min(rows[name], centroid_sample)
        acc = torch.zeros(d, dtype=torch.float64, device=dev)
        for _, block in device_rows(sid, 0, n_c, d, dev, buf, corpus):
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


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: 16 heterogeneous sources, mixed 384 / 768 / 1024-d encoders (SURVEY.md
# section 8d: "16 sources, D cycling {384,768,1024}, N_s log-uniform 50 k-5 M (seed 5), k=10, no scaler,
# in_dim = 2*1024+16").  The reference's FeB4RAG has 13 sources (config.py:34,44-58); this is its
# superset stress.  Sources are placed whole on the ranks (place_sources), never row-sharded.
# ---------------------------------------------------------------------------------------------------
FEB16_DIMS = (384, 768, 1024)
FEB16_K = 10
FEB16_DMAX = 1024


def feb4rag16_layout(rows_scale: float = 1.0):
    """(names, dims, rows, encoder of each source) of the synthetic 16-source federation."""
    rng = np.random.default_rng(5)
    n_rows = np.exp(rng.uniform(np.log(50_000), np.log(5_000_000), size=16))
    names = [f"src{i:02d}" for i in range(16)]
    dims = {n: FEB16_DIMS[i % 3] for i, n in enumerate(names)}
    rows = {n: max(FEB16_K, int(n_rows[i] * rows_scale)) for i, n in enumerate(names)}
    model = {n: f"enc{dims[n]}" for n in names}
    return names, dims, rows, model


def build_feb4rag16_federation(rank: int = 0, world: int = 1, device: Optional[int] = None,
                               rows_scale: float = 1.0, router_seed: int = 9, routing: str = "ragroute",
                               group=None):
    """This rank's sources (source i lives whole on rank i % world; an empty index stands for the others),
    a seeded router without scaler calibrated to route about half the pairs, and the Federation.
    Returns (fed, layout dict)."""
    from .federation import place_sources
    dev_index = torch.cuda.current_device() if device is None else device
    dev = torch.device("cuda", dev_index)
    names, dims, rows, model = feb4rag16_layout(rows_scale)
    where = place_sources(names, world)
    sources, centroids = {}, {}
    bufs = {d: torch.empty((GEN_CHUNK_ROWS, d), dtype=torch.float32, device=dev) for d in FEB16_DIMS}
    for i, n in enumerate(names):
        d, sid = dims[n], 100 + i
        mine = where[n] == rank
        ix = IndexFlatIP(d, device=dev_index, reserve=rows[n] if mine else 0)
        if mine:
            for _, block in device_rows(sid, 0, rows[n], d, dev, bufs[d]):
                ix.add(block)
            torch.cuda.synchronize(dev)
        sources[n] = ix
        n_c = min(rows[n], 65536)          # centroid from the first rows: identical on every rank
        acc = torch.zeros(d, dtype=torch.float64, device=dev)
        for _, block in device_rows(sid, 0, n_c, d, dev, bufs[d]):
            acc += block.double().sum(dim=0)
        centroids[n] = np.pad((acc / n_c).float().cpu().numpy(), (0, FEB16_DMAX - d)).astype(np.float32)
    del bufs
    router = None
    slot_names = sorted(set(model.values()))
    if routing == "ragroute":
        in_dim = 2 * FEB16_DMAX + len(names)
        sd = syn.router_state_dict(in_dim, router_seed)
        onehot = {n: i for i, n in enumerate(names)}

        def make(sd_):
            r = FusedCorpusRouter(in_dim, device=dev_index)
            r.load_state_dict(sd_)
            r.set_tables(names, FEB16_DMAX, model, centroids, onehot, len(names), scaler=None)
            return r

        router = make(sd)
        qs = feb4rag16_queries(512, 4242, names, dims)
        qr = pad_slots([torch.from_numpy(qs[m]).to(dev) for m in router.slot_names], FEB16_DMAX)
        _, lg = router.select(qr, 0.5, return_logits=True)
        sd = dict(sd)
        sd["fc3.bias"] = (sd["fc3.bias"] - np.float32(lg.median().item())).astype(np.float32)
        router = make(sd)
        slot_names = router.slot_names
    slot_of_source = [slot_names.index(model[n]) for n in names]
    fed = Federation(sources, FEB16_K, router=router, routing=routing, threshold=0.5, dataset="feb4rag",
                     distributed=world > 1, group=group, slot_of_source=slot_of_source, d_max=FEB16_DMAX)
    layout = {"names": names, "dims": dims, "rows": rows, "model": model, "placement": where,
              "slot_names": slot_names, "slot_dims": [int(m[3:]) for m in slot_names]}
    return fed, layout


def feb4rag16_queries(n: int, seed: int, names: Sequence[str], dims: Dict[str, int]) -> Dict[str, np.ndarray]:
    """One embedding per encoder and query (host): enc_d -> [n, d], each query near the mean of a random
    source of that dimension."""
    rng = np.random.default_rng(seed)
    out = {}
    for d in FEB16_DIMS:
        srcs = [100 + i for i, nm in enumerate(names) if dims[nm] == d]
        pick = rng.integers(0, len(srcs), size=n)
        mus = np.stack([syn.source_mean(s, d) for s in srcs])
        out[f"enc{d}"] = (rng.standard_normal((n, d), dtype=np.float32) * np.float32(1.0 / np.sqrt(d)) + mus[pick]).astype(np.float32)
    return out


def pad_slots(q_slots: Sequence[torch.Tensor], d_max: int) -> torch.Tensor:
    """Per-encoder embeddings [B, d_slot] -> the router's input [B, n_slots, d_max], zero padded
    (np.pad(query_embed, (0, d_max - len)), router.py:248)."""
    B = q_slots[0].shape[0]
    out = torch.zeros((B, len(q_slots), d_max), dtype=torch.float32, device=q_slots[0].device)
    for s, q in enumerate(q_slots):
        out[:, s, :q.shape[1]] = q
    return out

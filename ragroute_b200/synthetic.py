"""Seeded synthetic inputs for tests and benchmarks (host side, numpy only).

No corpus embeddings, router weights or scaler ship with the reference (its
loaders read absolute NFS paths: /root/reference/ragroute/router.py:108-130,
/root/reference/ragroute/data_source.py:45-57), so every test and benchmark input
is synthetic with the reference's shapes.

``hash_uniform`` is a counter-based generator made of integer operations only
(splitmix64), so the same seed gives bit-identical arrays on every platform and
numpy version; the golden fixtures under tests/golden/ store seeds, not weights.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def hash_uniform(seed: int, shape, lo: float = -1.0, hi: float = 1.0) -> np.ndarray:
    """float32 array, uniform in [lo, hi), bit-reproducible everywhere."""
    n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
    with np.errstate(over="ignore"):
        ctr = np.arange(n, dtype=np.uint64) + (np.uint64(seed) << np.uint64(32))
        bits = _splitmix64(_splitmix64(ctr))
    u = (bits >> np.uint64(40)).astype(np.float64) * (1.0 / (1 << 24))  # 24-bit, exact
    return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


# ---------------------------------------------------------------------------
# Router (CorpusRoutingNN-shaped) parameters, /root/reference/ragroute/router.py:37-48
# ---------------------------------------------------------------------------

ROUTER_H1 = 256
ROUTER_H2 = 128


def router_state_dict(in_dim: int, seed: int = 0) -> Dict[str, np.ndarray]:
    """Random parameters with torch's default Linear init ranges U(-1/sqrt(fan_in), ..)
    and non-trivial LayerNorm affine terms, keyed like the reference's state_dict."""
    def lin(s, out_f, in_f):
        b = 1.0 / np.sqrt(in_f)
        return hash_uniform(s, (out_f, in_f), -b, b), hash_uniform(s + 1, (out_f,), -b, b)

    w1, b1 = lin(seed * 100 + 1, ROUTER_H1, in_dim)
    w2, b2 = lin(seed * 100 + 3, ROUTER_H2, ROUTER_H1)
    w3, b3 = lin(seed * 100 + 5, 1, ROUTER_H2)
    return {
        "fc1.weight": w1, "fc1.bias": b1,
        "ln1.weight": hash_uniform(seed * 100 + 7, (ROUTER_H1,), 0.5, 1.5),
        "ln1.bias": hash_uniform(seed * 100 + 8, (ROUTER_H1,), -0.2, 0.2),
        "fc2.weight": w2, "fc2.bias": b2,
        "ln2.weight": hash_uniform(seed * 100 + 9, (ROUTER_H2,), 0.5, 1.5),
        "ln2.bias": hash_uniform(seed * 100 + 10, (ROUTER_H2,), -0.2, 0.2),
        "fc3.weight": w3, "fc3.bias": b3,
    }


def router_scaler(in_dim: int, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """(mean_, scale_) float64 [in_dim], shaped like sklearn StandardScaler's."""
    mean = hash_uniform(seed * 100 + 20, (in_dim,), -0.1, 0.1).astype(np.float64)
    scale = hash_uniform(seed * 100 + 21, (in_dim,), 0.02, 0.2).astype(np.float64)
    return mean, scale


# ---------------------------------------------------------------------------
# Corpora / queries: X = mu_s + N(0,1)/sqrt(D), |mu_s| = 0.5
# ---------------------------------------------------------------------------

MEDRAG_SOURCES = ("pubmed", "statpearls", "textbooks", "wikipedia")  # config.py:33
MEDRAG_ROWS = {"pubmed": 23_900_000, "statpearls": 301_000, "textbooks": 125_000,
               "wikipedia": 29_900_000}
MEDRAG_SOURCE_TO_ID = {"pubmed": 0, "statpearls": 1, "textbooks": 2, "wikipedia": 3}  # config.py:88-90
MEDRAG_DIM = 768  # config.py:93
MEDRAG_K = 32  # config.py:98


def source_mean(source_id: int, d: int, norm: float = 0.5) -> np.ndarray:
    rng = np.random.default_rng(7 + source_id)
    mu = rng.standard_normal(d)
    return (mu * (norm / np.linalg.norm(mu))).astype(np.float32)


def corpus_rows(source_id: int, n: int, d: int, seed_offset: int = 0,
                mu: Optional[np.ndarray] = None) -> np.ndarray:
    """Host-side corpus block (tests and the CPU-baseline sample)."""
    rng = np.random.default_rng(1000 + source_id * 16 + seed_offset)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x *= np.float32(1.0 / np.sqrt(d))
    x += source_mean(source_id, d) if mu is None else mu
    return x


def queries(n: int, d: int, source_ids: Sequence[int], seed: int = 42) -> np.ndarray:
    """Each query = mu_{s*} + N(0,1)/sqrt(D), s* uniform over the given sources."""
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, len(source_ids), size=n)
    q = rng.standard_normal((n, d), dtype=np.float32) * np.float32(1.0 / np.sqrt(d))
    mus = np.stack([source_mean(s, d) for s in source_ids])
    return (q + mus[pick]).astype(np.float32)

"""oracle/router_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy, fp32 network / float64 scaler exactly as the reference
mixes them) of the router's relevance inference:

* ``mlp_logits``        <- ``CorpusRoutingNN.forward``
  (/root/reference/ragroute/router.py:37-55): fc1 -> LayerNorm(256) -> ReLU ->
  fc2 -> LayerNorm(128) -> ReLU -> fc3; dropout is the identity in eval mode
  (router.py:119); LayerNorm eps = 1e-5, biased variance, affine (torch defaults).
* ``build_features``    <- ``Router.select_relevant_sources_ragroute``
  (router.py:241-267): per source ``[pad(q_model(s), Dmax) | centroid_s |
  onehot(id_s)]``; the one-hot comes from ``np.eye`` so the row is float64.
* ``select``            <- router.py:269-283: optional ``StandardScaler.transform``
  in float64 (``(x - mean_) / scale_``), cast to fp32, MLP, sigmoid, strict
  ``> threshold`` (0.4924 medrag, 0.5 otherwise), names in ``data_sources`` order.

Pinned against the reference itself: tests/golden/router_*.npz hold inputs and
outputs produced by importing /root/reference/ragroute/router.py (script:
tests/golden/gen_golden.py); tests/test_oracle_router.py replays them.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs
may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

MEDRAG_THRESHOLD = 0.4924  # router.py:278
DEFAULT_THRESHOLD = 0.5  # router.py:280
LN_EPS = 1e-5


def _layer_norm(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    x = x.astype(np.float32)
    mean = x.mean(axis=-1, keepdims=True, dtype=np.float32)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True, dtype=np.float32)
    return ((x - mean) / np.sqrt(var + np.float32(LN_EPS))).astype(np.float32) * w + b


def mlp_logits(sd: Dict[str, np.ndarray], x: np.ndarray) -> np.ndarray:
    """x fp32 [P, in_dim] -> logits fp32 [P, 1] (router.py:50-55)."""
    x = np.asarray(x, dtype=np.float32)
    h = x @ sd["fc1.weight"].T + sd["fc1.bias"]
    h = np.maximum(_layer_norm(h, sd["ln1.weight"], sd["ln1.bias"]), 0)
    h = h @ sd["fc2.weight"].T + sd["fc2.bias"]
    h = np.maximum(_layer_norm(h, sd["ln2.weight"], sd["ln2.bias"]), 0)
    return (h @ sd["fc3.weight"].T + sd["fc3.bias"]).astype(np.float32)


def build_features(query_embeddings: Dict[str, np.ndarray], data_sources: Sequence[str],
                   model_of_source: Dict[str, str], centroids: Dict[str, np.ndarray],
                   onehot_id: Dict[str, int], n_onehot: int, d_max: int) -> np.ndarray:
    """router.py:244-267. Returns float64 [S, 2*d_max + n_onehot]."""
    padded = {m: np.pad(q, (0, d_max - len(q))) for m, q in query_embeddings.items()}
    rows = []
    for corpus in data_sources:
        feats = np.concatenate([padded[model_of_source[corpus]], centroids[corpus]])
        feats = np.concatenate([feats, np.eye(n_onehot)[onehot_id[corpus]]])
        rows.append(feats)
    return np.asarray(rows)


def select(sd: Dict[str, np.ndarray], features: np.ndarray, threshold: float,
           scaler_mean: Optional[np.ndarray] = None, scaler_scale: Optional[np.ndarray] = None):
    """router.py:269-280. Returns (bool mask [S], logits fp32 [S])."""
    inputs = np.asarray(features)
    if scaler_mean is not None:
        inputs = (inputs.astype(np.float64) - scaler_mean) / scaler_scale  # StandardScaler.transform
    x = inputs.astype(np.float32)
    logits = mlp_logits(sd, x).reshape(-1)
    prob = (np.float32(1) / (np.float32(1) + np.exp(-logits, dtype=np.float32))).astype(np.float32)
    return prob > np.float32(threshold), logits


def select_names(mask: np.ndarray, data_sources: Sequence[str]) -> List[str]:
    """router.py:282."""
    return [c for p, c in zip(mask, data_sources) if p]


def logit_threshold(p: float) -> float:
    """sigmoid(l) > p  <=>  l > log(p / (1 - p))."""
    return float(np.log(p / (1.0 - p)))

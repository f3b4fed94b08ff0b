"""Loaders for the router artefacts the reference reads at start-up.

``Router.load_router`` (/root/reference/ragroute/router.py:106-151) gathers, from absolute paths on
the authors' NFS, a ``state_dict`` (``best_model.pth``), a scikit-learn ``StandardScaler`` -- the 4th
item of a pickled 5-tuple for medrag (:123-125), a bare pickle for wikipedia (:128-130), none for
feb4rag -- and one ``*_stats.json`` per corpus whose ``"centroid"`` is zero-padded to
``EMBEDDING_MAX_LENGTH`` (:149-151; wikipedia: one ``cluster_stats.json`` holding a list, indexed by
``int(corpus)``, :142-145).  ``load_router`` here takes the same files from explicit paths and
returns a ``FusedCorpusRouter`` with its tables set, ready for ``select`` /
``select_relevant_sources``.  Pure host-side file handling; the arithmetic stays in the library.
"""
from __future__ import annotations

import json
import pickle
from typing import Dict, Optional, Sequence

import numpy as np

# ragroute/config.py:32-36, :72-96; ragroute/router.py:32-34, :259-265, :278-280
DATASETS = {
    "medrag": dict(d_max=768, n_onehot=4, threshold=0.4924, input_dim=1540,
                   source_to_id={"pubmed": 0, "statpearls": 1, "textbooks": 2, "wikipedia": 3}),
    "feb4rag": dict(d_max=4096, n_onehot=13, threshold=0.5, input_dim=8205,
                    source_to_id={"arguana": 0, "climate-fever": 1, "dbpedia-entity": 2, "fever": 3, "fiqa": 4,
                                  "hotpotqa": 5, "msmarco": 6, "nfcorpus": 7, "nq": 8, "scidocs": 9,
                                  "scifact": 10, "trec-covid": 11, "webis-touche2020": 12}),
    "wikipedia": dict(d_max=768, n_onehot=10, threshold=0.5, input_dim=1546,
                      source_to_id={str(i): i for i in range(10)}),
}


def load_scaler(dataset: str, path: Optional[str]):
    """router.py:121-130: medrag -> 4th item of the pickled (train, val, test, scaler, val_qs) tuple
    written by scripts/train/train_medrag_router.py:240-245; wikipedia -> the pickled scaler itself;
    feb4rag -> no scaler."""
    if dataset == "feb4rag" or path is None:
        return None
    with open(path, "rb") as f:
        obj = pickle.load(f)
    if dataset == "medrag":
        _, _, _, scaler, _ = obj
        return scaler
    return obj


def load_centroids(dataset: str, data_sources: Sequence[str], stats_files, d_max: int) -> Dict[str, np.ndarray]:
    """router.py:132-151.  ``stats_files``: dict corpus -> path of its ``*_stats.json`` (medrag, feb4rag),
    or one path of ``cluster_stats.json`` (wikipedia: a list indexed by ``int(corpus)``)."""
    out = {}
    for corpus in data_sources:
        if dataset == "wikipedia":
            with open(stats_files if isinstance(stats_files, str) else stats_files[corpus], "r") as f:
                stats = json.load(f)[int(corpus)]
        else:
            with open(stats_files[corpus], "r") as f:
                stats = json.load(f)
        c = np.array(stats["centroid"], dtype=np.float32)
        if len(c) > d_max:
            raise ValueError(f"centroid of {corpus} has {len(c)} dims > EMBEDDING_MAX_LENGTH {d_max}")
        out[corpus] = np.pad(c, (0, d_max - len(c)))
    return out


def load_router(dataset: str, data_sources: Sequence[str], model_of_source: Dict[str, str], model_path: str,
                stats_files, scaler_path: Optional[str] = None, device: Optional[int] = None):
    """The whole of ``Router.load_router`` on explicit paths -> (FusedCorpusRouter, threshold)."""
    import torch
    from .router import FusedCorpusRouter
    if dataset not in DATASETS:
        raise ValueError(f"unknown dataset {dataset!r}")  # main.py:14 restricts the choices the same way
    cfg = DATASETS[dataset]
    sd = torch.load(model_path, map_location="cpu")
    in_dim = int(sd["fc1.weight"].shape[1])
    if in_dim != 2 * cfg["d_max"] + cfg["n_onehot"]:
        # the checked-in trainers disagree with the server about the one-hot columns (SURVEY 8a notes):
        # the layout of the state_dict decides, but it must be one the feature builder can produce
        raise ValueError(f"{model_path}: fc1 expects {in_dim} inputs, the {dataset} server builds "
                         f"{2 * cfg['d_max'] + cfg['n_onehot']} (router.py:32-34, :259-265)")
    router = FusedCorpusRouter(in_dim, device=device)
    router.load_state_dict({k: v for k, v in sd.items()})
    router.set_tables(list(data_sources), cfg["d_max"], model_of_source,
                      load_centroids(dataset, data_sources, stats_files, cfg["d_max"]),
                      {c: cfg["source_to_id"][c] for c in data_sources}, cfg["n_onehot"],
                      scaler=load_scaler(dataset, scaler_path))
    return router, cfg["threshold"]

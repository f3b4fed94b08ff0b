"""Generate golden fixtures by running the REFERENCE's own code.

Run in the authoring container only (needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/gen_golden.py

It imports /root/reference/ragroute/router.py and rerank.py unchanged (two absent
third-party modules are stubbed in sys.modules; they are only needed by the query
encoders, which are out of scope) and records

* router_{medrag,feb4rag,wikipedia}.npz -- for seeded weights / scaler / centroids
  (regenerated from seeds by ragroute_b200.synthetic, so only seeds are stored)
  and seeded query embeddings: the logits the reference feeds to torch.sigmoid
  (router.py:276) and the source list it returns (router.py:282);
* rerank.npz -- inputs and outputs of rerank_medrag / rerank_wikipedia
  (rerank.py:3-9, :28-34) on tie-free random score lists, incl. short and empty.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

from ragroute_b200 import synthetic as syn  # noqa: E402

# query encoders (out of scope) need these; stub them so router.py imports
for name in ("sentence_transformers", "sentence_transformers.models"):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.SentenceTransformer = object
        m.Transformer = object
        m.Pooling = object
        sys.modules[name] = m

import torch  # noqa: E402
import ragroute.router as ref_router  # noqa: E402
import ragroute.rerank as ref_rerank  # noqa: E402
from ragroute import config as ref_config  # noqa: E402

N_QUERIES = 24


class _Scaler:
    """Stand-in with sklearn StandardScaler.transform's arithmetic
    (sklearn/preprocessing/_data.py: X -= mean_; X /= scale_ on a float64 copy);
    the reference un-pickles a real StandardScaler (router.py:123-130)."""

    def __init__(self, mean, scale):
        self.mean_, self.scale_ = mean, scale

    def transform(self, X):
        X = np.array(X, dtype=np.float64)
        X -= self.mean_
        X /= self.scale_
        return X


def _use_sklearn_if_present(mean, scale):
    try:
        from sklearn.preprocessing import StandardScaler
        sc = StandardScaler()
        sc.mean_, sc.scale_, sc.var_ = mean, scale, scale ** 2
        sc.n_features_in_ = len(mean)
        sc.n_samples_seen_ = 1
        return sc
    except Exception:
        return _Scaler(mean, scale)


def make_reference_router(dataset: str, seed: int, bias_shift: float = 0.0):
    sources = ref_config.DATA_SOURCES[dataset]
    in_dim = {"medrag": ref_router.MEDRAG_INPUT_DIMENSION,
              "feb4rag": ref_router.FEDRAG_INPUT_DIMENSION,
              "wikipedia": ref_router.WIKIPEDIA_INPUT_DIMENSION}[dataset]
    d_max = ref_config.EMBEDDING_MAX_LENGTH[dataset]
    r = ref_router.Router.__new__(ref_router.Router)
    r.dataset, r.data_sources = dataset, sources
    r.routing_strategy, r.simulate, r.device = "ragroute", False, "cpu"
    net = ref_router.CorpusRoutingNN(in_dim)
    sd = syn.router_state_dict(in_dim, seed)
    sd["fc3.bias"] = (sd["fc3.bias"] + np.float32(bias_shift)).astype(np.float32)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd.items()})
    net.eval()
    r.router = net
    r.scaler = None
    if dataset in ("medrag", "wikipedia"):
        mean, scale = syn.router_scaler(in_dim, seed)
        r.scaler = _use_sklearn_if_present(mean, scale)
    # centroids: router.py:149-151 (fp32, zero-padded to d_max)
    r.centroids = {}
    for i, c in enumerate(sources):
        d_c = source_dim(dataset, c)
        cen = syn.hash_uniform(seed * 1000 + 500 + i, (d_c,), -0.3, 0.3)
        r.centroids[c] = np.pad(cen, (0, d_max - d_c))
    return r, in_dim, d_max


# embedding width of each FeB4RAG encoder (model cards; config.py:44-58 names them)
FEB4RAG_MODEL_DIM = {
    "e5-large": 1024, "SGPT-5.8B-weightedmean-msmarco-specb-bitfit": 4096,
    "UAE-Large-V1": 1024, "all-mpnet-base-v2": 768, "multilingual-e5-large": 1024,
    "ember-v1": 1024, "e5-base": 768, "gte-base": 768,
}


def source_dim(dataset: str, corpus: str) -> int:
    if dataset == "feb4rag":
        return FEB4RAG_MODEL_DIM[ref_config.EMBEDDING_MODELS_PER_DATA_SOURCE[dataset][corpus][0]]
    return 768


def gen_router(dataset: str, seed: int, bias_shift: float = 0.0):
    r, in_dim, d_max = make_reference_router(dataset, seed, bias_shift)
    models = sorted({ref_config.EMBEDDING_MODELS_PER_DATA_SOURCE[dataset][c][0] for c in r.data_sources})
    model_dim = {m: (FEB4RAG_MODEL_DIM[m] if dataset == "feb4rag" else 768) for m in models}
    S = len(r.data_sources)
    logits = np.zeros((N_QUERIES, S), dtype=np.float32)
    masks = np.zeros((N_QUERIES, S), dtype=np.bool_)
    captured = {}
    real_sigmoid = torch.sigmoid

    def spy(t):
        captured["logits"] = t.detach().clone()
        return real_sigmoid(t)

    for qi in range(N_QUERIES):
        qe = {m: syn.hash_uniform(seed * 100000 + 7000 + qi * 16 + mi, (model_dim[m],), -0.4, 0.4)
              for mi, m in enumerate(models)}
        torch.sigmoid = spy
        try:
            names = r.select_relevant_sources(qe)  # reference code, router.py:221-283
        finally:
            torch.sigmoid = real_sigmoid
        logits[qi] = captured["logits"].numpy()
        masks[qi] = [c in names for c in r.data_sources]
    out = os.path.join(HERE, f"router_{dataset}.npz")
    np.savez(out, seed=seed, fc3_bias_shift=np.float32(bias_shift), n_queries=N_QUERIES, in_dim=in_dim, d_max=d_max,
             data_sources=np.array(r.data_sources), models=np.array(models),
             model_dims=np.array([model_dim[m] for m in models]),
             model_of_source=np.array([ref_config.EMBEDDING_MODELS_PER_DATA_SOURCE[dataset][c][0]
                                       for c in r.data_sources]),
             source_dims=np.array([source_dim(dataset, c) for c in r.data_sources]),
             onehot_ids=np.array(onehot_ids(dataset, r.data_sources)),
             logits=logits, masks=masks,
             threshold=np.float64(0.4924 if dataset == "medrag" else 0.5))
    print(out, "selected fraction", masks.mean(), "logit range", logits.min(), logits.max())


def onehot_ids(dataset, sources):
    if dataset == "medrag":
        return [ref_config.MEDRAG_SOURCE_TO_ID[c] for c in sources]
    if dataset == "feb4rag":
        return [ref_config.FEB4RAG_SOURCE_TO_ID[c] for c in sources]
    return [int(c) for c in sources]


def gen_rerank():
    rng = np.random.default_rng(123)
    cases = {}
    for ci, (L, k) in enumerate([(128, 32), (96, 32), (20, 32), (0, 32), (40, 10), (7, 10), (1, 1)]):
        scores = rng.permutation(4096)[:L].astype(np.float32) / np.float32(64.0)  # distinct
        docs = list(range(L))
        s_list = [float(v) for v in scores]  # http_server.py:286 passes Python floats
        d1, s1 = ref_rerank.rerank_medrag(docs, s_list, k)
        d2, s2 = ref_rerank.rerank_wikipedia(docs, s_list, k)
        cases[f"c{ci}_scores"] = scores
        cases[f"c{ci}_k"] = np.int64(k)
        cases[f"c{ci}_medrag_docs"] = np.array(d1, dtype=np.int64)
        cases[f"c{ci}_medrag_scores"] = np.array(s1, dtype=np.float64)
        cases[f"c{ci}_wikipedia_docs"] = np.array(d2, dtype=np.int64)
        cases[f"c{ci}_wikipedia_scores"] = np.array(s2, dtype=np.float64)
    cases["n_cases"] = np.int64(7)
    out = os.path.join(HERE, "rerank.npz")
    np.savez(out, **cases)
    print(out)


if __name__ == "__main__":
    torch.set_num_threads(1)
    gen_router("medrag", seed=1)
    gen_router("feb4rag", seed=2)
    gen_router("wikipedia", seed=3, bias_shift=-0.6)
    gen_rerank()

"""Pins the numpy oracles (oracle/router_ref.py, oracle/rerank_ref.py) to the
golden vectors recorded from the reference's own code (tests/golden/gen_golden.py).
Tolerance: logits within 1e-5 absolute of the reference's torch-CPU logits;
decisions identical outside a +-1e-5 logit band around the threshold."""
import os

import numpy as np
import pytest

from oracle import rerank_ref, router_ref
from golden_util import GOLDEN, RouterCase


@pytest.mark.parametrize("dataset", ["medrag", "feb4rag", "wikipedia"])
def test_router_oracle_matches_reference_golden(dataset):
    c = RouterCase(dataset)
    thr_logit = router_ref.logit_threshold(c.threshold)
    for qi in range(c.n_queries):
        feats = router_ref.build_features(c.query(qi), c.data_sources, c.model_of_source,
                                          c.centroids, c.onehot_ids, c.n_onehot, c.d_max)
        assert feats.shape == (len(c.data_sources), c.in_dim)
        mean, scale = c.scaler if c.scaler else (None, None)
        mask, logits = router_ref.select(c.sd, feats, c.threshold, mean, scale)
        np.testing.assert_allclose(logits, c.ref_logits[qi], rtol=0, atol=1e-5)
        decided = np.abs(c.ref_logits[qi] - thr_logit) > 1e-5
        assert np.array_equal(mask[decided], c.ref_masks[qi][decided])


def test_rerank_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "rerank.npz"))
    for ci in range(int(g["n_cases"])):
        scores = g[f"c{ci}_scores"]
        k = int(g[f"c{ci}_k"])
        docs = list(range(len(scores)))
        s_list = [float(v) for v in scores]
        d1, s1 = rerank_ref.rerank_medrag(docs, s_list, k)
        d2, s2 = rerank_ref.rerank_wikipedia(docs, s_list, k)
        assert d1 == g[f"c{ci}_medrag_docs"].tolist() and s1 == g[f"c{ci}_medrag_scores"].tolist()
        assert d2 == g[f"c{ci}_wikipedia_docs"].tolist() and s2 == g[f"c{ci}_wikipedia_scores"].tolist()
        # array form agrees with list form on tie-free input
        L = len(scores)
        if L:
            ms, mi, mp = rerank_ref.merge_topk(scores[None], np.arange(L)[None], k, descending=True)
            n = min(k, L)
            assert mi[0, :n].tolist() == d1 and (mi[0, n:] == -1).all()
            ms, mi, mp = rerank_ref.merge_topk(scores[None], np.arange(L)[None], k, descending=False)
            assert mi[0, :n].tolist() == d2

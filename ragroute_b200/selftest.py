"""smoke(): one small routed search on cuda:0, checked against the CPU oracle.
(The oracle is test infrastructure; it is imported here only as the checker.)"""
from __future__ import annotations

import numpy as np
import torch

from . import synthetic as syn
from .faiss_compat import IndexFlatIP
from .federation import Federation
from .router import MEDRAG_THRESHOLD, FusedCorpusRouter


def build_small_medrag(n_rows=(6000, 9000), d=768, seed=11, device=0):
    """Two MedRAG-shaped sources + a seeded router over them (in_dim = 2*768 + 4)."""
    names = ["statpearls", "textbooks"]
    sids = [syn.MEDRAG_SOURCE_TO_ID[n] for n in names]
    corpora = {n: syn.corpus_rows(sid, nr, d) for n, sid, nr in zip(names, sids, n_rows)}
    in_dim = 2 * d + 4
    sd = syn.router_state_dict(in_dim, seed)
    scaler = syn.router_scaler(in_dim, seed)
    centroids = {n: corpora[n].mean(axis=0).astype(np.float32) for n in names}
    return names, sids, corpora, sd, scaler, centroids


def smoke() -> None:
    from oracle import flat_ip, rerank_ref, router_ref

    torch.cuda.set_device(0)
    d, k, B = 768, 32, 8
    names, sids, corpora, sd, scaler, centroids = build_small_medrag(d=d)
    q = syn.queries(B, d, sids, seed=5)

    # seeded router whose fc3 bias is calibrated so that about half the pairs are routed
    from .workloads import make_synthetic_router
    router = make_synthetic_router(names, d, centroids, 11, 0)
    sd, scaler = router.synthetic_state["sd"], router.synthetic_state["scaler"]
    sources = {}
    for n in names:
        ix = IndexFlatIP(d)
        ix.add(corpora[n])
        sources[n] = ix
    fed = Federation(sources, k, router=router, routing="ragroute", threshold=MEDRAG_THRESHOLD)
    out = fed.search(torch.from_numpy(q).cuda())
    torch.cuda.synchronize()
    got_s, got_i, got_src = out["scores"].cpu().numpy(), out["ids"].cpu().numpy(), out["source"].cpu().numpy()
    got_mask = out["mask"].cpu().numpy().astype(bool)

    # oracle: reference router arithmetic, FAISS restatement per source, reference merge
    thr_logit = router_ref.logit_threshold(MEDRAG_THRESHOLD)
    for b in range(B):
        feats = router_ref.build_features({"m": q[b]}, names, {n: "m" for n in names}, centroids,
                                          {n: syn.MEDRAG_SOURCE_TO_ID[n] for n in names}, 4, d)
        mask, logits = router_ref.select(sd, feats, MEDRAG_THRESHOLD, scaler[0], scaler[1])
        decided = np.abs(logits - thr_logit) > 1e-5
        assert np.array_equal(mask[decided], got_mask[b][decided]), f"router decision differs for query {b}"
        all_s, all_key = [], []
        for s, n in enumerate(names):
            if got_mask[b, s]:
                Ds, Is = flat_ip.search_seq(q[b:b + 1], corpora[n], k)
                all_s += Ds[0].tolist()
                all_key += [(s, int(i)) for i in Is[0]]
        _, ref_scores = rerank_ref.rerank_medrag(all_key, all_s, k)
        n_ref = len(ref_scores)
        assert (got_i[b, n_ref:] == -1).all(), "padding differs"
        np.testing.assert_allclose(got_s[b, :n_ref], np.asarray(ref_scores, dtype=np.float32), rtol=1e-5)
        ref_keys, _ = rerank_ref.rerank_medrag(all_key, all_s, k)
        got_keys = list(zip(got_src[b, :n_ref].tolist(), got_i[b, :n_ref].tolist()))
        if set(got_keys) != set(ref_keys):  # only near-ties may differ
            diff = set(got_keys) ^ set(ref_keys)
            kth = ref_scores[-1]
            for (s, i) in diff:
                sc = float(corpora[names[s]][i].astype(np.float64) @ q[b].astype(np.float64))
                assert abs(sc - kth) <= 1e-5 * abs(kth), f"query {b}: id {(s, i)} differs and is not a tie"
    assert 0.1 < got_mask.mean() < 0.9, "smoke router routes nothing / everything: the search is not exercised"
    st = {n: sources[n].stats() for n in names}
    print("smoke ok: routed top-%d over %s, mask density %.2f, exact rescans %s" %
          (k, names, got_mask.mean(), {n: st[n]["exact_rescans"] for n in names}))

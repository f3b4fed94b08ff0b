"""The whole hot path on one GPU: router -> routed per-source search -> merge,
against the reference arithmetic restated by the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_smoke_entry_point(cuda_required):
    import __graft_entry__ as entry
    entry.smoke()


@pytest.mark.parametrize("routing", ["all", "none", "random"])
def test_routing_strategies(cuda_required, routing):
    from oracle import flat_ip, rerank_ref
    from ragroute_b200 import synthetic as syn
    from ragroute_b200.faiss_compat import IndexFlatIP
    from ragroute_b200.federation import Federation
    d, k, B = 768, 32, 6
    names = ["pubmed", "statpearls", "textbooks", "wikipedia"]
    corpora = {n: syn.corpus_rows(i, 3000 + 500 * i, d) for i, n in enumerate(names)}
    sources = {}
    for n in names:
        ix = IndexFlatIP(d)
        ix.add(corpora[n])
        sources[n] = ix
    fed = Federation(sources, k, routing=routing, dataset="medrag", seed=1)
    q = syn.queries(B, d, [0, 1, 2, 3], seed=6)
    out = fed.search(torch.from_numpy(q).cuda())
    torch.cuda.synchronize()
    mask = out["mask"].cpu().numpy().astype(bool)
    if routing == "all":
        assert mask.all()
    elif routing == "none":
        assert not mask.any() and (out["ids"] == -1).all().item()      # router.py:236-237: nothing searched
    else:
        assert (mask.sum(axis=1) == 2).all()                           # router.py:229-230: random.sample(.., 2)
    got_s, got_i, got_src = (out[x].cpu().numpy() for x in ("scores", "ids", "source"))
    for b in range(B):
        all_s, all_key = [], []
        for s, n in enumerate(names):
            if mask[b, s]:
                Ds, Is = flat_ip.search_seq(q[b:b + 1], corpora[n], k)
                all_s += Ds[0].tolist()
                all_key += [(s, int(i)) for i in Is[0]]
        ref_keys, ref_scores = rerank_ref.rerank_medrag(all_key, all_s, k)
        n_ref = len(ref_keys)
        np.testing.assert_allclose(got_s[b, :n_ref], np.asarray(ref_scores, dtype=np.float32), rtol=1e-5)
        assert list(zip(got_src[b, :n_ref].tolist(), got_i[b, :n_ref].tolist())) == ref_keys
        assert (got_i[b, n_ref:] == -1).all()
    with pytest.raises(ValueError):
        Federation(sources, k, routing="bogus")

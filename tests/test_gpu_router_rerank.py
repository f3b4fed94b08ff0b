"""Parity of the fused router kernel and the merge kernel (through the C ABI) with
the golden vectors recorded from the reference's own code and with the numpy
oracles.  Tolerances: logits within 1e-5 absolute; source selections identical
except logits within 1e-5 of the decision threshold; merge bit-exact."""
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN, RouterCase
from oracle import rerank_ref, router_ref

pytestmark = pytest.mark.gpu


def _make_router(c):
    from ragroute_b200.router import FusedCorpusRouter
    r = FusedCorpusRouter(c.in_dim)
    r.load_state_dict({k: torch.from_numpy(v) for k, v in c.sd.items()})
    r.eval()
    r.set_tables(c.data_sources, c.d_max, c.model_of_source, c.centroids, c.onehot_ids, c.n_onehot, c.scaler)
    return r


@pytest.mark.parametrize("dataset", ["medrag", "feb4rag", "wikipedia"])
def test_router_matches_reference_golden(cuda_required, dataset):
    c = RouterCase(dataset)
    r = _make_router(c)
    assert r.fc1.in_features == c.in_dim
    thr_logit = router_ref.logit_threshold(c.threshold)
    # batched select over all golden queries at once
    slots = r.slot_names
    q = np.zeros((c.n_queries, len(slots), c.d_max), dtype=np.float32)
    for qi in range(c.n_queries):
        emb = c.query(qi)
        for si, m in enumerate(slots):
            q[qi, si, :len(emb[m])] = emb[m]
    mask, logits = r.select(torch.from_numpy(q).cuda(), c.threshold, return_logits=True)
    torch.cuda.synchronize()
    mask, logits = mask.cpu().numpy().astype(bool), logits.cpu().numpy()
    np.testing.assert_allclose(logits, c.ref_logits, rtol=0, atol=1e-5)
    decided = np.abs(c.ref_logits - thr_logit) > 1e-5
    assert np.array_equal(mask[decided], c.ref_masks[decided])
    # one query through the reference's own signature
    names = r.select_relevant_sources(c.query(0), c.threshold)
    ref_names = [s for s, m in zip(c.data_sources, c.ref_masks[0]) if m]
    undecided = {s for s, dcd in zip(c.data_sources, decided[0]) if not dcd}
    assert set(names) ^ set(ref_names) <= undecided
    assert names == [s for s in c.data_sources if s in names]      # data_sources order
    # module call on explicit feature rows == oracle features -> same logits
    feats = router_ref.build_features(c.query(3), c.data_sources, c.model_of_source, c.centroids,
                                      c.onehot_ids, c.n_onehot, c.d_max)
    if c.scaler:
        feats = (feats - c.scaler[0]) / c.scaler[1]
    out = r(torch.tensor(feats, dtype=torch.float32)).cpu().numpy()
    assert out.shape == (len(c.data_sources), 1)
    np.testing.assert_allclose(out[:, 0], c.ref_logits[3], rtol=0, atol=1e-5)


def test_router_batch_sizes_and_errors(cuda_required):
    c = RouterCase("medrag")
    r = _make_router(c)
    rng = np.random.default_rng(0)
    for B in (1, 3, 4, 5, 64, 257):
        q = rng.uniform(-0.4, 0.4, size=(B, c.d_max)).astype(np.float32)
        mask, logits = r.select(torch.from_numpy(q).cuda(), c.threshold, return_logits=True)
        logits = logits.cpu().numpy()
        for b in (0, B - 1):
            feats = router_ref.build_features({c.models[0]: q[b]}, c.data_sources, c.model_of_source,
                                              c.centroids, c.onehot_ids, c.n_onehot, c.d_max)
            m_ref, l_ref = router_ref.select(c.sd, feats, c.threshold, *c.scaler)
            np.testing.assert_allclose(logits[b], l_ref, rtol=0, atol=1e-5)
    from ragroute_b200.router import FusedCorpusRouter
    from ragroute_b200._lib import RRError
    with pytest.raises(RRError):
        FusedCorpusRouter(10)(torch.zeros(1, 10))      # used before load_state_dict
    with pytest.raises(KeyError):
        FusedCorpusRouter(10).load_state_dict({})


def test_router_logits_do_not_depend_on_the_batch(cuda_required):
    """fc1 is summed in one fixed tree whether a cluster of CTAs (small batches) or single CTAs (large ones)
    compute it: a query's logits and mask are bit-identical alone, in a batch of 100 and in a batch of 300."""
    c = RouterCase("medrag")
    r = _make_router(c)
    rng = np.random.default_rng(4)
    q = torch.from_numpy(rng.uniform(-0.4, 0.4, size=(300, c.d_max)).astype(np.float32)).cuda()
    m300, l300 = r.select(q, c.threshold, return_logits=True)
    m100, l100 = r.select(q[:100].contiguous(), c.threshold, return_logits=True)
    assert torch.equal(l300[:100], l100) and torch.equal(m300[:100], m100)
    for b in (0, 57, 99):
        m1, l1 = r.select(q[b:b + 1].contiguous(), c.threshold, return_logits=True)
        assert torch.equal(l1[0], l300[b]) and torch.equal(m1[0], m300[b])


def test_rerank_matches_reference_golden(cuda_required):
    from ragroute_b200 import rerank
    g = np.load(os.path.join(GOLDEN, "rerank.npz"))
    for ci in range(int(g["n_cases"])):
        scores = g[f"c{ci}_scores"]
        k = int(g[f"c{ci}_k"])
        docs = [f"doc{i}" for i in range(len(scores))]
        s_list = [float(v) for v in scores]
        d1, s1 = rerank.rerank_medrag(docs, s_list, k)
        d2, s2 = rerank.rerank_wikipedia(docs, s_list, k)
        assert d1 == [f"doc{i}" for i in g[f"c{ci}_medrag_docs"]] and s1 == g[f"c{ci}_medrag_scores"].tolist()
        assert d2 == [f"doc{i}" for i in g[f"c{ci}_wikipedia_docs"]] and s2 == g[f"c{ci}_wikipedia_scores"].tolist()


@pytest.mark.parametrize("B,L,k", [(1, 128, 32), (7, 96, 32), (33, 1024, 32), (5, 40, 10), (3, 4096, 32), (2, 5, 8)])
@pytest.mark.parametrize("descending", [True, False])
def test_merge_topk_vs_oracle_with_ties_and_holes(cuda_required, B, L, k, descending):
    from ragroute_b200.rerank import merge_topk
    rng = np.random.default_rng(B * 1000 + L)
    scores = rng.integers(0, 50, size=(B, L)).astype(np.float32) / 4     # many exact ties
    ids = rng.integers(0, 1 << 40, size=(B, L)).astype(np.int64)
    ids[rng.random((B, L)) < 0.2] = -1                                   # empty slots (faiss padding)
    scores[ids < 0] = -np.finfo(np.float32).max
    s, i, p = merge_topk(torch.from_numpy(scores).cuda(), torch.from_numpy(ids).cuda(), k, descending)
    rs, ri, rp = rerank_ref.merge_topk(scores, ids, k, descending)
    assert np.array_equal(s.cpu().numpy(), rs) and np.array_equal(i.cpu().numpy(), ri)
    assert np.array_equal(p.cpu().numpy(), rp)


@pytest.mark.parametrize("B,W,S,k", [(1, 1, 4, 32), (9, 2, 4, 32), (64, 8, 4, 32), (5, 3, 13, 10), (130, 4, 1, 32)])
@pytest.mark.parametrize("descending", [True, False])
def test_merge_segments_in_place_vs_oracle(cuda_required, B, W, S, k, descending):
    """rr_merge_topk_segments over W gathered blocks (the all-gather's receive buffer, read in place)
    equals the oracle's merge over their concatenation; source / rank are decoded on the device."""
    from ragroute_b200.federation import PackedLists, merge_packed, decode_pos
    rng = np.random.default_rng(B * 100 + W * 10 + S)
    L = S * k
    scores = rng.integers(0, 60, size=(W, B, L)).astype(np.float32) / 4   # many exact ties
    ids = rng.integers(0, 1 << 40, size=(W, B, L)).astype(np.int64)
    ids[rng.random((W, B, L)) < 0.3] = -1
    scores[ids < 0] = -np.finfo(np.float32).max if descending else np.finfo(np.float32).max
    pl = PackedLists(B, L, "cuda", n_blocks=W)
    pl.buf.fill_(0xAB)
    for r in range(W):
        pl.scores(r).copy_(torch.from_numpy(scores[r]))
        pl.ids(r).copy_(torch.from_numpy(ids[r]))
    s, i, pos, src, rank = merge_packed(pl, k, k, descending)
    cat_s = np.concatenate(list(scores), axis=1)
    cat_i = np.concatenate(list(ids), axis=1)
    rs, ri, rp = rerank_ref.merge_topk(cat_s, cat_i, k, descending, seg_len=L, list_len=k)
    assert np.array_equal(s.cpu().numpy(), rs) and np.array_equal(i.cpu().numpy(), ri)
    assert np.array_equal(pos.cpu().numpy(), rp)
    r_ref, s_ref = decode_pos(torch.from_numpy(rp), S, k)
    assert np.array_equal(rank.cpu().numpy(), r_ref.numpy()) and np.array_equal(src.cpu().numpy(), s_ref.numpy())


@pytest.mark.parametrize("W", [2, 3, 8])
def test_sharded_merge_equals_unsharded_with_cross_rank_ties(cuda_required, W):
    """equal scores from different sources / ranks must come back in the same order whether
    the corpus is one shard or W (duplicated documents across corpora produce exactly such ties).  Per
    (query, source) a pool of scored rows with many equal scores; the unsharded lists are each source's
    top-k by (score desc, row asc), the sharded ones each (rank, source)'s top-k of its contiguous rows."""
    from ragroute_b200.federation import PackedLists, merge_packed
    rng = np.random.default_rng(W)
    B, S, k, n = 40, 3, 8, 60
    per = -(-n // W)
    whole = PackedLists(B, S * k, "cuda", n_blocks=1)
    shard = PackedLists(B, S * k, "cuda", n_blocks=W)
    fmax = np.finfo(np.float32).max
    ws, wi = np.full((B, S * k), -fmax, np.float32), np.full((B, S * k), -1, np.int64)
    ss, si = np.full((W, B, S * k), -fmax, np.float32), np.full((W, B, S * k), -1, np.int64)
    for b in range(B):
        for s in range(S):
            sc = rng.integers(0, 6, size=n).astype(np.float32)             # 6 distinct scores over 60 rows
            order = np.lexsort((np.arange(n), -sc))
            top = order[:k]
            ws[b, s * k:(s + 1) * k], wi[b, s * k:(s + 1) * k] = sc[top], top
            for r in range(W):
                rows = np.arange(r * per, min(n, (r + 1) * per))
                o = rows[np.lexsort((rows, -sc[rows]))][:k]
                ss[r, b, s * k:s * k + len(o)], si[r, b, s * k:s * k + len(o)] = sc[o], o
    whole.scores().copy_(torch.from_numpy(ws)); whole.ids().copy_(torch.from_numpy(wi))
    for r in range(W):
        shard.scores(r).copy_(torch.from_numpy(ss[r])); shard.ids(r).copy_(torch.from_numpy(si[r]))
    a = merge_packed(whole, k, k)
    c = merge_packed(shard, k, k)
    torch.cuda.synchronize()
    assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1]) and torch.equal(a[3], c[3]), \
        "sharded merge must equal the unsharded one bit for bit, ties included"
    rs, ri, _ = rerank_ref.merge_topk(ws, wi, k, True, seg_len=S * k, list_len=k)
    assert np.array_equal(a[0].cpu().numpy(), rs) and np.array_equal(a[1].cpu().numpy(), ri)


def test_merge_segments_rejects_bad_arguments(cuda_required):
    from ragroute_b200 import _lib
    lib = _lib.load()
    z = torch.zeros(4096, device="cuda")
    zi = torch.zeros(4096, dtype=torch.int64, device="cuda")
    args = lambda n_seg, seg_len, ss, si, k, ll: (z.data_ptr(), zi.data_ptr(), 2, n_seg, seg_len, ss, si, k, 1, ll,
                                                  z.data_ptr(), zi.data_ptr(), None, None, None, None)
    assert lib.rr_merge_topk_segments(*args(2, 64, 128, 128, 32, 32)) == 0
    assert lib.rr_merge_topk_segments(*args(2, 64, 100, 128, 32, 32)) == -1     # overlapping segments
    assert lib.rr_merge_topk_segments(*args(2, 64, 128, 128, 32, 48)) == -1     # seg_len % list_len
    assert lib.rr_merge_topk_segments(*args(80, 64, 128, 128, 32, 32)) == -1    # > RR_MERGE_MAX_L
    assert lib.rr_merge_topk_segments(*args(0, 64, 128, 128, 32, 32)) == -1
    torch.cuda.synchronize()


def test_load_router_from_reference_style_artifacts(cuda_required, tmp_path):
    """best_model.pth + preprocessed_data.pkl + *_stats.json, laid out as Router.load_router reads them
    (router.py:106-151), loaded through ragroute_b200.artifacts and checked against the oracle."""
    import json
    import pickle
    import torch
    from sklearn.preprocessing import StandardScaler
    from oracle import router_ref
    from ragroute_b200 import artifacts, synthetic as syn
    names = ["pubmed", "statpearls", "textbooks", "wikipedia"]
    d, in_dim = 768, 1540
    sd = syn.router_state_dict(in_dim, 5)
    torch.save({k: torch.from_numpy(v) for k, v in sd.items()}, tmp_path / "best_model.pth")
    rng = np.random.default_rng(2)
    scaler = StandardScaler().fit(rng.standard_normal((64, in_dim)) * 0.1)
    with open(tmp_path / "preprocessed_data.pkl", "wb") as f:
        pickle.dump((None, None, None, scaler, None), f)
    files, centroids = {}, {}
    for n in names:
        centroids[n] = syn.source_mean(syn.MEDRAG_SOURCE_TO_ID[n], d)
        files[n] = str(tmp_path / f"{n}_stats.json")
        with open(files[n], "w") as f:
            json.dump({"centroid": centroids[n].tolist()}, f)
    model = {n: "ncbi/MedCPT-Query-Encoder" for n in names}
    router, thr = artifacts.load_router("medrag", names, model, str(tmp_path / "best_model.pth"), files,
                                        str(tmp_path / "preprocessed_data.pkl"))
    assert thr == 0.4924 and router.fc1.in_features == in_dim
    q = syn.queries(9, d, [0, 1, 2, 3], seed=3)
    thr_logit = router_ref.logit_threshold(thr)
    for b in range(9):
        picked = router.select_relevant_sources({"ncbi/MedCPT-Query-Encoder": q[b]}, thr)
        feats = router_ref.build_features({"m": q[b]}, names, {n: "m" for n in names}, centroids,
                                          syn.MEDRAG_SOURCE_TO_ID, 4, d)
        mask, logits = router_ref.select(sd, feats, thr, scaler.mean_, scaler.scale_)
        sure = np.abs(logits - thr_logit) > 1e-5
        want = [n for n, m, s in zip(names, mask, sure) if m and s]
        maybe = [n for n, s in zip(names, sure) if not s]
        assert all(n in picked for n in want) and all(n in want or n in maybe for n in picked)
    with pytest.raises(ValueError):
        artifacts.load_router("bogus", names, model, str(tmp_path / "best_model.pth"), files)

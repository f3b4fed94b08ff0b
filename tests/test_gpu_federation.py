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


# FeB4RAG shape (config.py:34,44-58,72-86,94,99): 13 sources, 8 query encoders of 768 / 1024 / 4096
# dims, EMBEDDING_MAX_LENGTH 4096, router in_dim 2*4096 + 13 = 8205, no scaler, threshold 0.5, k = 10.
_FEB_SOURCES = ["msmarco", "trec-covid", "nfcorpus", "scidocs", "nq", "hotpotqa", "fiqa", "arguana",
                "webis-touche2020", "dbpedia-entity", "fever", "climate-fever", "scifact"]
_FEB_MODEL = {"msmarco": "e5-large", "trec-covid": "SGPT-5.8B", "nfcorpus": "UAE-Large-V1",
              "scidocs": "all-mpnet-base-v2", "nq": "multilingual-e5-large", "hotpotqa": "ember-v1",
              "fiqa": "all-mpnet-base-v2", "arguana": "UAE-Large-V1", "webis-touche2020": "e5-base",
              "dbpedia-entity": "UAE-Large-V1", "fever": "UAE-Large-V1", "climate-fever": "UAE-Large-V1",
              "scifact": "gte-base"}
_FEB_DIM = {"e5-large": 1024, "SGPT-5.8B": 4096, "UAE-Large-V1": 1024, "all-mpnet-base-v2": 768,
            "multilingual-e5-large": 1024, "ember-v1": 1024, "e5-base": 768, "gte-base": 768}
_FEB_ID = {"arguana": 0, "climate-fever": 1, "dbpedia-entity": 2, "fever": 3, "fiqa": 4, "hotpotqa": 5,
           "msmarco": 6, "nfcorpus": 7, "nq": 8, "scidocs": 9, "scifact": 10, "trec-covid": 11,
           "webis-touche2020": 12}


def test_feb4rag_shaped_federation(cuda_required):
    """Mixed-dimension sources (768 / 1024 / 4096: the 4096-d one takes the swap-AB scan in
    16-query passes), several encoders, no scaler: router mask and per-source top-10 against the oracle."""
    from oracle import flat_ip, router_ref
    from ragroute_b200 import synthetic as syn
    from ragroute_b200.faiss_compat import IndexFlatIP
    from ragroute_b200.federation import Federation
    from ragroute_b200.router import FusedCorpusRouter
    d_max, k, B, S = 4096, 10, 70, 13
    rng = np.random.default_rng(5)
    corpora, centroids = {}, {}
    for i, n in enumerate(_FEB_SOURCES):
        d = _FEB_DIM[_FEB_MODEL[n]]
        mu = rng.standard_normal(d).astype(np.float32) * np.float32(0.5 / np.sqrt(d))
        corpora[n] = (rng.standard_normal((700 + 90 * i, d), dtype=np.float32) / np.float32(np.sqrt(d)) + mu)
        centroids[n] = np.pad(corpora[n].mean(axis=0), (0, d_max - d)).astype(np.float32)   # router.py:149-151
    models = sorted(set(_FEB_MODEL.values()))
    q_by_model = {m: (rng.standard_normal((B, _FEB_DIM[m]), dtype=np.float32) / np.float32(np.sqrt(_FEB_DIM[m])))
                  for m in models}
    in_dim = 2 * d_max + S
    sd = syn.router_state_dict(in_dim, 3)
    router = FusedCorpusRouter(in_dim)
    router.load_state_dict(sd)
    router.set_tables(_FEB_SOURCES, d_max, _FEB_MODEL, centroids, _FEB_ID, S, scaler=None)
    # calibrate fc3.bias so that about half the pairs are routed
    q_router = np.zeros((B, len(router.slot_names), d_max), dtype=np.float32)
    for si, m in enumerate(router.slot_names):
        q_router[:, si, :_FEB_DIM[m]] = q_by_model[m]
    qr = torch.from_numpy(q_router).cuda()
    _, lg = router.select(qr, 0.5, return_logits=True)
    sd = dict(sd)
    sd["fc3.bias"] = (sd["fc3.bias"] - np.float32(lg.median().item())).astype(np.float32)
    router.load_state_dict(sd)
    router.set_tables(_FEB_SOURCES, d_max, _FEB_MODEL, centroids, _FEB_ID, S, scaler=None)

    sources = {}
    for n in _FEB_SOURCES:
        ix = IndexFlatIP(corpora[n].shape[1])
        ix.add(corpora[n])
        sources[n] = ix
    fed = Federation(sources, k, router=router, routing="ragroute", threshold=0.5, dataset="feb4rag")
    mask = fed.route(qr)
    q_list = [torch.from_numpy(q_by_model[_FEB_MODEL[n]]).cuda() for n in _FEB_SOURCES]
    D_loc, I_loc = fed.search_sources(q_list, mask)
    torch.cuda.synchronize()
    mask, D_loc, I_loc = mask.cpu().numpy().astype(bool), D_loc.cpu().numpy(), I_loc.cpu().numpy()
    assert 0.2 < mask.mean() < 0.8
    for b in range(0, B, 7):
        feats = router_ref.build_features({m: q_by_model[m][b] for m in models}, _FEB_SOURCES, _FEB_MODEL,
                                          centroids, _FEB_ID, S, d_max)
        ref_mask, ref_logits = router_ref.select(sd, feats, 0.5)
        decided = np.abs(ref_logits) > 1e-5
        assert np.array_equal(ref_mask[decided], mask[b][decided])
        for s, n in enumerate(_FEB_SOURCES):
            Ds, Is = D_loc[b:b + 1, s * k:(s + 1) * k], I_loc[b:b + 1, s * k:(s + 1) * k]
            if not mask[b, s]:
                assert (Is == -1).all()
                continue
            qv = q_by_model[_FEB_MODEL[n]][b:b + 1]
            Dr, Ir = flat_ip.search_seq(qv, corpora[n], k)
            flat_ip.compare_topk(Ds, Is, Dr, Ir, x=qv, xb=corpora[n])


def test_routed_large_batch_federation_matches_single_queries(cuda_required):
    """A routed batch of 300 (large-batch kernels, compaction of the routed queries) must return,
    query by query, exactly what each query returns alone (swap-AB kernel)."""
    from ragroute_b200 import synthetic as syn
    from ragroute_b200.selftest import build_small_medrag
    from ragroute_b200.faiss_compat import IndexFlatIP
    from ragroute_b200.federation import Federation
    from ragroute_b200.router import MEDRAG_THRESHOLD
    from ragroute_b200.workloads import make_synthetic_router
    d, k, B = 768, 32, 300
    names, sids, corpora, _, _, centroids = build_small_medrag(n_rows=(20000, 30000), d=d)
    router = make_synthetic_router(names, d, centroids, 11, 0)
    sources = {}
    for n in names:
        ix = IndexFlatIP(d)
        ix.add(corpora[n])
        sources[n] = ix
    fed = Federation(sources, k, router=router, routing="ragroute", threshold=MEDRAG_THRESHOLD)
    q = torch.from_numpy(syn.queries(B, d, sids, seed=77)).cuda()
    out = fed.search(q)
    torch.cuda.synchronize()
    assert all(sources[n].stats()["exact_rescans"] == 0 for n in names)
    for b in range(0, B, 23):
        one = fed.search(q[b:b + 1].contiguous())
        for key in ("scores", "ids", "source", "mask"):
            assert torch.equal(one[key][0], out[key][b]), f"query {b}: {key} differs between batch and single"


def test_row_shards_merge_to_the_unsharded_result(cuda_required):
    """Top-k is decomposable: searching W row shards (ids offset by the shard start) and
    merging their lists equals the unsharded search -- here with all shards on one GPU, which is the
    per-rank computation of the multi-GPU path without the collective (covered on CPU by the gloo test)."""
    from ragroute_b200 import synthetic as syn
    from ragroute_b200.faiss_compat import IndexFlatIP
    from ragroute_b200.federation import shard_rows
    from ragroute_b200.rerank import merge_topk
    d, k, W = 768, 32, 3
    xb = syn.corpus_rows(0, 50001, d)
    for B in (5, 130):
        q = torch.from_numpy(syn.queries(B, d, [0], seed=B)).cuda()
        whole = IndexFlatIP(d)
        whole.add(xb)
        D0, I0 = whole.search_tensor(q, k)
        Ds, Is = [], []
        for r in range(W):
            lo, hi = shard_rows(len(xb), W, r)
            ix = IndexFlatIP(d)
            ix.add(xb[lo:hi])
            ix.set_id_offset(lo)
            Dr, Ir = ix.search_tensor(q, k)
            Ds.append(Dr)
            Is.append(Ir)
        Dm, Im, _ = merge_topk(torch.cat(Ds, dim=1).contiguous(), torch.cat(Is, dim=1).contiguous(), k)
        torch.cuda.synchronize()
        assert torch.equal(Im, I0) and torch.equal(Dm, D0)


def test_micro_batcher_over_a_real_index(cuda_required):
    """40 concurrent single-query requests (what --parallel 40 sends, run_benchmark.py:90-114) through the
    collector equal 40 separate index.search calls, and really were batched."""
    import asyncio
    from ragroute_b200 import synthetic as syn
    from ragroute_b200.collector import MicroBatcher, search_rows
    from ragroute_b200.faiss_compat import IndexFlatIP
    d, k = 768, 32
    xb = syn.corpus_rows(1, 20000, d)
    ix = IndexFlatIP(d)
    ix.add(xb)
    qs = syn.queries(40, d, [1], seed=8)

    async def main():
        mb = MicroBatcher(lambda embs: search_rows(ix, embs, k), max_batch=64, max_wait_s=0.02)
        res = await asyncio.gather(*[mb.submit(qs[i].reshape(1, -1)) for i in range(40)])
        await mb.close()
        return res, mb

    res, mb = asyncio.run(main())
    assert mb.batches < 40 and mb.largest_batch > 1
    for i, (D, I) in enumerate(res):
        D1, I1 = ix.search(qs[i:i + 1], k)
        assert np.array_equal(I, I1) and np.array_equal(D, D1)


@pytest.mark.parametrize("B,density", [(1, 0.6), (7, 0.6), (64, 0.6), (100, 0.6), (300, 0.6), (300, 0.25), (2100, 0.6),
                                       (2100, 0.04)])
def test_fused_fan_out_equals_per_source_calls(cuda_required, B, density):
    """rr_federation_search (one launch per stage for all sources) returns bit for bit what one
    rr_index_search_ex per source returns, on every dispatch path (swap-AB scan, queries-in-TMEM scan,
    CTA pairs, two chunks), with a routing mask, an empty source, an exact-mode source and a source
    whose near-duplicate rows force the second pass."""
    from ragroute_b200 import synthetic as syn
    from ragroute_b200.faiss_compat import IndexFlatIP
    from ragroute_b200.federation import Federation
    d, k = 768, 32
    rng = np.random.default_rng(B)
    corpora = [syn.corpus_rows(0, 9000, d), syn.corpus_rows(1, 4100, d), np.zeros((0, d), np.float32),
               syn.corpus_rows(2, 2500, d)]
    base = syn.corpus_rows(3, 12, d)
    dup = np.repeat(base, 200, axis=0) * (1 + 1e-4 * rng.standard_normal((2400, 1)).astype(np.float32))
    corpora.append(dup.astype(np.float32))                       # 200 near-copies of each row: certificates fail
    sources = {}
    for i, xb in enumerate(corpora):
        ix = IndexFlatIP(d)
        if len(xb):
            ix.add(xb)
        ix.set_id_offset(1000 * i)
        if i == 3:
            ix.set_mode("exact")
        sources[f"s{i}"] = ix
    q = torch.from_numpy(syn.queries(B, d, [0, 1, 2, 3], seed=100 + B)).cuda()
    # the density decides, on the device, between single CTAs (<= 128 routed queries) and CTA pairs
    mask = torch.from_numpy((rng.random((B, len(corpora))) < density).astype(np.uint8)).cuda()
    res = {}
    for fused in (True, False):
        fed = Federation(sources, k, routing="all", fused=fused)
        D, I = fed.search_sources(q, mask)
        torch.cuda.synchronize()
        res[fused] = (D.clone(), I.clone())
    assert torch.equal(res[True][1], res[False][1])
    assert torch.equal(res[True][0], res[False][0])
    m = mask.cpu().numpy().astype(bool)
    I = res[True][1].cpu().numpy().reshape(B, len(corpora), k)
    assert (I[~m] == -1).all() and (I[:, 2] == -1).all()
    assert (I[m[:, 0], 0] >= 0).all() and (I[m[:, 4], 4] >= 4000).all()
    if B >= 7 and density > 0.5:
        assert sources["s4"].stats()["exact_rescans"] > 0          # the second pass really ran
    # unmasked call (mask NULL) through the same entry point
    fed = Federation(sources, k, routing="all", fused=True)
    D1, I1 = fed.search_sources(q, None)
    fed2 = Federation(sources, k, routing="all", fused=False)
    D2, I2 = fed2.search_sources(q, None)
    torch.cuda.synchronize()
    assert torch.equal(I1, I2) and torch.equal(D1, D2)


def test_fused_fan_out_rejects_bad_arguments(cuda_required):
    from ragroute_b200 import _lib
    from ragroute_b200.faiss_compat import IndexFlatIP
    lib = _lib.load()
    ix = IndexFlatIP(64)
    ix.add(np.ones((10, 64), np.float32))
    q = torch.ones((2, 64), device="cuda")
    D = torch.empty((2, 64), device="cuda")
    I = torch.empty((2, 64), dtype=torch.int64, device="cuda")
    hs = (_lib.c_void_p * 2)(ix._h, ix._h)
    qs = (_lib.c_void_p * 2)(q.data_ptr(), q.data_ptr())
    rc = lib.rr_federation_search(hs, 2, qs, 2, 32, None, 0, D.data_ptr(), I.data_ptr(), 0, None)
    assert rc == -1 and b"twice" in lib.rr_last_error()
    rc = lib.rr_federation_search(hs, 1, qs, 2, 33, None, 0, D.data_ptr(), I.data_ptr(), 0, None)
    assert rc == -4
    rc = lib.rr_federation_search(hs, 1, qs, 2, 32, None, 0, D.data_ptr(), I.data_ptr(), 16, None)
    assert rc == -1
    rc = lib.rr_federation_search(hs, 0, qs, 2, 32, None, 0, D.data_ptr(), I.data_ptr(), 0, None)
    assert rc == -1


def test_config5_sixteen_mixed_dim_sources_placed_on_two_ranks(cuda_required):
    """BASELINE config 5: 16 sources, dims cycling 384 / 768 / 1024, k = 10, no scaler,
    in_dim = 2*1024 + 16, 64 queries in flight (--parallel 64).  (a) routed per-source top-10 and the
    router mask against the oracle; (b) whole-source placement over 2 ranks (each rank holds every other
    source and an empty index for the rest), the ranks' packed lists merged as the all-gather would
    deliver them, equals the single-process result."""
    from oracle import flat_ip, router_ref
    from ragroute_b200 import synthetic as syn
    from ragroute_b200.faiss_compat import IndexFlatIP
    from ragroute_b200.federation import Federation, PackedLists, merge_packed, place_sources
    from ragroute_b200.router import FusedCorpusRouter
    S, k, B, d_max = 16, 10, 64, 1024
    dims = [(384, 768, 1024)[i % 3] for i in range(S)]
    names = [f"src{i:02d}" for i in range(S)]
    model = {n: f"enc{dims[i]}" for i, n in enumerate(names)}
    rng = np.random.default_rng(55)
    n_rows = np.exp(rng.uniform(np.log(500), np.log(20000), size=S)).astype(int)   # log-uniform sizes, scaled down
    corpora, centroids = {}, {}
    for i, n in enumerate(names):
        d = dims[i]
        mu = rng.standard_normal(d).astype(np.float32) * np.float32(0.5 / np.sqrt(d))
        corpora[n] = rng.standard_normal((n_rows[i], d), dtype=np.float32) / np.float32(np.sqrt(d)) + mu
        centroids[n] = np.pad(corpora[n].mean(axis=0), (0, d_max - d)).astype(np.float32)
    q_by_model = {f"enc{d}": rng.standard_normal((B, d), dtype=np.float32) / np.float32(np.sqrt(d))
                  for d in (384, 768, 1024)}
    in_dim = 2 * d_max + S
    sd = syn.router_state_dict(in_dim, 9)
    onehot = {n: i for i, n in enumerate(names)}

    def make_router(sd_):
        r = FusedCorpusRouter(in_dim)
        r.load_state_dict(sd_)
        r.set_tables(names, d_max, model, centroids, onehot, S, scaler=None)
        return r

    router = make_router(sd)
    q_router = np.zeros((B, len(router.slot_names), d_max), dtype=np.float32)
    for si, m in enumerate(router.slot_names):
        q_router[:, si, :q_by_model[m].shape[1]] = q_by_model[m]
    qr = torch.from_numpy(q_router).cuda()
    _, lg = router.select(qr, 0.5, return_logits=True)
    sd = dict(sd)
    sd["fc3.bias"] = (sd["fc3.bias"] - np.float32(lg.median().item())).astype(np.float32)
    router = make_router(sd)

    def build(keep):
        out = {}
        for i, n in enumerate(names):
            ix = IndexFlatIP(dims[i])
            if keep(i):
                ix.add(corpora[n])
            out[n] = ix
        return out

    q_list = [torch.from_numpy(q_by_model[model[n]]).cuda() for n in names]
    fed = Federation(build(lambda i: True), k, router=router, routing="ragroute", threshold=0.5, dataset="feb4rag")
    mask = fed.route(qr)
    D_loc, I_loc = fed.search_sources(q_list, mask)
    full = fed.merge(D_loc, I_loc, mask)
    torch.cuda.synchronize()
    m, Dn, In = mask.cpu().numpy().astype(bool), D_loc.cpu().numpy(), I_loc.cpu().numpy()
    assert 0.2 < m.mean() < 0.8
    for b in range(0, B, 9):
        feats = router_ref.build_features({mm: q_by_model[mm][b] for mm in q_by_model}, names, model, centroids,
                                          onehot, S, d_max)
        ref_mask, ref_logits = router_ref.select(sd, feats, 0.5)
        decided = np.abs(ref_logits) > 1e-5
        assert np.array_equal(ref_mask[decided], m[b][decided])
        for s, n in enumerate(names):
            Ds, Is = Dn[b:b + 1, s * k:(s + 1) * k], In[b:b + 1, s * k:(s + 1) * k]
            if not m[b, s]:
                assert (Is == -1).all()
                continue
            qv = q_by_model[model[n]][b:b + 1]
            Dr, Ir = flat_ip.search_seq(qv, corpora[n], k)
            flat_ip.compare_topk(Ds, Is, Dr, Ir, x=qv, xb=corpora[n])
    # (b) two ranks, whole-source placement
    where = place_sources(names, 2)
    gathered = PackedLists(B, S * k, "cuda", n_blocks=2)
    for r in range(2):
        fr = Federation(build(lambda i: where[names[i]] == r), k, router=router, routing="ragroute", threshold=0.5,
                        dataset="feb4rag")
        Dr_, Ir_ = fr.search_sources(q_list, mask)
        gathered.scores(r).copy_(Dr_)
        gathered.ids(r).copy_(Ir_)
    s2, i2, _, src2, rank2 = merge_packed(gathered, k, k)
    torch.cuda.synchronize()
    assert torch.equal(s2, full["scores"]) and torch.equal(i2, full["ids"]) and torch.equal(src2, full["source"])
    valid = (src2 >= 0).cpu().numpy()
    want_rank = np.vectorize(lambda s: where[names[s]] if s >= 0 else -1)(src2.cpu().numpy())
    assert np.array_equal(rank2.cpu().numpy()[valid], want_rank[valid])


def test_search_host_equals_device_search(cuda_required):
    """The host-buffer entry point (pinned queries in, one packed copy out) returns what the device path returns."""
    from ragroute_b200 import synthetic as syn
    from ragroute_b200.selftest import build_small_medrag
    from ragroute_b200.faiss_compat import IndexFlatIP
    from ragroute_b200.federation import Federation
    from ragroute_b200.router import MEDRAG_THRESHOLD
    from ragroute_b200.workloads import make_synthetic_router
    d, k = 768, 32
    names, sids, corpora, _, _, centroids = build_small_medrag(n_rows=(6000, 9000), d=d)
    router = make_synthetic_router(names, d, centroids, 3, 0)
    sources = {}
    for n in names:
        ix = IndexFlatIP(d)
        ix.add(corpora[n])
        sources[n] = ix
    fed = Federation(sources, k, router=router, routing="ragroute", threshold=MEDRAG_THRESHOLD)
    for B in (1, 33, 33, 200):                                       # 33 twice: the pinned buffer is reused
        qh = torch.from_numpy(syn.queries(B, d, sids, seed=B)).pin_memory()
        s_h, i_h, src_h = fed.search_host(qh)
        s_h, i_h, src_h = s_h.clone(), i_h.clone(), src_h.clone()
        dev = fed.search(qh.cuda())
        torch.cuda.synchronize()
        assert torch.equal(s_h, dev["scores"].cpu()) and torch.equal(i_h, dev["ids"].cpu())
        assert torch.equal(src_h, dev["source"].cpu())
        assert s_h.shape == (B, k) and i_h.dtype == torch.int64 and src_h.dtype == torch.int32

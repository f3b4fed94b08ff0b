"""Parity of the CUDA flat-IP index (through the C ABI) with the CPU oracle.

Tolerances (north_star): ids equal to the oracle's FAISS-fp32 restatement except
ids whose float64 score is within 1e-5 relative of the k-th score (documented
ties); scores within 1e-5 relative.  'auto' (tcgen05 bf16 scan + exact rescoring +
certificate) must return bit-identical results to 'exact'."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import flat_ip as oracle
from ragroute_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
FLT_MAX = np.finfo(np.float32).max


@pytest.fixture(scope="module")
def faiss(cuda_required):
    import ragroute_b200.faiss_compat as faiss
    return faiss


def _index(faiss, xb, mode="auto"):
    ix = faiss.IndexFlatIP(xb.shape[1])
    ix.add(xb)
    ix.set_mode(mode)
    return ix


def _check_vs_oracle(faiss, xb, q, k, mode, expect_no_rescan=False):
    ix = _index(faiss, xb, mode)
    D, I = ix.search(q, k)
    Dr, Ir = oracle.search(q, xb, k)
    info = oracle.compare_topk(D, I, Dr, Ir, x=q, xb=xb)
    st = ix.stats()
    if expect_no_rescan:
        assert st["exact_rescans"] == 0, f"certificate failed on benign data: {st}"
    return D, I, info, st


@pytest.mark.parametrize("mode", ["exact", "auto"])
def test_hand_computable_duplicates(faiss, mode):
    xb = np.zeros((64, 8), dtype=np.float32)
    for j in range(64):
        xb[j, j % 8] = j // 8 + 1
    xb[10] = xb[58]
    q = np.zeros((1, 8), dtype=np.float32)
    q[0, 2] = 2.0
    D, I = _index(faiss, xb, mode).search(q, 5)
    assert D[0].tolist() == [16.0, 16.0, 14.0, 12.0, 10.0]
    assert I[0].tolist() == [10, 58, 50, 42, 34]   # ties: lower row id first (deterministic)


@pytest.mark.parametrize("mode", ["exact", "auto"])
def test_fewer_rows_than_k_and_empty(faiss, mode):
    rng = np.random.default_rng(1)
    xb = rng.standard_normal((5, 16), dtype=np.float32)
    q = rng.standard_normal((3, 16), dtype=np.float32)
    D, I = _index(faiss, xb, mode).search(q, 8)
    Dr, Ir = oracle.search_seq(q, xb, 8)
    oracle.compare_topk(D, I, Dr, Ir, x=q, xb=xb)
    assert (I[:, 5:] == -1).all() and (D[:, 5:] == -FLT_MAX).all()
    ix = faiss.IndexFlatIP(16)
    ix.set_mode(mode)
    D0, I0 = ix.search(q, 4)
    assert (I0 == -1).all() and (D0 == -FLT_MAX).all()


@pytest.mark.parametrize("mode", ["exact", "auto"])
@pytest.mark.parametrize("n,d,k,nq", [(5000, 768, 32, 5), (3000, 1024, 10, 3), (4000, 384, 10, 7),
                                      (2500, 100, 7, 2), (700, 4096, 10, 2), (129, 768, 32, 1)])
def test_random_gaussian_vs_oracle(faiss, mode, n, d, k, nq):
    rng = np.random.default_rng(n + d)
    xb = rng.standard_normal((n, d), dtype=np.float32)
    q = rng.standard_normal((nq, d), dtype=np.float32)
    _check_vs_oracle(faiss, xb, q, k, mode)


def test_medrag_shaped_corpus_auto_equals_exact_without_rescans(faiss):
    d, k = 768, 32
    xb = syn.corpus_rows(2, 60000, d)
    q = syn.queries(64, d, [2], seed=3)
    Da, Ia, info, st = _check_vs_oracle(faiss, xb, q, k, "auto", expect_no_rescan=True)
    De, Ie = _index(faiss, xb, "exact").search(q, k)
    assert np.array_equal(Ia, Ie) and np.array_equal(Da, De), "auto and exact must agree bit for bit"
    assert st["queries"] == 64


def test_batches_larger_than_one_pass_and_chunk(faiss):
    d, k = 768, 10
    xb = syn.corpus_rows(1, 20000, d)
    q = syn.queries(300, d, [1], seed=9)       # > 64 (one tc pass) and > 256 (one chunk)
    D, I, _, st = _check_vs_oracle(faiss, xb, q, k, "auto", expect_no_rescan=True)
    D1, I1 = _index(faiss, xb, "auto").search(q[17:18], k)  # same answer alone as in the batch
    assert np.array_equal(I1[0], I[17]) and np.array_equal(D1[0], D[17])


def test_tc_scan_candidates_match_bf16_reference(faiss):
    """The raw tensor-core scan: the union of its candidate lists must contain the
    top rows of a bf16-input / fp32-accumulate reference, with matching scores."""
    d = 768
    xb = syn.corpus_rows(0, 30000, d)
    q = syn.queries(9, d, [0], seed=4)
    ix = _index(faiss, xb, "auto")
    qt = torch.from_numpy(q).cuda()
    cs, cr, n_lists, depth = ix.debug_tc_candidates(qt)
    torch.cuda.synchronize()
    cs, cr = cs.cpu().numpy(), cr.cpu().numpy().astype(np.int64) & 0xFFFFFFFF

    def bf16_rna(a):  # the library's rounding: nearest, ties away (bits + 0x8000) >> 16
        b = (a.view(np.uint32) + np.uint32(0x8000)) & np.uint32(0xFFFF0000)
        return b.view(np.float32)

    ref = bf16_rna(q).astype(np.float64) @ bf16_rna(xb).astype(np.float64).T
    for b in range(q.shape[0]):
        valid = cr[b] != 0xFFFFFFFF
        rows, sc = cr[b][valid], cs[b][valid]
        assert len(np.unique(rows)) == len(rows), "a row appears in two candidate lists"
        np.testing.assert_allclose(sc, ref[b][rows], rtol=0, atol=2e-5)
        top = np.argsort(-ref[b])[:64]
        assert set(top.tolist()) <= set(rows.tolist()), "tensor-core scan lost a top-64 row"
    assert depth == 16 and n_lists >= 4


def test_adversarial_near_ties_fall_back_to_exact(faiss):
    """Rows identical after bf16 rounding but different in fp32: the bf16 scan cannot
    rank them, the certificate must notice and the exact path must return the fp32
    order."""
    d, k, n = 768, 32, 4096
    rng = np.random.default_rng(0)
    base = rng.standard_normal(d).astype(np.float32)
    base = ((base.view(np.uint32) + np.uint32(0x8000)) & np.uint32(0xFFFF0000)).view(np.float32)  # exactly bf16
    xb = np.tile(base, (n, 1))
    # perturb below bf16 resolution (relative 2^-10 < half an ulp = 2^-8), different per row
    xb *= (1.0 + rng.uniform(-1, 1, size=(n, 1)).astype(np.float32) * np.float32(2.0 ** -10))
    q = base[None, :].copy()
    ix = _index(faiss, xb, "auto")
    D, I = ix.search(q, k)
    Dr, Ir = oracle.search_seq(q, xb, k)
    oracle.compare_topk(D, I, Dr, Ir, x=q, xb=xb)
    st = ix.stats()
    assert st["exact_rescans"] >= 1, "the certificate should have rejected the bf16 ranking"
    assert st["exact_scans"] >= 1, "4096 rows inside one bf16 ulp overflow the second pass: exact scan expected"
    # the library's own contract: fp64-accumulated score rounded to fp32, ties by lower id
    t32 = (xb.astype(np.float64) @ q[0].astype(np.float64)).astype(np.float32)
    want = np.lexsort((np.arange(n), -t32.astype(np.float64)))[:k]
    assert I[0].tolist() == want.tolist() and np.array_equal(D[0], t32[want])


def test_all_rows_identical(faiss):
    xb = np.ones((1000, 64), dtype=np.float32)
    q = np.ones((2, 64), dtype=np.float32)
    for mode in ("exact", "auto"):
        D, I = _index(faiss, xb, mode).search(q, 8)
        assert (D == 64.0).all() and I[0].tolist() == list(range(8))


def test_search_tensor_mask_and_strided_output(faiss):
    d, k = 768, 32
    xb = syn.corpus_rows(2, 9000, d)
    q = syn.queries(12, d, [2], seed=8)
    ix = _index(faiss, xb, "auto")
    qt = torch.from_numpy(q).cuda()
    mask = torch.zeros((12, 3), dtype=torch.uint8, device="cuda")
    mask[::2, 1] = 1
    D_all = torch.full((12, 3 * k), 7.0, dtype=torch.float32, device="cuda")
    I_all = torch.full((12, 3 * k), 7, dtype=torch.int64, device="cuda")
    ix.search_tensor(qt, k, mask=mask, mask_col=1, out=(D_all, I_all), out_col=k)
    torch.cuda.synchronize()
    D_all, I_all = D_all.cpu().numpy(), I_all.cpu().numpy()
    Dr, Ir = ix.search(q, k)
    assert (D_all[:, :k] == 7).all() and (I_all[:, 2 * k:] == 7).all(), "wrote outside its columns"
    assert np.array_equal(I_all[::2, k:2 * k], Ir[::2]) and np.array_equal(D_all[::2, k:2 * k], Dr[::2])
    assert (I_all[1::2, k:2 * k] == -1).all() and (D_all[1::2, k:2 * k] == -FLT_MAX).all()
    # nothing routed at all: padding everywhere, corpus untouched
    mask.zero_()
    D2, I2 = ix.search_tensor(qt, k, mask=mask, mask_col=0)
    assert (I2 == -1).all().item()


def test_id_offset_and_incremental_add(faiss):
    d, k = 384, 10
    xb = syn.corpus_rows(1, 3000, d)
    q = syn.queries(4, d, [1], seed=2)
    ix = faiss.IndexFlatIP(d)
    ix.add(xb[:1000])
    ix.add(torch.from_numpy(xb[1000:]).cuda())     # device-side add, grows the planes
    assert ix.ntotal == 3000
    D, I = ix.search(q, k)
    Dr, Ir = oracle.search_seq(q, xb, k)
    oracle.compare_topk(D, I, Dr, Ir, x=q, xb=xb)
    ix.set_id_offset(1_000_000_000_000)
    D2, I2 = ix.search(q, k)
    assert np.array_equal(I2, I + 1_000_000_000_000) and np.array_equal(D2, D)


def test_reconstruct_is_bit_exact_and_index_file_roundtrip(faiss):
    rng = np.random.default_rng(3)
    xb = rng.standard_normal((777, 100), dtype=np.float32)
    xb[0, :4] = [0.0, -0.0, 1e-30, 3.0e38]
    ix = _index(faiss, xb)
    back = ix.reconstruct_n(0, 777)
    assert np.array_equal(back.view(np.uint32), xb.view(np.uint32)), "hi/lo planes must rebuild fp32 exactly"
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "faiss.index")
        faiss.write_index(ix, path)
        ix2 = faiss.read_index(path)
        assert ix2.ntotal == 777 and ix2.d == 100
        q = rng.standard_normal((2, 100), dtype=np.float32)
        a, b = ix.search(q, 5), ix2.search(q, 5)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])


def _faiss_flat_file(path, xb, fourcc=b"IxFI", metric=0):
    """An index file laid out as upstream faiss 1.7.4 writes IndexFlat (index_write.cpp: fourcc, write_index_header,
    then the vector of floats), produced here byte by byte -- independent of the library's own writer."""
    import struct
    n, d = xb.shape
    with open(path, "wb") as f:
        f.write(fourcc)
        f.write(struct.pack("<i", d)); f.write(struct.pack("<q", n))
        f.write(struct.pack("<q", 1 << 20)); f.write(struct.pack("<q", 1 << 20))
        f.write(struct.pack("<B", 1)); f.write(struct.pack("<i", metric))
        f.write(struct.pack("<Q", n * d))
        f.write(np.ascontiguousarray(xb, dtype="<f4").tobytes())


def test_read_index_streams_a_faiss_laid_out_file(faiss, tmp_path):
    """faiss.read_index (data_source.py:71) on a hand-laid-out IxFI file larger than one streaming chunk (64 MB:
    the double-buffered disk -> pinned -> device path really alternates buffers), bit-exact rows, same search
    results as index.add of the same rows; the writer's output is byte-identical to the hand-laid-out file;
    non-flat / L2 / truncated files are rejected loudly."""
    from ragroute_b200._lib import RRError
    rng = np.random.default_rng(12)
    n, d = 70_001, 300                                   # 84 MB of rows: two chunks, the second partial
    xb = rng.standard_normal((n, d), dtype=np.float32)
    path = str(tmp_path / "faiss.index")
    _faiss_flat_file(path, xb)
    ix = faiss.read_index(path)
    assert ix.ntotal == n and ix.d == d
    for r0 in (0, 55_923, n - 7):                        # around the chunk boundary (55 924 rows per chunk) and the end
        back = ix.reconstruct_n(r0, 7)
        assert np.array_equal(back.view(np.uint32), xb[r0:r0 + 7].view(np.uint32))
    q = rng.standard_normal((3, d), dtype=np.float32)
    D, I = ix.search(q, 10)
    D2, I2 = _index(faiss, xb).search(q, 10)
    assert np.array_equal(I, I2) and np.array_equal(D, D2)
    out = str(tmp_path / "written.index")
    faiss.write_index(ix, out)
    assert open(out, "rb").read() == open(path, "rb").read(), "write_index must reproduce the faiss layout byte for byte"
    bad = str(tmp_path / "bad.index")
    _faiss_flat_file(bad, xb[:10], fourcc=b"IxF2", metric=1)          # IndexFlatL2
    with pytest.raises(RRError):
        faiss.read_index(bad)
    _faiss_flat_file(bad, xb[:10], fourcc=b"IwFl")                     # IVF: not a flat index
    with pytest.raises(RRError):
        faiss.read_index(bad)
    with open(path, "rb") as f:
        head = f.read(100_000)
    with open(bad, "wb") as f:
        f.write(head)                                                  # truncated rows
    with pytest.raises(RRError):
        faiss.read_index(bad)
    with pytest.raises(RRError):
        faiss.read_index(str(tmp_path / "missing.index"))


def test_normalize_l2_matches_oracle(faiss):
    rng = np.random.default_rng(2)
    x = rng.standard_normal((3, 768), dtype=np.float32)
    x[1] = 0
    y, z = x.copy(), x.copy()
    faiss.normalize_L2(y)
    oracle.normalize_L2(z)
    np.testing.assert_allclose(y, z, rtol=1e-6, atol=1e-8)
    assert (y[1] == 0).all()


def test_errors_like_faiss(faiss):
    ix = _index(faiss, np.ones((10, 8), dtype=np.float32))
    with pytest.raises(AssertionError):
        ix.search(np.ones((1, 9), dtype=np.float32), 3)     # wrong d
    with pytest.raises(AssertionError):
        ix.search(np.ones((1, 8), dtype=np.float32), 0)     # k must be > 0
    from ragroute_b200._lib import RRError
    with pytest.raises(RRError):
        ix.search(np.ones((1, 8), dtype=np.float32), 33)    # k > RR_MAX_K: loud, not silent


# ---------------------------------------------------------------------------
# Large-batch path (rr_scan_gemm.cu: queries stationary in TMEM).  "scan_path" 2 (one CTA
# per 128-query tile) / 3 (CTA pairs, tcgen05 cta_group::2, 256 queries per pair) force it for
# any batch size so that small cases reach it too.
# ---------------------------------------------------------------------------
def _gemm_index(faiss, xb, path=2):
    ix = _index(faiss, xb, "auto")
    ix.set_param("scan_path", path)
    return ix


@pytest.mark.parametrize("path", [2, 3])
@pytest.mark.parametrize("n,d,k,nq", [(5000, 768, 32, 5), (20000, 768, 32, 130), (4000, 384, 10, 64),
                                      (2500, 100, 7, 2), (63, 768, 32, 3), (9000, 512, 10, 257),
                                      (7000, 640, 10, 300)])
def test_gemm_path_vs_oracle(faiss, n, d, k, nq, path):
    rng = np.random.default_rng(n + d + nq)
    xb = rng.standard_normal((n, d), dtype=np.float32)
    q = rng.standard_normal((nq, d), dtype=np.float32)
    ix = _gemm_index(faiss, xb, path)
    D, I = ix.search(q, k)
    Dr, Ir = oracle.search(q, xb, k)
    oracle.compare_topk(D, I, Dr, Ir, x=q, xb=xb)
    De, Ie = _index(faiss, xb, "exact").search(q, k)
    assert np.array_equal(I, Ie) and np.array_equal(D, De), "gemm path and exact must agree bit for bit"


@pytest.mark.parametrize("d", [832, 1000, 1024])
def test_gemm_path_wide_rows_up_to_1024_dims(faiss, d):
    """768 < d_pad <= 1024 (FeB4RAG's 1024-d encoders, config.py:44-58): the query tile's tail beyond
    TMEM lives in shared memory -- 768 + 256 dims with one accumulator stage for a single query tile,
    512 + 512 with two for CTA pairs.  Default dispatch (single CTA / pairs picked on the device), both
    routed densities, against the oracle and bit for bit against exact mode and the swap-AB scan."""
    k = 10
    rng = np.random.default_rng(d)
    xb = (rng.standard_normal((21000, d), dtype=np.float32) / np.float32(np.sqrt(d))
          + rng.standard_normal(d).astype(np.float32) * np.float32(0.5 / np.sqrt(d)))
    ixe = _index(faiss, xb, "exact")
    for nq, density in ((30, 1.0), (100, 1.0), (300, 1.0), (300, 0.3), (2100, 0.7)):
        q = rng.standard_normal((nq, d), dtype=np.float32) / np.float32(np.sqrt(d))
        ix = _index(faiss, xb, "auto")
        qt = torch.from_numpy(q).cuda()
        mask = torch.from_numpy((rng.random((nq, 1)) < density).astype(np.uint8)).cuda()
        D, I = ix.search_tensor(qt, k, mask=mask)
        torch.cuda.synchronize()
        D, I, m = D.cpu().numpy(), I.cpu().numpy(), mask.cpu().numpy()[:, 0].astype(bool)
        assert ix.stats()["exact_rescans"] == 0, f"certificate failed on benign data: {ix.stats()}"
        assert (I[~m] == -1).all()
        De, Ie = ixe.search(q, k)
        assert np.array_equal(I[m], Ie[m]) and np.array_equal(D[m], De[m]), "auto and exact must agree bit for bit"
        sel = np.nonzero(m)[0][:20]
        Dr, Ir = oracle.search(q[sel], xb, k)
        oracle.compare_topk(D[sel], I[sel], Dr, Ir, x=q[sel], xb=xb)
    # forced CTA pairs on a small batch, and the swap-AB scan, agree too
    q = rng.standard_normal((70, d), dtype=np.float32) / np.float32(np.sqrt(d))
    De, Ie = ixe.search(q, k)
    for path in (1, 3):
        ix = _gemm_index(faiss, xb, path)
        D, I = ix.search(q, k)
        assert np.array_equal(I, Ie) and np.array_equal(D, De)


@pytest.mark.parametrize("depth", [8, 32])
@pytest.mark.parametrize("nq,path", [(40, 2), (128, 2), (300, 3), (1100, 0)])
def test_gemm_path_list_depths_agree_with_exact(faiss, depth, nq, path):
    """Both candidate-list depths of the queries-in-TMEM scan (8 for small shards, 32 for large ones; forced here)
    on a MedRAG-shaped corpus: no certificate failure and bit-identical to the exact mode."""
    d, k = 768, 32
    xb = syn.corpus_rows(1, 50000, d)
    q = syn.queries(nq, d, [1], seed=nq + depth)
    ix = _gemm_index(faiss, xb, path)
    ix.set_param("gemm_depth", depth)
    D, I = ix.search(q, k)
    if depth == 32 or nq <= 256:     # (8-entry lists are for ONE query tile; forced beyond that, certificates may fail --
        assert ix.stats()["exact_rescans"] == 0, f"certificate failed on benign data (depth {depth}): {ix.stats()}"
    De, Ie = _index(faiss, xb, "exact").search(q, k)   # -- and the repair path keeps the result exact anyway)
    assert np.array_equal(I, Ie) and np.array_equal(D, De)


def test_gemm_path_medrag_shaped_no_rescans_and_equals_small_path(faiss):
    """On benign MedRAG-shaped data the certificate must pass for every query (otherwise
    the exact rescan would hide a broken tensor-core kernel), for 1, 2 and 8 query tiles."""
    d, k = 768, 32
    xb = syn.corpus_rows(2, 70000, d)
    for nq, path in ((100, 2), (100, 3), (256, 2), (256, 3), (1024, 2), (1024, 3), (700, 3), (2048, 3), (60, 0), (200, 0)):
        q = syn.queries(nq, d, [2], seed=nq)
        ix = _gemm_index(faiss, xb, path)
        D, I = ix.search(q, k)
        assert ix.stats()["exact_rescans"] == 0, f"certificate failed on benign data (path {path}): {ix.stats()}"
        ixs = _index(faiss, xb, "auto")
        ixs.set_param("scan_path", 1)
        Ds, Is = ixs.search(q, k)
        assert ixs.stats()["exact_rescans"] == 0
        assert np.array_equal(I, Is) and np.array_equal(D, Ds), "both tensor-core paths must agree bit for bit"
        Dr, Ir = oracle.search(q[:40], xb, k)
        oracle.compare_topk(D[:40], I[:40], Dr, Ir, x=q[:40], xb=xb)


@pytest.mark.parametrize("nq,path", [(600, 3), (2048, 3), (300, 2), (900, 0)])
def test_gemm_path_lockstep_does_not_change_results(faiss, nq, path):
    """The lockstep of the units sharing a corpus slice (several query tiles: 2 ... 8 siblings per slice) only paces
    their TMA producers: results are bit-identical with it switched off, and no launch hangs or times out."""
    d, k = 768, 32
    xb = syn.corpus_rows(3, 90000, d)
    q = syn.queries(nq, d, [3], seed=nq)
    out = []
    for lockstep in (4, 0, 1):                 # default window, free running, the tightest window
        ix = _gemm_index(faiss, xb, path)
        ix.set_param("lockstep", lockstep)
        for _ in range(3):                     # repeated launches on one handle: the launch number keeps counting
            D, I = ix.search(q, k)
        assert ix.stats()["exact_rescans"] == 0
        out.append((D, I))
    for D, I in out[1:]:
        assert np.array_equal(I, out[0][1]) and np.array_equal(D, out[0][0])
    De, Ie = _index(faiss, xb, "exact").search(q[:64], k)
    assert np.array_equal(out[0][1][:64], Ie) and np.array_equal(out[0][0][:64], De)


@pytest.mark.parametrize("n,d,nq,path", [(60000, 768, 700, 3), (60000, 768, 2048, 0), (40000, 384, 600, 3),
                                         (30000, 640, 300, 3), (20000, 1024, 520, 3), (9000, 100, 257, 3)])
def test_gemm_path_64k_stages_agree_with_exact(faiss, n, d, nq, path):
    """CTA pairs with 64 KB TMA stages (16 MMAs per barrier round trip; batches > 256, d_pad <= 768 -- wider rows fall
    back to 32 KB stages): bit-identical to the exact mode and to the 32 KB-stage launch, no certificate failure."""
    k = 32 if d == 768 else 10
    rng = np.random.default_rng(n + d + nq)
    xb = syn.corpus_rows(1, n, d) if d == 768 else rng.standard_normal((n, d), dtype=np.float32) / np.float32(np.sqrt(d))
    q = syn.queries(nq, d, [1], seed=nq) if d == 768 else rng.standard_normal((nq, d), dtype=np.float32)
    res = {}
    for skc in (4, 2):
        ix = _gemm_index(faiss, xb, path)
        ix.set_param("gemm_stage_kc", skc)
        res[skc] = ix.search(q, k)
        if d == 768:
            assert ix.stats()["exact_rescans"] == 0, f"certificate failed on benign data (stage_kc {skc}): {ix.stats()}"
    assert np.array_equal(res[4][1], res[2][1]) and np.array_equal(res[4][0], res[2][0])
    De, Ie = _index(faiss, xb, "exact").search(q[:96], k)
    assert np.array_equal(res[4][1][:96], Ie) and np.array_equal(res[4][0][:96], De)


def test_gemm_path_mask_chunks_and_default_dispatch(faiss):
    d, k = 768, 10
    xb = syn.corpus_rows(1, 30000, d)
    q = syn.queries(2300, d, [1], seed=12)     # > 2048: two chunks on the large-batch path
    ix = _index(faiss, xb, "auto")             # default dispatch: nq > 64 takes the large-batch path
    qt = torch.from_numpy(q).cuda()
    mask = (torch.arange(2300, device="cuda") % 3 != 0).to(torch.uint8).reshape(-1, 1).contiguous()
    D, I = ix.search_tensor(qt, k, mask=mask)
    torch.cuda.synchronize()
    D, I = D.cpu().numpy(), I.cpu().numpy()
    assert ix.stats()["exact_rescans"] == 0
    off = (np.arange(2300) % 3 == 0)
    assert (I[off] == -1).all() and (D[off] == -FLT_MAX).all()
    sel = np.nonzero(~off)[0][::37]
    Dr, Ir = oracle.search(q[sel], xb, k)
    oracle.compare_topk(D[sel], I[sel], Dr, Ir, x=q[sel], xb=xb)


# ---------------------------------------------------------------------------
# BASELINE.json sizes.  configs[0] (Textbooks alone, 125 000 x 768, k = 32) is checked against
# the oracle itself; a multi-million-row corpus is checked through size-independent properties
# against a float64 torch reference on the device (the CPU oracle would need minutes per query).
# ---------------------------------------------------------------------------
def test_config0_textbooks_full_size_vs_oracle(faiss):
    d, k = syn.MEDRAG_DIM, syn.MEDRAG_K
    xb = syn.corpus_rows(syn.MEDRAG_SOURCE_TO_ID["textbooks"], syn.MEDRAG_ROWS["textbooks"], d)
    q = syn.queries(6, d, [2], seed=21)
    ix = _index(faiss, xb, "auto")
    D, I = ix.search(q, k)                       # one query at a time is what the reference does; 6 at once here
    Dr, Ir = oracle.search(q, xb, k)
    oracle.compare_topk(D, I, Dr, Ir, x=q, xb=xb)
    D1, I1 = ix.search(q[:1], k)                 # nq = 1, the reference's calling convention (data_source.py:114)
    assert np.array_equal(I1[0], I[0]) and np.array_equal(D1[0], D[0])
    assert ix.stats()["exact_rescans"] == 0


@pytest.mark.parametrize("nq", [3, 64, 600])
def test_multi_million_rows_properties(faiss, nq):
    """3 M x 768 rows generated on the device (seeded): for every kernel family, the returned lists
    must be sorted, duplicate-free, carry the float64 scores of their rows (<= 1e-5 relative) and no
    other row of the corpus may beat the k-th score (checked exhaustively in float64 on the GPU)."""
    from ragroute_b200.workloads import device_rows
    d, k, n = 768, 32, 3_000_000
    dev = torch.device("cuda", 0)
    ix = faiss.IndexFlatIP(d, reserve=n)
    for _, block in device_rows(0, 0, n, d, dev):
        ix.add(block)
    q = torch.from_numpy(syn.queries(nq, d, [0], seed=100 + nq)).cuda()
    D, I = ix.search_tensor(q, k)
    torch.cuda.synchronize()
    assert ix.stats()["exact_rescans"] == 0
    assert (D[:, :-1] >= D[:, 1:]).all().item(), "scores not descending"
    assert all(len(set(r)) == k for r in I.cpu().tolist()), "duplicate ids"
    check = torch.arange(0, nq, max(1, nq // 12), device=dev)           # a dozen queries, exhaustively
    qd = q[check].double()
    best = torch.full((len(check), k), -float("inf"), dtype=torch.float64, device=dev)
    for g0, block in device_rows(0, 0, n, d, dev):
        s = qd @ block.double().T                                         # [c, rows] float64
        best = torch.topk(torch.cat([best, s], dim=1), k, dim=1).values
        sel = (I[check] >= g0) & (I[check] < g0 + block.shape[0])       # float64 scores of the returned rows
        if sel.any():
            rows = (I[check] - g0).clamp(0, block.shape[0] - 1)
            got = torch.gather(s, 1, rows)
            ref = D[check].double()
            assert ((got - ref).abs() <= 1e-5 * ref.abs().clamp_min(1e-30))[sel].all().item(), "score mismatch"
    kth_ref, kth_got = best[:, -1], D[check][:, -1].double()
    assert ((kth_ref - kth_got).abs() <= 1e-5 * kth_ref.abs()).all().item(), "a better row exists than the returned k-th"
    assert ((best - D[check].double()).abs() <= 1e-5 * best.abs()).all().item(), "top-k scores differ from float64 truth"


@pytest.mark.parametrize("nq,path", [(1, 0), (40, 0), (64, 0), (300, 0), (300, 2), (1500, 0)])
def test_near_duplicate_documents_are_settled_by_the_second_pass(faiss, nq, path):
    """Documents stored as runs of ~100 near-duplicate chunks: the whole top-32 of a query falls inside one
    candidate list's territory, so the first-pass certificate MUST fail; the thresholded second tensor-core
    pass collects every row that can matter and the result is bit-identical to the exact scan, without it."""
    d, k, docs, per = 768, 32, 300, 100
    rng = np.random.default_rng(11)
    centers = (rng.standard_normal((docs, d)) / np.sqrt(d) + 0.02).astype(np.float32)
    xb = (np.repeat(centers, per, axis=0) +
          0.03 * rng.standard_normal((docs * per, d)).astype(np.float32) / np.float32(np.sqrt(d))).astype(np.float32)
    pick = rng.integers(0, docs, size=nq)
    q = (centers[pick] + 0.05 * rng.standard_normal((nq, d)).astype(np.float32) / np.float32(np.sqrt(d))).astype(np.float32)
    ix = _index(faiss, xb, "auto")
    if path:
        ix.set_param("scan_path", path)
    D, I = ix.search(q, k)
    st = ix.stats()
    assert st["exact_rescans"] > 0.5 * nq, f"expected certificate failures on near-duplicate runs: {st}"
    assert st["exact_scans"] == 0, f"the second pass should have settled every query: {st}"
    De, Ie = _index(faiss, xb, "exact").search(q, k)
    assert np.array_equal(I, Ie) and np.array_equal(D, De), "second pass and exact scan must agree bit for bit"
    assert ((I // per) == pick[:, None]).mean() > 0.95     # the top-32 really are the query's own document
    sel = np.arange(0, nq, max(1, nq // 5))
    Dr, Ir = oracle.search(q[sel], xb, k)
    oracle.compare_topk(D[sel], I[sel], Dr, Ir, x=q[sel], xb=xb)


def test_plain_c_caller_known_answers(faiss, tmp_path):
    """tests/c/abi_check.c --compute: a C program builds a 64 x 8 index through the C ABI and checks the
    hand-computable top-5 (ties: lower row first) in auto and exact mode."""
    import subprocess
    from golden_util import build_abi_check
    exe = build_abi_check(tmp_path)
    res = subprocess.run([exe, "--compute"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "abi_check ok (compute" in res.stdout


@pytest.mark.parametrize("seed", range(int(os.environ.get("RR_FUZZ_CASES", "36"))))
def test_fuzz_auto_equals_exact(faiss, seed):
    """Randomised shapes through the default dispatch (swap-AB scan, queries in TMEM on single CTAs / CTA
    pairs picked on the device, chained launches, second pass) against the exact-scan mode of the same
    library, which the oracle pins: ids and scores bit for bit, padding for unrouted queries, id offsets.
    The exact mode is one simple kernel; agreement on ragged sizes, duplicates and partial tiles is the
    property that the tensor-core machinery changes HOW FAST, never WHAT."""
    rng = np.random.default_rng(9000 + seed)
    d = int(rng.choice([8, 64, 100, 384, 512, 640, 768, 832, 1024]))
    n = int(np.exp(rng.uniform(np.log(1), np.log(40000))))
    nq = int(np.exp(rng.uniform(np.log(1), np.log(2300))))
    nq = max(1, min(nq, int(3e7 // max(n, 1))))
    k = int(rng.integers(1, 33))
    density = float(rng.choice([0.0, 0.05, 0.3, 0.7, 1.0]))
    offset = int(rng.choice([0, 12345, 1 << 33]))
    mu = rng.standard_normal(d).astype(np.float32) * np.float32(0.5 / np.sqrt(d))
    xb = rng.standard_normal((n, d), dtype=np.float32) / np.float32(np.sqrt(d)) + mu
    flavour = seed % 4
    if flavour == 1 and n > 8:                       # exact duplicates (ties broken by row)
        xb[rng.integers(0, n, size=n // 3)] = xb[rng.integers(0, n)]
    elif flavour == 2 and n > 8:                     # a run of near-duplicates: certificates fail
        j = int(rng.integers(0, n))
        m = min(n, 300)
        lo = max(0, min(j, n - m))
        xb[lo:lo + m] = xb[j] * (1 + 1e-4 * rng.standard_normal((m, 1)).astype(np.float32))
    elif flavour == 3:                               # small integers: exact in bf16, many equal scores
        xb = rng.integers(-3, 4, size=(n, d)).astype(np.float32)
    q = rng.standard_normal((nq, d), dtype=np.float32) / np.float32(np.sqrt(d)) + mu
    if flavour == 3:
        q = rng.integers(-2, 3, size=(nq, d)).astype(np.float32)
    mask = (rng.random((nq, 1)) < density).astype(np.uint8)
    res = {}
    for mode in ("auto", "exact"):
        ix = _index(faiss, xb, mode)
        ix.set_id_offset(offset)
        D, I = ix.search_tensor(torch.from_numpy(q).cuda(), k, mask=torch.from_numpy(mask).cuda())
        torch.cuda.synchronize()
        res[mode] = (D.cpu().numpy(), I.cpu().numpy())
    (Da, Ia), (De, Ie) = res["auto"], res["exact"]
    tag = f"seed={seed} n={n} d={d} nq={nq} k={k} density={density} flavour={flavour}"
    assert np.array_equal(Ia, Ie), tag
    assert np.array_equal(Da, De), tag
    m = mask[:, 0].astype(bool)
    assert (Ia[~m] == -1).all() and (Da[~m] == -FLT_MAX).all(), tag
    if m.any():
        kk = min(k, n)
        assert (Ia[m][:, :kk] >= offset).all() and (Ia[m][:, :kk] < offset + n).all(), tag
        assert (Ia[m][:, kk:] == -1).all(), tag
        if n > 1:
            assert (np.diff(Da[m][:, :kk], axis=1) <= 0).all(), tag      # descending
    # the exact mode itself against the oracle on a few routed queries
    sel = np.nonzero(m)[0][:4]
    if len(sel):
        Dr, Ir = oracle.search_seq(q[sel], xb, k)
        oracle.compare_topk(De[sel], np.where(Ie[sel] >= 0, Ie[sel] - offset, -1), Dr, Ir, x=q[sel], xb=xb)


@pytest.mark.parametrize("nq", [1, 40, 300])
def test_certificate_is_sound_at_bf16_midpoints(faiss, nq):
    """Worst case of the bf16 rounding error: query and the true top rows sit just below bf16
    midpoints in every coordinate, so the tensor-core scan under-scores them by 2^-7 relative, while decoy rows
    whose coordinates are (mostly) bf16-exact are under-scored by much less and fill every candidate list.  A
    certificate that assumes 2^-9 per operand accepts the decoys; the bound from the measured rounding-error
    norms must reject them and the repair path must return the exact top-k."""
    d, k, n = 768, 32, 6000
    u = np.float32(2.0 ** -8)
    m = np.array([0x3F807FFF], dtype=np.uint32).view(np.float32)[0]      # 1 + 2^-8 - 2^-23: rounds DOWN to 1.0
    p, lo = np.float32(1.0) + 2 * u, np.float32(1.0) - u                  # bf16-exact neighbours
    rng = np.random.default_rng(5)
    xb = (0.3 * rng.standard_normal((n, d))).astype(np.float32)          # filler: low scores
    pattern = np.array([p] * (d // 4) + [m] * (d // 2) + [lo] * (d // 4), dtype=np.float32)
    decoys = np.arange(0, n, 3)
    for r in decoys:                                                     # exact 768 m (1 + 0.75 u), bf16 768 (1 + 0.25 u)
        xb[r] = rng.permutation(pattern)
    top = np.array([r for r in range(1, n, 149)][:40])                   # exact 768 m^2, bf16 768.0: below every decoy
    xb[top] = m
    q = np.full((nq, d), m, dtype=np.float32)
    ix = _index(faiss, xb, "auto")
    D, I = ix.search(q, k)
    st = ix.stats()
    De, Ie = _index(faiss, xb, "exact").search(q, k)
    assert np.array_equal(I, Ie) and np.array_equal(D, De), f"auto differs from exact at bf16 midpoints: {st}"
    assert I[0].tolist() == sorted(top.tolist())[:k], "the rows at the midpoints are the true top-k"
    assert st["exact_rescans"] == nq, f"every certificate must fail here: {st}"


def test_search_with_normalize_is_one_call_and_in_place(faiss):
    """retrieve_docs_wikipedia (data_source.py:199-203): normalize_L2 + search as one host round trip."""
    rng = np.random.default_rng(8)
    xb = rng.standard_normal((5000, 768), dtype=np.float32)
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    q = rng.standard_normal((3, 768), dtype=np.float32) * 7
    ix = _index(faiss, xb, "auto")
    q1, q2 = q.copy(), q.copy()
    faiss.normalize_L2(q1)
    D1, I1 = ix.search(q1, 10)
    D2, I2 = ix.search(q2, 10, normalize=True)
    assert np.array_equal(I1, I2) and np.array_equal(D1, D2)
    assert np.array_equal(q1, q2), "normalize=True must leave the normalised queries in the caller's array"
    # a device-API search followed by a host search on the same handle (different streams) stays ordered
    qt = torch.from_numpy(q1).cuda()
    Dt, It = ix.search_tensor(qt, 10)
    D3, I3 = ix.search(q1, 10)
    torch.cuda.synchronize()
    assert np.array_equal(It.cpu().numpy(), I3) and np.array_equal(Dt.cpu().numpy(), D3)

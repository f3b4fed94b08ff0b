"""oracle/flat_ip.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU oracle for the reference's per-source search
``index.search(query_embed, k)`` (/root/reference/ragroute/data_source.py:158,
:186, :203), i.e. FAISS ``IndexFlatIP.search`` from the un-vendored dependency
faiss-cpu==1.7.4 (/root/reference/environment.yml:55).

* ``search_seq``   -- ctypes binding of ``flat_ip_ref.c`` (FAISS' nq < 20 branch:
  sequential scan, AVX2-ordered fp32 dot product, strict-``>`` min-heap).  This is
  the branch the reference always takes (one query per call, data_source.py:114).
* ``search_blas``  -- FAISS' nq >= 20 branch (exhaustive_inner_product_blas):
  query blocks of 4096 x corpus blocks of 1024, ``sgemm`` (numpy/OpenBLAS fp32),
  then heap ``add_results`` per block with the same strict-``>`` rule.
* ``search``       -- FAISS' dispatch: ``nq < 20`` -> seq, else blas
  (``distance_compute_blas_threshold = 20``).
* ``truth_f64``    -- exhaustive float64 scores; ``classify_mismatch`` applies the
  north_star carve-out (an id mismatch is a documented tie iff the two float64
  scores differ by <= 1e-5 relative).

PARITY UNPINNED: no golden vector exists in the reference for this path and FAISS
cannot be installed here; see the header of flat_ip_ref.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs
may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libflat_ip_ref.so")
_lib = None

FLT_MAX = np.float32(np.finfo(np.float32).max)
BLAS_THRESHOLD = 20  # faiss::distance_compute_blas_threshold
BLAS_QUERY_BS = 4096  # faiss::distance_compute_blas_query_bs
BLAS_DB_BS = 1024  # faiss::distance_compute_blas_database_bs


def build(force: bool = False) -> str:
    """Compile flat_ip_ref.c with the Makefile next to it (gcc, no other deps)."""
    src = os.path.join(_HERE, "flat_ip_ref.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libflat_ip_ref.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        lib = ctypes.CDLL(_LIB_PATH)
        f32p = ctypes.POINTER(ctypes.c_float)
        i64p = ctypes.POINTER(ctypes.c_int64)
        f64p = ctypes.POINTER(ctypes.c_double)
        i64 = ctypes.c_int64
        for name in ("ref_flat_ip_search_seq", "ref_flat_ip_search_rowpar"):
            fn = getattr(lib, name)
            fn.argtypes = [f32p, f32p, i64, i64, i64, i64, f32p, i64p, ctypes.c_int]
            fn.restype = ctypes.c_int
        lib.ref_heap_begin.argtypes = [i64, i64, f32p, i64p]
        lib.ref_heap_begin.restype = ctypes.c_int
        lib.ref_heap_add_results.argtypes = [f32p, i64, i64, i64, i64, i64, f32p, i64p, ctypes.c_int]
        lib.ref_heap_add_results.restype = ctypes.c_int
        lib.ref_heap_end.argtypes = [i64, i64, f32p, i64p]
        lib.ref_heap_end.restype = ctypes.c_int
        lib.ref_flat_ip_scores_f64.argtypes = [f32p, f32p, i64, i64, f64p]
        lib.ref_flat_ip_scores_f64.restype = ctypes.c_int
        lib.ref_normalize_L2.argtypes = [f32p, i64, i64]
        lib.ref_normalize_L2.restype = ctypes.c_int
        _lib = lib
    return _lib


def _f32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _check(x, xb):
    x = np.ascontiguousarray(x, dtype=np.float32)
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    if x.ndim != 2 or xb.ndim != 2 or x.shape[1] != xb.shape[1]:
        raise AssertionError("x must be [nq, d] and xb [nb, d]")  # faiss asserts d
    return x, xb


def search_seq(x, xb, k, n_threads: int = 0, row_parallel: bool = False):
    """FAISS nq<20 branch. Returns (D float32[nq,k] descending, I int64[nq,k])."""
    lib = _load()
    x, xb = _check(x, xb)
    if k <= 0:
        raise AssertionError("k > 0")
    nq, d = x.shape
    nb = xb.shape[0]
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    fn = lib.ref_flat_ip_search_rowpar if row_parallel else lib.ref_flat_ip_search_seq
    rc = fn(_f32(x), _f32(xb), nq, nb, d, k, _f32(D),
            I.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), int(n_threads))
    if rc != 0:
        raise RuntimeError(f"flat_ip_ref rc={rc}")
    return D, I


def _heap_add_block(D, I, ip_block, j0):
    """HeapResultHandler::add_results for one corpus block, vectorised in numpy (kept as the
    independent cross-check of the C heap pass: tests/test_oracle_flat_ip.py): keeping the
    running top-k under "replace the minimum on strict >" gives the k best scores; an equal
    score never displaces an earlier row, but WHICH of several rows tied at the k-th score is
    evicted by a better one depends on the heap's shape -- here the latest.  ip_block: [nq_blk, nb_blk] fp32."""
    k = D.shape[1]
    nq_blk, nb_blk = ip_block.shape
    ids = np.arange(j0, j0 + nb_blk, dtype=np.int64)
    cat_s = np.concatenate([D, ip_block], axis=1)
    cat_i = np.concatenate([I, np.broadcast_to(ids, (nq_blk, nb_blk))], axis=1)
    # stable sort on -score keeps earlier (already-kept / lower-id) rows first at ties
    order = np.argsort(-cat_s, axis=1, kind="stable")[:, :k]
    D[:] = np.take_along_axis(cat_s, order, axis=1)
    I[:] = np.take_along_axis(cat_i, order, axis=1)


def search_blas(x, xb, k, n_threads: int = 0, numpy_heap: bool = False):
    """FAISS nq>=20 branch: blocked sgemm (numpy -> OpenBLAS, all cores) + the C heap pass
    ``add_results`` per block (OpenMP over queries, as FAISS).  ``numpy_heap=True`` runs the
    vectorised numpy restatement of the heap pass instead (slow; cross-check only).

    Within a query the result SET equals search_seq's; the order of equal scores inside the
    final list follows heap_reorder in both."""
    lib = _load()
    x, xb = _check(x, xb)
    nq, _ = x.shape
    nb = xb.shape[0]
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    i64p = ctypes.POINTER(ctypes.c_int64)
    if numpy_heap:
        D.fill(-FLT_MAX)
        I.fill(-1)
    else:
        lib.ref_heap_begin(nq, k, _f32(D), I.ctypes.data_as(i64p))
    ip_buf = np.empty((min(nq, BLAS_QUERY_BS), BLAS_DB_BS), dtype=np.float32)
    for i0 in range(0, nq, BLAS_QUERY_BS):
        i1 = min(nq, i0 + BLAS_QUERY_BS)
        Db, Ib = D[i0:i1], I[i0:i1]
        for j0 in range(0, nb, BLAS_DB_BS):
            j1 = min(nb, j0 + BLAS_DB_BS)
            ip = ip_buf[:i1 - i0, :j1 - j0]
            np.matmul(x[i0:i1], xb[j0:j1].T, out=ip)  # sgemm_
            if numpy_heap:
                _heap_add_block(Db, Ib, ip, j0)
            else:
                lib.ref_heap_add_results(_f32(ip_buf), i1 - i0, j1 - j0, BLAS_DB_BS, j0, k, _f32(Db),
                                         Ib.ctypes.data_as(i64p), int(n_threads))
    if not numpy_heap:
        lib.ref_heap_end(nq, k, _f32(D), I.ctypes.data_as(i64p))
    return D, I


def search(x, xb, k, n_threads: int = 0):
    """IndexFlatIP.search with FAISS' own dispatch on nq."""
    x, xb = _check(x, xb)
    if x.shape[0] < BLAS_THRESHOLD:
        return search_seq(x, xb, k, n_threads)
    return search_blas(x, xb, k, n_threads)


def normalize_L2(x):
    """faiss.normalize_L2 (in place), data_source.py:199."""
    if x.dtype != np.float32 or not x.flags.c_contiguous or x.ndim != 2:
        raise AssertionError("normalize_L2 needs a C-contiguous float32 [n, d] array")
    _load().ref_normalize_L2(_f32(x), x.shape[0], x.shape[1])


def truth_f64(x, xb):
    """Exhaustive float64 scores [nq, nb]."""
    lib = _load()
    x, xb = _check(x, xb)
    out = np.empty((x.shape[0], xb.shape[0]), dtype=np.float64)
    for i in range(x.shape[0]):
        lib.ref_flat_ip_scores_f64(_f32(x[i]), _f32(xb), xb.shape[0], xb.shape[1],
                                   out[i].ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def compare_topk(D_got, I_got, D_ref, I_ref, truth=None, x=None, xb=None,
                 rel_tol: float = 1e-5):
    """Parity check with the north_star tolerance, written out:

    * scores: |D_got - D_ref| <= rel_tol * max(|D_ref|, tiny) element-wise
      (padding slots must be exactly (-FLT_MAX, -1) on both sides);
    * ids: equal as per-query SETS, except that an id present on one side only is
      accepted iff its float64 score is within rel_tol (relative) of the float64
      k-th score -- a documented tie.  Order inside a query is checked through the
      scores (descending on both sides).

    Returns a dict with counts; raises AssertionError on a real mismatch.
    """
    D_got = np.asarray(D_got)
    I_got = np.asarray(I_got)
    assert D_got.shape == D_ref.shape and I_got.shape == I_ref.shape, (D_got.shape, D_ref.shape)
    nq, k = D_ref.shape
    pad_ref = I_ref < 0
    pad_got = I_got < 0
    assert np.array_equal(pad_ref, pad_got), "padding (-1) slots differ"
    assert np.all(D_got[pad_got] == -FLT_MAX), "padding score must be -FLT_MAX"
    v = ~pad_ref
    denom = np.maximum(np.abs(D_ref[v]), 1e-30)
    rel = np.abs(D_got[v].astype(np.float64) - D_ref[v].astype(np.float64)) / denom
    assert rel.size == 0 or rel.max() <= rel_tol, f"score mismatch: max rel {rel.max():.3e}"
    # descending order
    assert np.all(np.diff(D_got.astype(np.float64), axis=1) <= 0), "scores not descending"
    n_tie = 0

    def f64score(i, j):
        if truth is not None:
            return float(truth[i][j])
        assert x is not None and xb is not None, "id mismatch and no truth to classify it"
        return float(np.asarray(xb[j], dtype=np.float64) @ np.asarray(x[i], dtype=np.float64))

    for i in range(nq):
        a = set(I_got[i][I_got[i] >= 0].tolist())
        b = set(I_ref[i][I_ref[i] >= 0].tolist())
        if a == b:
            continue
        kth = min(f64score(i, j) for j in b)
        for j in (a - b) | (b - a):
            tj = f64score(i, j)
            assert abs(tj - kth) <= rel_tol * max(abs(kth), 1e-30), (
                f"query {i}: id {j} differs and is not a tie "
                f"(f64 score {tj!r} vs k-th {kth!r})")
            n_tie += 1
    return {"queries": nq, "tie_ids": n_tie,
            "max_rel_score_err": float(rel.max()) if rel.size else 0.0}

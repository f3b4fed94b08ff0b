"""Known-answer tests for the FAISS IndexFlatIP restatement (oracle/flat_ip*.{c,py}).
The reference holds no fixture for this path (parity unpinned), so the oracle is
pinned by hand-computable cases and by an exhaustive float64 truth pass."""
import numpy as np
import pytest

from oracle import flat_ip as o

FLT_MAX = np.finfo(np.float32).max


def test_hand_computable_integers_with_duplicates():
    # N=64, D=8 integer vectors: scores are exact small integers
    xb = np.zeros((64, 8), dtype=np.float32)
    for j in range(64):
        xb[j, j % 8] = j // 8 + 1          # row j has a single non-zero: value (j//8+1) at dim j%8
    xb[10] = xb[58]                         # exact duplicate of the best row for dim 2
    q = np.zeros((1, 8), dtype=np.float32)
    q[0, 2] = 2.0
    D, I = o.search_seq(q, xb, 5)
    # rows with dim 2: j = 2,10(dup of 58),18,...,58 -> values 1,(8),3,4,5,6,7,8 ; scores 2*value
    assert D[0].tolist() == [16.0, 16.0, 14.0, 12.0, 10.0]
    assert I[0].tolist()[2:] == [50, 42, 34]
    assert set(I[0].tolist()[:2]) == {10, 58}   # tie: both present, order implementation-defined


def test_fewer_rows_than_k_pads_with_minus_one():
    rng = np.random.default_rng(1)
    xb = rng.standard_normal((5, 16), dtype=np.float32)
    q = rng.standard_normal((2, 16), dtype=np.float32)
    D, I = o.search_seq(q, xb, 8)
    assert (I[:, 5:] == -1).all() and (D[:, 5:] == -FLT_MAX).all()
    assert sorted(I[0, :5].tolist()) == [0, 1, 2, 3, 4]
    assert np.all(np.diff(D[:, :5], axis=1) <= 0)
    D0, I0 = o.search_seq(q, xb[:0], 4)
    assert (I0 == -1).all()


@pytest.mark.parametrize("d,k", [(768, 32), (1024, 10), (384, 10), (100, 7)])
def test_seq_equals_float64_truth(d, k):
    rng = np.random.default_rng(d)
    xb = rng.standard_normal((3000, d), dtype=np.float32)
    q = rng.standard_normal((4, d), dtype=np.float32)
    D, I = o.search_seq(q, xb, k)
    t = o.truth_f64(q, xb)
    ti = np.argsort(-t, axis=1, kind="stable")[:, :k]
    Dt = np.take_along_axis(t, ti, axis=1).astype(np.float32)
    o.compare_topk(D, I, Dt, ti, truth=t)


def test_blas_branch_and_dispatch_agree_with_seq():
    rng = np.random.default_rng(5)
    xb = rng.standard_normal((2500, 64), dtype=np.float32)
    q = rng.standard_normal((25, 64), dtype=np.float32)
    Ds, Is = o.search_seq(q, xb, 10)
    Db, Ib = o.search(q, xb, 10)               # nq >= 20 -> blas branch
    o.compare_topk(Db, Ib, Ds, Is, x=q, xb=xb)
    Dr, Ir = o.search_seq(q, xb, 10, row_parallel=True)
    assert np.array_equal(Ir, Is) and np.array_equal(Dr, Ds)


def test_blas_branch_c_heap_pass_equals_its_numpy_restatement():
    """The C add_results heap pass (what bench.py's CPU legs time) against the vectorised numpy restatement
    of the same rule, over several corpus blocks, with ties and with fewer rows than k."""
    rng = np.random.default_rng(6)
    xb = rng.integers(-3, 4, size=(2600, 32)).astype(np.float32)      # many equal scores, 3 corpus blocks
    q = rng.integers(-2, 3, size=(33, 32)).astype(np.float32)
    Dc, Ic = o.search_blas(q, xb, 10)
    Dn, In = o.search_blas(q, xb, 10, numpy_heap=True)
    assert np.array_equal(Dc, Dn)                                      # same scores; which of several rows
    o.compare_topk(Dc, Ic, Dn, In, x=q, xb=xb)                         # TIED at the k-th score stays is heap-shaped
    Ds, Is = o.search_seq(q, xb, 10)                                   # the sequential branch feeds the same heap
    assert np.array_equal(Dc, Ds) and np.array_equal(Ic, Is)           # in the same row order: identical
    D5, I5 = o.search_blas(q[:21], xb[:5], 8)
    assert (I5[:, 5:] == -1).all() and (D5[:, 5:] == -FLT_MAX).all() and np.all(np.diff(D5[:, :5], axis=1) <= 0)


def test_strict_greater_keeps_earlier_row_on_entry_tie():
    xb = np.ones((40, 4), dtype=np.float32)    # all scores equal
    q = np.ones((1, 4), dtype=np.float32)
    D, I = o.search_seq(q, xb, 8)
    assert sorted(I[0].tolist()) == list(range(8))   # later equal rows never displace


def test_normalize_l2():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((3, 768), dtype=np.float32)
    x[1] = 0
    y = x.copy()
    o.normalize_L2(y)
    np.testing.assert_allclose(np.linalg.norm(y[[0, 2]], axis=1), 1.0, atol=1e-6)
    assert (y[1] == 0).all()

"""CPU-only checks: the C-ABI library loads and exports every symbol the public
header declares (no compute without a GPU), fails loudly without a device, and
the host-side sharding / gather / decode logic is right (world_size-2 gloo)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_header_symbols():
    import __graft_entry__ as entry
    entry.build_cuda()
    from ragroute_b200 import _lib
    lib = _lib.load()
    declared = _lib.header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ragroute_b200.h but not exported"
    assert set(declared) == set(_lib._SIGNATURES), "ctypes signatures out of sync with the header"
    assert lib.rr_abi_version() == 1


def test_built_for_sm100a_with_tcgen05_and_tma():
    import __graft_entry__ as entry
    lib = entry.build_cuda()
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):    # tcgen05.mma, TMA load, tcgen05.ld
        assert mnemonic in sass, f"{mnemonic} missing from SASS"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_fails_loudly():
    from ragroute_b200 import _lib
    from ragroute_b200.faiss_compat import IndexFlatIP
    from ragroute_b200.router import FusedCorpusRouter
    with pytest.raises(_lib.RRError, match="no CPU fallback"):
        IndexFlatIP(768)
    with pytest.raises(_lib.RRError, match="no CPU fallback"):
        FusedCorpusRouter(1540)


def test_shard_rows_cover_exactly():
    from ragroute_b200.federation import shard_rows
    for n in (0, 1, 7, 125_000, 23_900_000, 29_900_001):
        for w in (1, 2, 4, 8):
            spans = [shard_rows(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= -(-n // w)


def test_decode_pos():
    from ragroute_b200.federation import decode_pos
    pos = torch.tensor([[0, 31, 32, 127, 128, 300, -1]], dtype=torch.int32)
    rank, src = decode_pos(pos, n_sources=4, k=32)
    assert rank.tolist() == [[0, 0, 0, 0, 1, 2, -1]]
    assert src.tolist() == [[0, 0, 1, 3, 0, 1, -1]]


_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {repo!r})
from ragroute_b200.federation import gather_lists, decode_pos, shard_rows, PackedLists
from ragroute_b200 import synthetic as syn
from oracle import flat_ip, rerank_ref
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, W = dist.get_rank(), 2
d, k, B, S = 64, 8, 5, 2
corpora = [syn.corpus_rows(s, 1001 + 37 * s, d) for s in range(S)]
q = syn.queries(B, d, [0, 1], seed=1)
D_loc = np.empty((B, S * k), np.float32); I_loc = np.empty((B, S * k), np.int64)
for s in range(S):
    lo, hi = shard_rows(len(corpora[s]), W, rank)
    Ds, Is = flat_ip.search_seq(q, corpora[s][lo:hi], k)       # CPU stand-in for the shard-local search
    D_loc[:, s*k:(s+1)*k] = Ds; I_loc[:, s*k:(s+1)*k] = np.where(Is >= 0, Is + lo, -1)
D_cat, I_cat = gather_lists(torch.from_numpy(D_loc), torch.from_numpy(I_loc))
assert D_cat.shape == (B, W * S * k)
# the packed block the GPU path sends (one collective, merge kernel reads the blocks in place)
pl = PackedLists(B, S * k, "cpu")
pl.scores().copy_(torch.from_numpy(D_loc)); pl.ids().copy_(torch.from_numpy(I_loc))
g = pl.gather()
assert g.n_blocks == W and g.block_bytes % 16 == 0 and g.n_scores % 16 == 0
assert torch.equal(torch.cat([g.scores(r) for r in range(W)], dim=1), D_cat)
assert torch.equal(torch.cat([g.ids(r) for r in range(W)], dim=1), I_cat)
ms, mi, mp = rerank_ref.merge_topk(D_cat.numpy(), I_cat.numpy(), k)
r, src = decode_pos(torch.from_numpy(mp), S, k)
# unsharded truth
all_s = np.concatenate([flat_ip.search_seq(q, corpora[s], k)[0] for s in range(S)], axis=1)
all_i = np.concatenate([flat_ip.search_seq(q, corpora[s], k)[1] for s in range(S)], axis=1)
ts, ti, tp = rerank_ref.merge_topk(all_s, all_i, k)
assert np.array_equal(ms, ts) and np.array_equal(mi, ti), "sharded merge differs from unsharded"
assert np.array_equal(src.numpy(), tp // k)
for b in range(B):
    for j in range(k):
        lo, hi = shard_rows(len(corpora[src[b, j]]), W, int(r[b, j]))
        assert lo <= mi[b, j] < hi
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_sharded_gather_merge_world_size_2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(repo=REPO, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


from golden_util import build_abi_check as _build_abi_check


def test_c_abi_is_callable_from_plain_c(tmp_path):
    """The header compiles as C99 and the library links into a plain C program (no CUDA / torch types in
    the boundary); without a GPU every compute entry point refuses loudly."""
    exe = _build_abi_check(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "abi_check ok" in res.stdout


def test_packed_lists_merge_out_layout_and_placement():
    """The byte blocks the step moves around: views must tile their block exactly (16-byte aligned parts, no
    overlap), because the merge kernel and the single device-to-host copy address them by offset."""
    from ragroute_b200.federation import MergeOut, PackedLists, place_sources
    for B, L, W in ((5, 12, 3), (1, 128, 1), (7, 130, 2), (64, 128, 8)):
        pl = PackedLists(B, L, "cpu", n_blocks=W)
        assert pl.n_scores % 16 == 0 and pl.block_bytes % 16 == 0 and pl.buf.numel() == W * pl.block_bytes
        pl.buf.fill_(0)
        for r in range(W):
            pl.scores(r).fill_(float(r + 1))
            pl.ids(r).fill_(-(r + 1))
        for r in range(W):
            assert (pl.scores(r) == float(r + 1)).all() and (pl.ids(r) == -(r + 1)).all()
            assert pl.scores(r).data_ptr() == pl.buf.data_ptr() + r * pl.block_bytes
            assert pl.ids(r).data_ptr() == pl.buf.data_ptr() + r * pl.block_bytes + pl.n_scores
            assert pl.ids(r).data_ptr() % 8 == 0
    mo = MergeOut(3, 10, "cpu")
    n = 30
    assert mo.buf.numel() == 24 * n and mo.host_bytes == 16 * n
    ptrs = [mo.ids.data_ptr(), mo.scores.data_ptr(), mo.source.data_ptr(), mo.rank.data_ptr(), mo.pos.data_ptr()]
    base = mo.buf.data_ptr()
    assert [p - base for p in ptrs] == [0, 8 * n, 12 * n, 16 * n, 20 * n]
    names = [f"s{i}" for i in range(13)]
    where = place_sources(names, 8)
    assert [where[n_] for n_ in names] == [i % 8 for i in range(13)]
    assert set(place_sources(names, 1).values()) == {0}

"""Host-side logic that needs no GPU: the micro-batching collector and the
schema of bench.py's reference arm (the oracle timed as the reference's CPU stand-in)."""
import asyncio
import json
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_micro_batcher_batches_orders_and_flushes():
    from ragroute_b200.collector import MicroBatcher
    seen = []

    def batch_fn(items):
        seen.append(list(items))
        return [x * 10 for x in items]

    async def main():
        mb = MicroBatcher(batch_fn, max_batch=4, max_wait_s=0.05)
        # 10 concurrent requests -> batches of at most 4, results matched to their request
        res = await asyncio.gather(*[mb.submit(i) for i in range(10)])
        assert res == [i * 10 for i in range(10)]
        assert [len(b) for b in seen] == [4, 4, 2]
        assert sum(seen, []) == list(range(10))          # submission order kept
        # a lone request is flushed by the timeout, not held for a full batch
        t0 = asyncio.get_running_loop().time()
        assert await mb.submit(7) == 70
        assert asyncio.get_running_loop().time() - t0 < 1.0
        assert mb.batches == 4 and mb.items == 11 and mb.largest_batch == 4
        await mb.close()
        with pytest.raises(RuntimeError):
            mb.submit(1)

    asyncio.run(main())


def test_micro_batcher_delivers_exceptions_to_the_whole_batch_and_survives():
    from ragroute_b200.collector import MicroBatcher
    calls = []

    def batch_fn(items):
        calls.append(len(items))
        if any(x < 0 for x in items):
            raise ValueError("bad query")          # faiss asserts on a wrong dimension, data_source.py:137-138
        return items

    async def main():
        mb = MicroBatcher(batch_fn, max_batch=8, max_wait_s=0.01)
        futs = [mb.submit(x) for x in (1, -1, 2)]
        out = await asyncio.gather(*futs, return_exceptions=True)
        assert all(isinstance(o, ValueError) for o in out)
        assert await mb.submit(5) == 5              # the batcher keeps serving
        await mb.close()

    asyncio.run(main())


def test_search_rows_splits_a_batched_search_per_request():
    from ragroute_b200.collector import search_rows

    class FakeIndex:                                 # index.search(x, k) contract: D [nq,k], I [nq,k]
        def search(self, x, k):
            assert x.dtype == np.float32 and x.ndim == 2
            s = x.sum(axis=1, keepdims=True)
            return np.repeat(s, k, axis=1).astype(np.float32), np.tile(np.arange(k, dtype=np.int64), (len(x), 1))

    embs = [np.full((1, 8), i, dtype=np.float32) for i in range(3)]     # as data_source.py:113-114 builds them
    out = search_rows(FakeIndex(), embs, 4)
    assert len(out) == 3
    for i, (D, I) in enumerate(out):
        assert D.shape == (1, 4) and I.shape == (1, 4) and D[0, 0] == 8 * i


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """bench.py --impl reference: the oracle (CPU restatement of the reference's algorithm) timed on a
    bounded row sample; one JSON line, same metric/unit/config keys as the GPU arm."""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--batch", "64", "--cpu-sample-rows", "20000", "--workload", "medqa_ts"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=REPO)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["config"]["workload"] == "medqa_ts" and d["config"]["batch"] == 64
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_bench_textbooks_workload_is_one_source_alone_without_the_router():
    """BASELINE configs[0] (Textbooks source alone): both arms default to routing 'all' -- every query searches the
    one source, no router MLP in the step (router.py:226-227) -- while the routed workloads keep 'ragroute'."""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--batch", "24", "--cpu-sample-rows", "4000", "--workload", "textbooks"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=REPO)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.strip()][0])
    assert d["config"]["workload"] == "textbooks" and d["config"]["routing"] == "all"
    assert "mask density 1.00" in d["cpu_baseline"]["sample"] and d["value"] > 0


def test_bench_gpu_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "1",
                        "--workload", "textbooks"], capture_output=True, text=True, timeout=600, cwd=REPO)
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)
    assert not [l for l in p.stdout.splitlines() if l.strip().startswith("{")], "no number may be printed"


def test_router_artifact_files_scaler_and_centroids(tmp_path):
    """File formats of Router.load_router (router.py:121-151), no GPU involved."""
    import pickle
    from sklearn.preprocessing import StandardScaler
    from ragroute_b200 import artifacts
    rng = np.random.default_rng(0)
    scaler = StandardScaler().fit(rng.standard_normal((50, 1540)))
    p = tmp_path / "preprocessed_data.pkl"
    with open(p, "wb") as f:
        pickle.dump(("train", "val", "test", scaler, ["q1"]), f)     # train_medrag_router.py:240-245
    got = artifacts.load_scaler("medrag", str(p))
    assert np.array_equal(got.mean_, scaler.mean_) and np.array_equal(got.scale_, scaler.scale_)
    p2 = tmp_path / "scaler.pkl"
    with open(p2, "wb") as f:
        pickle.dump(scaler, f)
    assert np.array_equal(artifacts.load_scaler("wikipedia", str(p2)).mean_, scaler.mean_)
    assert artifacts.load_scaler("feb4rag", None) is None
    # per-corpus stats (medrag) and the clustered list (wikipedia); short centroids are zero padded
    files = {}
    for c in ("pubmed", "textbooks"):
        files[c] = str(tmp_path / f"{c}_stats.json")
        with open(files[c], "w") as f:
            json.dump({"centroid": rng.standard_normal(700).tolist(), "n": 3}, f)
    cen = artifacts.load_centroids("medrag", ["pubmed", "textbooks"], files, 768)
    assert cen["pubmed"].shape == (768,) and cen["pubmed"].dtype == np.float32 and (cen["pubmed"][700:] == 0).all()
    cl = str(tmp_path / "cluster_stats.json")
    with open(cl, "w") as f:
        json.dump([{"centroid": [float(i)] * 768} for i in range(10)], f)
    cw = artifacts.load_centroids("wikipedia", ["3", "7"], cl, 768)
    assert cw["3"][0] == 3.0 and cw["7"][5] == 7.0
    with pytest.raises(ValueError):
        artifacts.load_centroids("medrag", ["pubmed"], files, 512)


def test_wire_format_is_bit_exact_and_smaller_than_json():
    from ragroute_b200 import wire
    rng = np.random.default_rng(1)
    emb = {"ncbi/MedCPT-Query-Encoder": rng.standard_normal(768).astype(np.float32),
           "e5-large": rng.standard_normal((1, 1024)).astype(np.float32)}
    frames = wire.pack_message({"query_id": "abc", "data_sources": ["pubmed", "textbooks"]}, emb)
    h, got = wire.unpack_message(frames)
    assert h == {"query_id": "abc", "data_sources": ["pubmed", "textbooks"]}
    for k, v in emb.items():
        assert got[k].dtype == np.float32 and got[k].shape == v.shape and np.array_equal(got[k], v)
    as_json = len(json.dumps({k: v.tolist() for k, v in emb.items()}))       # router.py:317-319
    assert sum(map(len, frames)) < as_json / 3
    fr = wire.pack_results("abc", 2, "textbooks", np.arange(32), np.linspace(1, 0, 32, dtype=np.float32), 0.001)
    hh, ids, sc = wire.unpack_results(fr)
    assert hh["k"] == 32 and ids.tolist() == list(range(32)) and sc.dtype == np.float32 and sc[0] == 1.0
    with pytest.raises(ValueError):
        wire.unpack_message([b"nope", b"{}"])

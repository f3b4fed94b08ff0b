"""Latency of the reference's own call, one query at a time (data_source.py:186: index.search(q[1,768], k=32),
numpy in / numpy out), on the Textbooks- and StatPearls-sized sources: this library through
faiss_compat.IndexFlatIP.search against the CPU oracle (FAISS-restated sequential scan; nq = 1 means one
thread in FAISS) on the same rows and queries.  Prints one JSON line per source."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import flat_ip
from ragroute_b200 import synthetic as syn
from ragroute_b200.faiss_compat import IndexFlatIP

flat_ip.build()
k, d = 32, 768
for name, n in (("textbooks", 125_000), ("statpearls", 301_000)):
    sid = syn.MEDRAG_SOURCE_TO_ID[name]
    xb = syn.corpus_rows(sid, n, d)
    qs = syn.queries(300, d, [sid], seed=11)
    ix = IndexFlatIP(d)
    ix.add(xb)
    for i in range(20):
        ix.search(qs[i:i + 1], k)
    t = []
    for i in range(20, 300):
        t0 = time.perf_counter()
        D, I = ix.search(qs[i:i + 1], k)
        t.append((time.perf_counter() - t0) * 1e3)
    tc = []
    same = True
    for i in range(20, 30):
        t0 = time.perf_counter()
        Dr, Ir = flat_ip.search_seq(qs[i:i + 1], xb, k, n_threads=1)
        tc.append((time.perf_counter() - t0) * 1e3)
        Dg, Ig = ix.search(qs[i:i + 1], k)
        same &= bool(np.array_equal(Ig, Ir)) and bool(np.allclose(Dg, Dr, rtol=1e-5))
    print(json.dumps({"source": name, "rows": n, "dim": d, "k": k, "call": "index.search(q[1,768], 32) numpy in/out",
                      "gpu_ms_p50": float(np.percentile(t, 50)), "gpu_ms_p99": float(np.percentile(t, 99)),
                      "cpu_oracle_1thread_ms_p50": float(np.percentile(tc, 50)),
                      "speedup_p50": float(np.percentile(tc, 50) / np.percentile(t, 50)), "same_results": same}))

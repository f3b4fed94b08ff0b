"""Record what the REFERENCE's own data-source glue returns, with `faiss` = the CPU oracle.

Run in the authoring container only (needs /root/reference):

    python tests/golden/gen_glue_fixture.py

/root/reference/ragroute/data_source.py is imported unchanged; the absent third-party module `faiss` is
replaced in sys.modules by a shim whose `read_index` returns an index object backed by the oracle
(oracle/flat_ip.py) over a seeded corpus, so that DataSource.load_faiss_index (:69-80) and
retrieve_docs_medrag / retrieve_docs_fed4rag / retrieve_docs_wikipedia (:143-215) run verbatim: faiss.read_index,
metadatas.jsonl / docids.json / titles+texts loading, index.search, row -> metadata -> chunk text.

The fixture (tests/golden/glue.npz + the seeded generator in tests/glue_util.py) stores the queries' seeds and the
reference's outputs: per query the returned `indices`, `docs` and `scores`.  tests/test_gpu_glue.py rebuilds the same
files on the GPU box, loads them through ragroute_b200.faiss_compat.read_index + ragroute_b200.metadata.DocStore
(what patches/ragroute.patch wires into the reference) and must reproduce them.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, "/root/reference")

from oracle import flat_ip  # noqa: E402
import glue_util  # noqa: E402


class _OracleIndex:
    """faiss.IndexFlatIP as far as data_source.py uses it: .search(x, k) / .search(x, k=k)."""

    def __init__(self, xb):
        self.xb, self.d, self.ntotal = xb, xb.shape[1], xb.shape[0]

    def search(self, x, k):
        return flat_ip.search(np.ascontiguousarray(x, dtype=np.float32), self.xb, k)


def install_faiss_shim():
    m = types.ModuleType("faiss")
    m.read_index = lambda path: _OracleIndex(glue_util.read_flat_file(path))
    m.normalize_L2 = flat_ip.normalize_L2
    sys.modules["faiss"] = m


def main():
    install_faiss_shim()
    import ragroute.config as cfg
    out = {}
    with tempfile.TemporaryDirectory() as root:
        cfg.MEDRAG_DIR, cfg.FEB4RAG_DIR, cfg.WIKIPEDIA_DIR = (os.path.join(root, n) for n in ("medrag", "feb4rag", "wiki"))
        import ragroute.data_source as ds_mod
        ds_mod.MEDRAG_DIR, ds_mod.FEB4RAG_DIR, ds_mod.WIKIPEDIA_DIR = cfg.MEDRAG_DIR, cfg.FEB4RAG_DIR, cfg.WIKIPEDIA_DIR
        for dataset, name in glue_util.CASES:
            glue_util.write_source_files(dataset, name, {"medrag": cfg.MEDRAG_DIR, "feb4rag": cfg.FEB4RAG_DIR,
                                                         "wikipedia": cfg.WIKIPEDIA_DIR}[dataset])
            ds = ds_mod.DataSource(0, dataset, name, simulate=False)   # builds paths, loads the wikipedia side files
            ds.load_faiss_index()                                      # faiss.read_index + metadata, verbatim
            k = cfg.K[dataset]
            qs = glue_util.queries(dataset, name)
            rec = []
            for q in qs:
                emb = np.array(q.tolist(), dtype=np.float32).reshape(1, -1)   # data_source.py:113-114
                if dataset == "medrag":
                    ids, docs, scores = ds.retrieve_docs_medrag(emb, k)
                elif dataset == "feb4rag":
                    ids, docs, scores = ds.retrieve_docs_fed4rag(emb, k)
                else:
                    ids, docs, scores = ds.retrieve_docs_wikipedia(emb, k)
                rec.append({"indices": ids, "docs": docs, "scores": scores})
            out[f"{dataset}/{name}"] = json.dumps(rec)
            ds.context.term()
    np.savez_compressed(os.path.join(HERE, "glue.npz"), **{k: np.frombuffer(v.encode(), dtype=np.uint8) for k, v in out.items()})
    print("wrote", os.path.join(HERE, "glue.npz"), {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()

"""Seeded side files of a data source as the reference lays them out on disk (data_source.py:44-62): the flat
index file, metadatas.jsonl + chunk/*.jsonl (MedRAG), docids.json + corpus.jsonl (FeB4RAG), titles / texts
(Wikipedia).  Shared by tests/golden/gen_glue_fixture.py (which runs the reference's glue over them) and the
tests that replay the recorded outputs."""
import json
import os
import struct

import numpy as np

from ragroute_b200 import synthetic as syn

CASES = [("medrag", "textbooks"), ("feb4rag", "scifact"), ("wikipedia", "3")]
N_ROWS = {"medrag": 4000, "feb4rag": 2500, "wikipedia": 3000}
DIM = {"medrag": 768, "feb4rag": 768, "wikipedia": 768}      # config.py:44-58,93-95: scifact uses gte-base (768)
FEB_MODEL = "gte-base"
N_QUERIES = 6


def corpus(dataset, name):
    x = syn.corpus_rows(40 + len(name), N_ROWS[dataset], DIM[dataset], seed_offset=len(dataset))
    if dataset == "wikipedia":                                  # the reference's wikipedia indexes are normalised
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.ascontiguousarray(x, dtype=np.float32)


def queries(dataset, name):
    rng = np.random.default_rng(len(dataset) * 31 + len(name))
    q = rng.standard_normal((N_QUERIES, DIM[dataset])).astype(np.float32)
    return q * np.float32(3.0 if dataset == "wikipedia" else 1.0 / np.sqrt(DIM[dataset]))


def write_flat_file(path, xb):
    """upstream faiss 1.7.4 IndexFlatIP layout (index_write.cpp)"""
    n, d = xb.shape
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(b"IxFI" + struct.pack("<iqqqBiQ", d, n, 1 << 20, 1 << 20, 1, 0, n * d))
        f.write(xb.tobytes())


def read_flat_file(path):
    with open(path, "rb") as f:
        assert f.read(4) == b"IxFI"
        d, n, _, _, _, metric, count = struct.unpack("<iqqqBiQ", f.read(struct.calcsize("<iqqqBiQ")))
        assert metric == 0 and count == n * d
        return np.frombuffer(f.read(n * d * 4), dtype="<f4").reshape(n, d).copy()


def index_path(dataset, name, root):
    if dataset == "medrag":
        return os.path.join(root, name, "index", "ncbi/MedCPT-Article-Encoder", "faiss.index")
    if dataset == "feb4rag":
        return os.path.join(root, "dataset_creation", "2_search", "embeddings", name, f"{name}_{FEB_MODEL}.faiss")
    return os.path.join(root, "faiss_clusters", "normalized_indexes", f"faiss_index_{name}_normalized.index")


def write_source_files(dataset, name, root):
    xb = corpus(dataset, name)
    n = len(xb)
    write_flat_file(index_path(dataset, name, root), xb)
    if dataset == "medrag":
        files = [f"book{j}" for j in range(7)]
        per = {f: [] for f in files}
        metas = []
        for i in range(n):
            f = files[(i * 7 + i // 13) % 7]
            metas.append({"index": len(per[f]), "source": f})
            per[f].append({"id": f"{f}_{len(per[f])}", "title": f"Title {i}", "content": f"chunk text of row {i}",
                           "contents": f"Title {i}. chunk text of row {i}"})
        idx_dir = os.path.dirname(index_path(dataset, name, root))
        with open(os.path.join(idx_dir, "metadatas.jsonl"), "w") as fh:
            fh.write("\n".join(json.dumps(m) for m in metas) + "\n")
        os.makedirs(os.path.join(root, name, "chunk"), exist_ok=True)
        for f in files:
            with open(os.path.join(root, name, "chunk", f"{f}.jsonl"), "w") as fh:
                fh.write("\n".join(json.dumps(e) for e in per[f]) + "\n")
    elif dataset == "feb4rag":
        docids = [f"doc-{(i * 7919) % n}" for i in range(n)]
        idx_dir = os.path.dirname(index_path(dataset, name, root))
        with open(os.path.join(idx_dir, f"{name}_{FEB_MODEL}.docids.json"), "w") as fh:
            json.dump(docids, fh)
        cdir = os.path.join(root, "dataset_creation/original_dataset", name, name)
        os.makedirs(cdir, exist_ok=True)
        with open(os.path.join(cdir, "corpus.jsonl"), "w") as fh:
            for i in range(n):
                fh.write(json.dumps({"_id": f"doc-{i}", "title": f"T{i}", "text": f"text of document {i}"}) + "\n")
    else:
        sdir = os.path.join(root, "faiss_clusters", "split_texts_titles")
        os.makedirs(sdir, exist_ok=True)
        with open(os.path.join(sdir, f"titles_{name}.txt"), "w", encoding="utf-8") as fh:
            fh.write("\n".join(f"Wiki title {i}" for i in range(n)) + "\n")
        with open(os.path.join(sdir, f"texts_{name}.txt"), "w", encoding="utf-8") as fh:
            fh.write("\n".join(f"wiki passage number {i}" for i in range(n)) + "\n")


def load_golden():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "glue.npz"))
    return {k: json.loads(bytes(g[k]).decode()) for k in g.files}

"""Row-id -> document maps of the reference's data sources.

``faiss.read_index`` returns bare row numbers; the reference turns them into documents with three
side files, loaded next to the index in ``DataSource.load_faiss_index``
(/root/reference/ragroute/data_source.py:69-80) and used in the three ``retrieve_docs_*``:

* MedRAG  (:48, :73, :190, :166-183): ``metadatas.jsonl`` -- one JSON object per index row,
  ``{"index": <line number>, "source": <chunk file stem>}``; the document is line ``index`` of
  ``<dataset_dir>/<corpus>/chunk/<source>.jsonl`` (raw lines cached per file, parsed on demand).
* FeB4RAG (:53, :75-76, :159, :144-156): ``<name>_<model>.docids.json`` -- a JSON list, row -> BEIR
  ``_id``; documents come from ``corpus.jsonl`` keyed by ``_id``.
* Wikipedia (:59-62, :207-211): ``titles_<c>.txt`` / ``texts_<c>.txt``, row -> line.

Pure host-side lookups (no arithmetic); they exist so that ``patches/ragroute.patch`` can return ids + scores
from every source and fetch the text of the k documents that survive the global merge only
(``DocStore.fetch``), instead of every source shipping k full documents of which (S-1)/S are discarded.
Row -1 (FAISS padding when a source has fewer than k rows) maps to ``None``; the reference would silently
return the LAST row for it (``metadatas[-1]``, data_source.py:190 -- a latent bug).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence


def load_medrag_metadatas(path: str) -> List[dict]:
    """data_source.py:73 -- ``[json.loads(line) for line in open(path).read().strip().split('\\n')]``."""
    with open(path, "r") as f:
        text = f.read().strip()
    return [json.loads(line) for line in text.split("\n")] if text else []


def load_feb4rag_docids(path: str) -> List[str]:
    """data_source.py:75-76 -- ``json.load`` of the docids list."""
    with open(path, "r") as f:
        return json.load(f)


def load_lines(path: str) -> List[str]:
    """data_source.py:59-62 -- ``f.read().splitlines()`` (Wikipedia titles / texts)."""
    with open(path, "r", encoding="utf-8") as f:
        return f.read().splitlines()


class DocStore:
    """What a data source needs to turn the row ids of ``index.search`` into the reference's reply
    (``indices``, ``docs``), for one source."""

    def __init__(self, dataset: str, name: str, dataset_dir: str, metadatas: Optional[Sequence] = None,
                 titles: Optional[Sequence[str]] = None, texts: Optional[Sequence[str]] = None):
        if dataset not in ("medrag", "feb4rag", "wikipedia"):
            raise ValueError(f"Unknown dataset: {dataset}")
        self.dataset, self.name, self.dataset_dir = dataset, name, dataset_dir
        self.metadatas = metadatas if metadatas is not None else []
        self.titles, self.texts = titles, texts
        self.cache_jsonl: Dict[str, object] = {}

    # -- row id -> what the reference calls "indices" / "ids" --------------------------------------
    def indices(self, rows: Sequence[int]) -> list:
        """medrag: metadata dicts (data_source.py:190); feb4rag: docid strings (:159); wikipedia: ints (:215)."""
        if self.dataset == "wikipedia":
            return [int(i) for i in rows]
        return [self.metadatas[int(i)] if int(i) >= 0 else None for i in rows]

    # -- "indices" -> documents -----------------------------------------------------------------------
    def fetch(self, indices: Sequence) -> list:
        """idx2txt of the three retrieve_docs_* (data_source.py:144-156, :166-183, :207-211)."""
        if self.dataset == "medrag":
            out = []
            for i in indices:
                if i is None:
                    out.append(None)
                    continue
                source, index = i["source"], i["index"]
                if source not in self.cache_jsonl:  # raw lines cached as strings, parsed when needed (:173-181)
                    path = os.path.join(self.dataset_dir, self.name, "chunk", f"{source}.jsonl")
                    with open(path, "r") as f:
                        self.cache_jsonl[source] = f.read().strip().split("\n")
                out.append(json.loads(self.cache_jsonl[source][index]))
            return out
        if self.dataset == "feb4rag":
            if self.name not in self.cache_jsonl:  # :145-153
                path = os.path.join(self.dataset_dir, "dataset_creation/original_dataset", self.name, self.name, "corpus.jsonl")
                corpus = {}
                with open(path, "r") as f:
                    for line in f:
                        entry = json.loads(line)
                        corpus[entry["_id"]] = entry
                self.cache_jsonl[self.name] = corpus
            corpus = self.cache_jsonl[self.name]
            return [corpus.get(doc_id, None) for doc_id in indices]
        return [(self.titles[i], self.texts[i]) if i >= 0 else None for i in indices]

    def lookup(self, rows: Sequence[int]):
        """rows -> (indices, docs), the first two members of the reference's reply (data_source.py:123-132)."""
        ind = self.indices(rows)
        return ind, self.fetch(ind)

"""Compact wire format for the router -> server -> data-source hops.

The reference ships every query embedding across two process hops as a JSON list of floats
(/root/reference/ragroute/router.py:317-332 ``embedding.tolist()``; http_server.py:205-209;
parsed back with ``np.array(..., float32)`` at data_source.py:113-114): about 15 KB of text per
768-d embedding, and a decimal round trip.  These helpers carry the same message as a small JSON
header plus the raw little-endian fp32 bytes (3 KB per 768-d embedding, bit-exact), suitable
for a ZeroMQ multipart message (``send_multipart(frames)``) or any byte pipe.  Pure host-side
plumbing; opt-in, the JSON path keeps working.
"""
from __future__ import annotations

import json
from typing import Dict, List, Tuple

import numpy as np

MAGIC = b"RRW1"


def pack_message(header: dict, embeddings: Dict[str, np.ndarray]) -> List[bytes]:
    """-> frames [MAGIC, header-json, emb_0, emb_1, ...]; the header gains the embedding index
    (name, dtype, shape per frame) so that any other key of the reference's message
    (``id``, ``query_id``, ``data_sources``, ...) travels unchanged."""
    names, frames = [], []
    for name, e in embeddings.items():
        a = np.ascontiguousarray(e, dtype="<f4")
        names.append({"name": name, "shape": list(a.shape)})
        frames.append(a.tobytes())
    h = dict(header)
    h["_embeddings"] = names
    return [MAGIC, json.dumps(h).encode("utf-8")] + frames


def unpack_message(frames: List[bytes]) -> Tuple[dict, Dict[str, np.ndarray]]:
    if len(frames) < 2 or frames[0] != MAGIC:
        raise ValueError("not a ragroute_b200 wire message")
    h = json.loads(frames[1].decode("utf-8"))
    idx = h.pop("_embeddings")
    if len(idx) != len(frames) - 2:
        raise ValueError(f"header lists {len(idx)} embeddings, message carries {len(frames) - 2}")
    out = {}
    for meta, raw in zip(idx, frames[2:]):
        a = np.frombuffer(raw, dtype="<f4")
        if a.size != int(np.prod(meta["shape"])):
            raise ValueError(f"embedding {meta['name']!r}: {a.size} floats for shape {meta['shape']}")
        out[meta["name"]] = a.reshape(meta["shape"])
    return h, out


def decode_request(frames: List[bytes]) -> Tuple[dict, "np.ndarray | None"]:
    """What a data source receives (data_source.py:102,113-114): either the frames of ``pack_message`` with one
    embedding named "embedding", or the reference's single JSON frame with ``"embedding": [floats]`` (also the
    "fetch" message of the second phase, which carries no embedding).  -> (message dict, fp32 embedding or None)."""
    if frames and frames[0] == MAGIC:
        header, embs = unpack_message(frames)
        return header, embs.get("embedding")
    msg = json.loads(frames[0].decode("utf-8"))
    emb = msg.get("embedding")
    return msg, (np.asarray(emb, dtype=np.float32) if emb is not None else None)


def decode_routing(frames: List[bytes]) -> Tuple[dict, Dict[str, np.ndarray]]:
    """What the server receives from the router (http_server.py:174-179): ``pack_message`` frames with one
    embedding per encoder, or the reference's JSON with ``"embeddings": {encoder: [floats]}``."""
    if frames and frames[0] == MAGIC:
        return unpack_message(frames)
    msg = json.loads(frames[0].decode("utf-8"))
    embs = {m: np.asarray(e, dtype=np.float32) for m, e in msg.pop("embeddings", {}).items()}
    return msg, embs


def pack_results(query_id, client_id, name: str, ids: np.ndarray, scores: np.ndarray, duration: float) -> List[bytes]:
    """A data source's reply without document text (ids + scores only): the server fetches the text of the
    k documents that survive the global merge instead of receiving S*k documents of which (S-1)/S are
    discarded (data_source.py:123-132, http_server.py:280-293)."""
    i = np.ascontiguousarray(ids, dtype="<i8")
    s = np.ascontiguousarray(scores, dtype="<f4")
    h = {"query_id": query_id, "client_id": client_id, "name": name, "duration": duration, "k": int(i.size)}
    return [MAGIC, json.dumps(h).encode("utf-8"), i.tobytes(), s.tobytes()]


def unpack_results(frames: List[bytes]) -> Tuple[dict, np.ndarray, np.ndarray]:
    if len(frames) != 4 or frames[0] != MAGIC:
        raise ValueError("not a ragroute_b200 result message")
    h = json.loads(frames[1].decode("utf-8"))
    return h, np.frombuffer(frames[2], dtype="<i8"), np.frombuffer(frames[3], dtype="<f4")

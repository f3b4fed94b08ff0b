"""Micro-batching collector for the reference's serve loops.

The reference handles strictly one query at a time: the data-source loop awaits one
``recv_json`` and runs one ``index.search`` on a ``[1, d]`` array
(/root/reference/ragroute/data_source.py:99-132), the router dequeues one query per
iteration and otherwise sleeps 100 ms (/root/reference/ragroute/router.py:207-219), so
``--parallel 64`` (run_benchmark.py:90-114) never forms a batch and the GPU path would only
ever see B = 1.  ``MicroBatcher`` sits between such a loop and the batched entry points of
this package (``IndexFlatIP.search``, ``FusedCorpusRouter.select``, ``Federation.search_host``):

    batcher = MicroBatcher(lambda embs: search_rows(index, embs, k), max_batch=64, max_wait_s=300e-6)
    ...
    query_data = await self.receiver.recv_json()                  # unchanged reference code
    ids, scores = await batcher.submit(np.asarray(query_data["embedding"], np.float32))

Requests arriving while a batch is being processed are collected and go out together as the
next batch; a lone request waits at most ``max_wait_s``.  Results come back per request, in
submission order inside a batch; an exception raised by ``batch_fn`` is delivered to every
request of that batch (the reference's loop logs it and sends no reply, data_source.py:137-138).

Pure host-side plumbing: no arithmetic happens here.
"""
from __future__ import annotations

import asyncio
import time
from typing import Any, Awaitable, Callable, List, Optional, Sequence, Tuple

import numpy as np


class MicroBatcher:
    """Collect single requests into batches for ``batch_fn(items) -> results`` (one result per item)."""

    def __init__(self, batch_fn: Callable[[List[Any]], Sequence[Any]], max_batch: int = 64,
                 max_wait_s: float = 300e-6, run_in_executor: bool = False):
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self.batch_fn = batch_fn
        self.max_batch = int(max_batch)
        self.max_wait_s = float(max_wait_s)
        self.run_in_executor = run_in_executor
        self._pending: List[Tuple[Any, asyncio.Future]] = []
        self._wakeup: Optional[asyncio.Event] = None
        self._task: Optional[asyncio.Task] = None
        self._closed = False
        # counters (the reference logs per-query durations; these expose what batching achieved)
        self.batches = 0
        self.items = 0
        self.largest_batch = 0

    # -- public API ---------------------------------------------------------------------------
    def submit(self, item: Any) -> Awaitable[Any]:
        """Queue one request; the returned awaitable resolves to its result."""
        if self._closed:
            raise RuntimeError("MicroBatcher is closed")
        loop = asyncio.get_running_loop()
        if self._task is None:
            self._wakeup = asyncio.Event()
            self._task = loop.create_task(self._run())
        fut = loop.create_future()
        self._pending.append((item, fut))
        self._wakeup.set()
        return fut

    async def close(self) -> None:
        """Flush what is pending and stop the background task."""
        self._closed = True
        if self._task is not None:
            self._wakeup.set()
            await self._task
            self._task = None

    @property
    def mean_batch(self) -> float:
        return self.items / self.batches if self.batches else 0.0

    # -- background task ------------------------------------------------------------------------
    async def _run(self) -> None:
        loop = asyncio.get_running_loop()
        while True:
            if not self._pending:
                if self._closed:
                    return
                self._wakeup.clear()
                await self._wakeup.wait()
                continue
            # a batch is open: give later arrivals max_wait_s to join, unless it is already full
            deadline = time.perf_counter() + self.max_wait_s
            while len(self._pending) < self.max_batch and not self._closed:
                left = deadline - time.perf_counter()
                if left <= 0:
                    break
                self._wakeup.clear()
                try:
                    await asyncio.wait_for(self._wakeup.wait(), timeout=left)
                except asyncio.TimeoutError:
                    break
            batch, self._pending = self._pending[:self.max_batch], self._pending[self.max_batch:]
            items = [it for it, _ in batch]
            try:
                if self.run_in_executor:
                    results = await loop.run_in_executor(None, self.batch_fn, items)
                else:
                    results = self.batch_fn(items)
                if len(results) != len(items):
                    raise RuntimeError(f"batch_fn returned {len(results)} results for {len(items)} items")
            except Exception as e:  # noqa: BLE001 -- delivered to every waiter of the batch
                for _, fut in batch:
                    if not fut.done():
                        fut.set_exception(e)
            else:
                for (_, fut), res in zip(batch, results):
                    if not fut.done():
                        fut.set_result(res)
            self.batches += 1
            self.items += len(items)
            self.largest_batch = max(self.largest_batch, len(items))
            await asyncio.sleep(0)  # let the waiters run (and submit their next request)


# -- batch functions for the three call sites ---------------------------------------------------
def search_rows(index, embeddings: Sequence[np.ndarray], k: int, normalize: bool = False):
    """batch_fn for a data source: embeddings as the serve loop builds them
    (``np.array(..., float32).reshape(1, -1)``, data_source.py:113-114) -> per request
    ``(D[1, k], I[1, k])`` exactly as ``index.search(embedding, k)`` returns them.
    normalize: ``faiss.normalize_L2`` first, in the same call (retrieve_docs_wikipedia, data_source.py:199)."""
    x = np.concatenate([np.asarray(e, dtype=np.float32).reshape(1, -1) for e in embeddings], axis=0)
    D, I = index.search(x, k, normalize=True) if normalize else index.search(x, k)
    return [(D[i:i + 1], I[i:i + 1]) for i in range(len(embeddings))]


def select_rows(router, query_embeddings: Sequence[dict], threshold: float):
    """batch_fn for the router: per request the dict encoder-name -> embedding that
    ``select_relevant_sources_ragroute`` takes (router.py:241) -> per request the list of
    selected source names, in ``data_sources`` order (router.py:282)."""
    import torch
    slots = router.slot_names
    q = np.zeros((len(query_embeddings), len(slots), router.d_max), dtype=np.float32)
    for b, emb in enumerate(query_embeddings):
        for s, m in enumerate(slots):
            e = np.asarray(emb[m], dtype=np.float32).reshape(-1)
            q[b, s, :len(e)] = e  # np.pad(query_embed, (0, d_max - len)), router.py:248
    mask = router.select(torch.from_numpy(q).cuda(router.device), threshold).cpu().numpy().astype(bool)
    return [[c for p, c in zip(row, router.data_sources) if p] for row in mask]

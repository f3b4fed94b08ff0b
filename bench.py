#!/usr/bin/env python
"""bench.py -- routed top-k retrieval throughput on synthetic corpora of the reference's shapes.

One "step" = one pass of the hot path over one batch of B queries:
router MLP (ragroute/router.py:241-283) -> masked exact inner-product top-k in every
source (data_source.py:186) -> [all-gather of the per-rank lists when N > 1] ->
cross-source merge (rerank.py:3-9).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--workload NAME] [--corpus iid|clustered]
    python bench.py --impl reference ...      # the reference's CPU algorithm on host cores

Workloads (BASELINE.json configs): federation = configs[3] (default: the configuration the metric is
quoted on, 54.226 M x 768, it fits one B200), pubmed = configs[2], medqa_ts = configs[1],
textbooks = configs[0], feb4rag16 = configs[4] (16 mixed-dimension sources placed whole on the ranks,
64 queries in flight through the micro-batching collector).

N > 1 is launched by torchrun (one rank per GPU, NCCL); MedRAG corpora are row-sharded over the
ranks, feb4rag16 sources are placed whole (strong scaling: the federation is fixed).
Rank 0 prints ONE JSON line on stdout; everything else goes to stderr.

Every line carries "verify": after the timed region a few queries of the batch are re-done by an
exhaustive float64 scan over the stored rows of every rank (reconstructed bit-exactly from the planes)
and compared with what the timed path returned, under the north_star tolerance.

The oracle (oracle/) is imported only by the cpu_baseline leg and by --impl reference,
where it is the thing being timed as the reference's stand-in, never by the GPU arm.
"""
from __future__ import annotations

import os

# the CPU legs alternate OpenBLAS sgemm (its own threads) with an OpenMP heap pass: a spinning OpenMP team
# starves the BLAS threads (measured 4.5x slower); must be set before libgomp starts
os.environ.setdefault("OMP_WAIT_POLICY", "passive")

import argparse  # noqa: E402
import datetime  # noqa: E402
import gc  # noqa: E402
import json  # noqa: E402
import sys  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402
from dataclasses import dataclass, field  # noqa: E402
from typing import Dict, List  # noqa: E402

import numpy as np  # noqa: E402

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "routed top-32 queries/sec (MedRAG-shaped synthetic federation, exact IP search)"
METRIC_FEB = "routed top-10 queries/sec (16 mixed-dimension synthetic sources, exact IP search, 64 queries in flight)"
UNIT = "queries/s"

# global row counts per MedRAG workload, in the reference's data_sources order (config.py:33)
WORKLOADS = {
    # BASELINE.json configs[3]: the configuration the metric is quoted on
    "federation": {"pubmed": 23_900_000, "statpearls": 301_000, "textbooks": 125_000, "wikipedia": 29_900_000},
    "pubmed": {"pubmed": 23_900_000},                                   # configs[2]
    "medqa_ts": {"statpearls": 301_000, "textbooks": 125_000},          # configs[1]
    "textbooks": {"textbooks": 125_000},                                # configs[0]
    "feb4rag16": None,                                                  # configs[4] (ragroute_b200/workloads.py)
}
DEFAULT_SWEEP = {"federation": "1,8,64,256", "pubmed": "1,4,16,64,256,1024,4096", "medqa_ts": "1,8,64,256",
                 "textbooks": "1,8,64,256", "feb4rag16": "1,8,256,2048"}
BYTES_PER_ELEM_STORED = 4  # two 16-bit planes
MEDRAG_ENCODER = "ncbi/MedCPT-Query-Encoder"  # config.py:38-43


_REAL_STDOUT = None


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ------------------------------------------------------------------------- workload spec
@dataclass
class Spec:
    """What both arms need to know about a workload (shapes only; no data)."""
    workload: str
    dataset: str
    names: List[str]
    dims: Dict[str, int]
    rows: Dict[str, int]           # global rows per source
    model: Dict[str, str]          # query encoder of each source
    gen_id: Dict[str, int]         # generator id of each source (ragroute_b200/synthetic.py)
    onehot: Dict[str, int]
    n_onehot: int
    d_max: int
    k: int
    scaler: bool
    threshold: float
    router_seed: int
    placement: str                 # "row-shard" | "whole-source"
    slots: List[str] = field(default_factory=list)   # encoder names, router slot order (sorted)

    @property
    def in_dim(self) -> int:
        return 2 * self.d_max + self.n_onehot

    def d_pad(self, n: str) -> int:
        return (self.dims[n] + 63) // 64 * 64

    def query_bytes(self) -> int:
        return sum(4 * int(m[3:]) if m.startswith("enc") else 4 * self.d_max for m in self.slots)


def make_spec(workload: str, rows_scale: float) -> Spec:
    from ragroute_b200 import synthetic as syn
    if workload == "feb4rag16":
        from ragroute_b200.workloads import FEB16_DMAX, FEB16_K, feb4rag16_layout
        names, dims, rows, model = feb4rag16_layout(rows_scale)
        return Spec(workload, "feb4rag", names, dims, rows, model, {n: 100 + i for i, n in enumerate(names)},
                    {n: i for i, n in enumerate(names)}, len(names), FEB16_DMAX, FEB16_K, False, 0.5, 9,
                    "whole-source", sorted(set(model.values())))
    rows = {n: max(syn.MEDRAG_K, int(r * rows_scale)) for n, r in WORKLOADS[workload].items()}
    names = list(rows)
    return Spec(workload, "medrag", names, {n: syn.MEDRAG_DIM for n in names}, rows,
                {n: MEDRAG_ENCODER for n in names}, {n: syn.MEDRAG_SOURCE_TO_ID[n] for n in names},
                {n: syn.MEDRAG_SOURCE_TO_ID[n] for n in names}, len(syn.MEDRAG_SOURCE_TO_ID), syn.MEDRAG_DIM,
                syn.MEDRAG_K, True, 0.4924, 0, "row-shard", [MEDRAG_ENCODER])


def host_queries(spec: Spec, batch: int, seed: int) -> Dict[str, np.ndarray]:
    """One embedding per encoder: encoder name -> fp32 [batch, d]."""
    from ragroute_b200 import synthetic as syn
    if spec.workload == "feb4rag16":
        from ragroute_b200.workloads import feb4rag16_queries
        return feb4rag16_queries(batch, seed, spec.names, spec.dims)
    return {MEDRAG_ENCODER: syn.queries(batch, spec.d_max, [spec.gen_id[n] for n in spec.names], seed=seed)}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                phys = int(vis.split(",")[index])
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            log(f"[bench] NVML unavailable ({e}); clocks will be null")
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:  # noqa: BLE001
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for b, name in self.REASONS.items():
                    if bits & b and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def start(self):
        if self._nv is not None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------- CPU reference arm
def cpu_sample_plan(rows: dict, sample_rows_total: int):
    total = sum(rows.values())
    plan = {}
    for n, r in rows.items():
        plan[n] = max(min(r, 1024), int(round(sample_rows_total * r / total)))
        plan[n] = min(plan[n], r)
    return plan


class CpuReference:
    """The reference's CPU algorithm for the same step on host cores: router MLP per query (numpy
    restatement of router.py:241-283), the FAISS-restated flat-IP search of every ROUTED (query, source)
    pair (oracle/flat_ip*.{c,py}), numpy argsort merge (rerank.py:3-9).  Every source is represented by a
    ROW SAMPLE (proportional, >= 1 M rows in total by default: far beyond the caches); the reported
    queries/s is the measured rate scaled by sample_rows / total_rows (a flat scan is linear in N)."""

    def __init__(self, spec: Spec, sample_rows_total: int, routing: str = "ragroute"):
        from oracle import flat_ip, rerank_ref, router_ref
        from ragroute_b200 import synthetic as syn
        flat_ip.build()
        self.flat_ip, self.rerank_ref, self.router_ref = flat_ip, rerank_ref, router_ref
        self.spec, self.routing = spec, routing
        self.cores = os.cpu_count() or 1
        self.plan = cpu_sample_plan(spec.rows, sample_rows_total)
        t0 = time.time()
        self.corpora = {n: syn.corpus_rows(spec.gen_id[n], self.plan[n], spec.dims[n]) for n in spec.names}
        self.gen_s = time.time() - t0
        self.frac = sum(self.plan.values()) / float(sum(spec.rows.values()))
        self.sd = syn.router_state_dict(spec.in_dim, spec.router_seed)
        self.scaler = syn.router_scaler(spec.in_dim, spec.router_seed) if spec.scaler else (None, None)
        self.centroids = {n: np.pad(syn.source_mean(spec.gen_id[n], spec.dims[n]), (0, spec.d_max - spec.dims[n]))
                          for n in spec.names}
        # same calibration as the GPU arm: median logit at the decision threshold
        qc = host_queries(spec, 64, 4242)
        lg = [self._route_one({m: qc[m][b] for m in qc})[1] for b in range(64)]
        self.sd = dict(self.sd)
        shift = router_ref.logit_threshold(spec.threshold) - float(np.median(np.concatenate(lg)))
        self.sd["fc3.bias"] = (self.sd["fc3.bias"] + np.float32(shift)).astype(np.float32)

    def _route_one(self, q_by_model):
        s, rr = self.spec, self.router_ref
        f = rr.build_features(q_by_model, s.names, s.model, self.centroids, s.onehot, s.n_onehot, s.d_max)
        return rr.select(self.sd, f, s.threshold, self.scaler[0], self.scaler[1])

    def step(self, q: Dict[str, np.ndarray], mode: str, n_threads: int):
        """mode 'batch': FAISS given the whole batch per source (nq >= 20 -> blocked sgemm + heap pass, all
        threads; nq < 20 -> sequential scan, OpenMP over queries).  mode 'deployed': the reference as it
        runs, one query at a time on one thread (data_source.py:114,186)."""
        s = self.spec
        B = next(iter(q.values())).shape[0]
        mask = np.ones((B, len(s.names)), dtype=bool)  # routing "all": every source, no router (router.py:226-227)
        for b in range(B if self.routing == "ragroute" else 0):  # the reference routes one query at a time (router.py:212-214)
            mask[b] = self._route_one({m: q[m][b] for m in q})[0]
        per_src = {}
        for si, n in enumerate(s.names):
            sel = np.nonzero(mask[:, si])[0]
            if not len(sel):
                continue
            qs = q[s.model[n]][sel]
            if mode == "deployed":
                Ds = np.empty((len(sel), s.k), np.float32)
                Is = np.empty((len(sel), s.k), np.int64)
                for j in range(len(sel)):
                    Ds[j:j + 1], Is[j:j + 1] = self.flat_ip.search_seq(qs[j:j + 1], self.corpora[n], s.k, n_threads=1)
            else:
                Ds, Is = self.flat_ip.search(qs, self.corpora[n], s.k, n_threads=n_threads)
            per_src[si] = (sel, Ds, Is)
        out = []
        for b in range(B):
            sc, key = [], []
            for si, (sel, Ds, Is) in per_src.items():
                j = np.searchsorted(sel, b)
                if j < len(sel) and sel[j] == b:
                    sc += Ds[j].tolist()
                    key += [(si, int(i)) for i in Is[j]]
            out.append(self.rerank_ref.rerank_medrag(key, sc, s.k))
        return out, float(mask.mean())

    def run(self, batch: int, steps: int, warmup: int, mode: str, n_threads: int = 0):
        """-> (queries/s scaled to the full workload, ms per step on the sample, mask density)"""
        nt = n_threads if n_threads > 0 else self.cores
        times, dens = [], []
        for it in range(warmup + steps):
            q = host_queries(self.spec, batch, 42 + it)
            t0 = time.perf_counter()
            _, dn = self.step(q, mode, nt)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
                dens.append(dn)
        t_step = float(np.mean(times))
        return batch / t_step * self.frac, t_step * 1e3, float(np.mean(dens))

    def sample_text(self) -> str:
        return (f"{sum(self.plan.values())} of {sum(self.spec.rows.values())} rows ({self.frac:.4%}, proportional per "
                f"source); queries/s scaled linearly in rows to the full workload")


def cpu_baseline_legs(spec: Spec, batch: int, sample_rows: int, steps: int, warmup: int, deployed_queries: int = 16,
                      routing: str = "ragroute"):
    """Both CPU figures: FAISS given the same batch (all cores) and the reference as
    deployed (nq = 1, one thread)."""
    t0 = time.time()
    ref = CpuReference(spec, sample_rows, routing)
    qps_b, ms_b, dens_b = ref.run(batch, steps, warmup, "batch")
    qps_d, ms_d, dens_d = ref.run(min(batch, deployed_queries), 1, 0, "deployed", n_threads=1)
    nq_note = ("blocked sgemm (OpenBLAS) + heap pass, FAISS' nq >= 20 branch" if batch * dens_b >= 20
               else "sequential scan, OpenMP over queries, FAISS' nq < 20 branch")
    same = {"value": qps_b, "unit": UNIT, "cores": ref.cores, "kind": "port",
            "sample": f"{ref.sample_text()}; batch {batch}, {steps} timed steps, routed pairs only (mask density "
                      f"{dens_b:.2f}), {nq_note}; {ms_b:.0f} ms per sampled step"}
    dep = {"value": qps_d, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"{ref.sample_text()}; {min(batch, deployed_queries)} queries one at a time on one thread "
                     f"(index.search(q[1,d], k) as data_source.py:114,186 calls it), mask density {dens_d:.2f}"}
    return {**same, "faiss_same_batch": same, "as_deployed": dep, "wall_s": time.time() - t0,
            "corpus_generation_s": ref.gen_s}


def reference_main(args, spec: Spec):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ref = CpuReference(spec, args.cpu_sample_rows, args.routing)
    qps, ms_sample, dens = ref.run(args.batch, args.steps, args.warmup, "batch")
    sample = (f"{ref.sample_text()}; batch {args.batch}, FAISS-restated search of routed pairs only (mask density "
              f"{dens:.2f}) with FAISS' own dispatch (nq >= 20: blocked sgemm + heap pass; below: sequential scan), "
              f"all {ref.cores} host threads")
    line = {
        "impl": "reference", "metric": METRIC_FEB if spec.workload == "feb4rag16" else METRIC, "value": qps, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_sample / ref.frac,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": spec.workload, "rows": spec.rows, "dim": spec.d_max, "k": spec.k, "batch": args.batch,
                   "routing": args.routing},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": ref.cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------- GPU arm
def pct(xs, p):
    return float(np.percentile(np.asarray(xs, dtype=np.float64), p)) if len(xs) else None


class GpuWorkload:
    """This rank's part of a workload on its GPU plus the query plumbing of both timed paths."""

    def __init__(self, spec: Spec, rank: int, world: int, local_rank: int, routing: str, corpus: str):
        import torch
        from ragroute_b200 import workloads as wl
        self.spec, self.rank, self.world, self.corpus = spec, rank, world, corpus
        self.dev = torch.device("cuda", local_rank)
        if spec.workload == "feb4rag16":
            self.fed, layout = wl.build_feb4rag16_federation(rank=rank, world=world, device=local_rank,
                                                             rows_scale=1.0, routing=routing)
            # (rows_scale is already applied in the spec: rebuild with the spec's rows if it was scaled)
            self.local_rows = {n: (spec.rows[n] if layout["placement"][n] == rank else 0) for n in spec.names}
            assert layout["rows"] == spec.rows, "feb4rag16: --rows-scale is applied through make_spec only at 1.0"
        else:
            self.fed, self.local_rows = wl.build_medrag_federation(spec.rows, d=spec.d_max, k=spec.k, rank=rank,
                                                                   world=world, routing=routing, device=local_rank,
                                                                   corpus=corpus)
        self.multi = spec.workload == "feb4rag16"
        torch.cuda.synchronize(self.dev)

    def close(self):
        import torch
        for ix in self.fed.indexes:
            ix.close()
        self.fed = None
        gc.collect()
        torch.cuda.empty_cache()

    # -- queries -------------------------------------------------------------------------------
    def pinned_queries(self, batch: int, seed: int):
        """Host queries of one step as pinned tensors, one per encoder slot (a bare tensor for MedRAG)."""
        import torch
        if self.corpus == "clustered":
            from ragroute_b200.workloads import clustered_queries
            q = clustered_queries(batch, self.spec.d_max, self.spec.rows, self.spec.gen_id, seed, self.dev)
            return q.cpu().pin_memory()
        q = host_queries(self.spec, batch, seed)
        t = [torch.from_numpy(q[m]).pin_memory() for m in self.spec.slots]
        return t if self.multi else t[0]

    def to_device(self, qh):
        return [t.to(self.dev) for t in qh] if self.multi else qh.to(self.dev)

    def step_dev(self, qd):
        return self.fed.search_slots(qd) if self.multi else self.fed.search(qd)

    def step_host(self, qh):
        return self.fed.search_host(qh)

    def per_source_queries(self, qd):
        return [qd[e] for e in self.fed.slot_of_source] if self.multi else [qd] * len(self.spec.names)


def verify_step(w: GpuWorkload, qd, out, n_sample: int, block_rows: int = 1 << 17, extra: int = 32):
    """Exhaustive float64 check of a few queries of a finished step (after the timed region).

    For `n_sample` queries spread over the batch: every rank scores ALL stored rows of every source the
    router selected for the query in float64 (rows rebuilt bit-exactly from the planes by
    rr_index_reconstruct, fp32 products are exact in fp64), keeps its best k + `extra`, the ranks' lists
    are all-gathered and merged, and rank 0 compares with what the timed path returned:
      * an id returned but not in the float64 top-k (or the reverse) is a documented TIE iff its float64
        score is within 1e-5 relative of the float64 k-th score, otherwise a MISMATCH;
      * every returned score must be within 1e-5 relative of the float64 score of its row;
      * returned lists must be descending and the right length."""
    import torch
    import torch.distributed as dist
    fed, spec, dev, world = w.fed, w.spec, w.dev, w.world
    k, kk = spec.k, spec.k + extra
    B = out["scores"].shape[0]
    nv = max(1, min(n_sample, B))
    qs = torch.linspace(0, B - 1, nv, device=dev).round().long().unique()
    nv = int(qs.numel())
    mask = out["mask"][qs].bool()                      # [nv, S] (identical on every rank)
    q_src = w.per_source_queries(qd)
    best_s = torch.full((nv, kk), -float("inf"), dtype=torch.float64, device=dev)
    best_src = torch.full((nv, kk), -1, dtype=torch.int64, device=dev)
    best_id = torch.full((nv, kk), -1, dtype=torch.int64, device=dev)
    rows_checked = 0
    buf = {}
    for s, ix in enumerate(fed.indexes):
        n = ix.ntotal
        routed = mask[:, s]
        if n == 0 or not bool(routed.any().item()):
            continue
        d = ix.d
        if d not in buf:
            buf[d] = torch.empty((block_rows, d), dtype=torch.float32, device=dev)
        qv = q_src[s][qs].double()                     # [nv, d]
        for r0 in range(0, n, block_rows):
            nr = min(block_rows, n - r0)
            blk = ix.reconstruct_tensor(r0, nr, out=buf[d][:nr])
            sc = qv @ blk.double().T                   # [nv, nr] float64, products exact
            sc[~routed] = -float("inf")
            t = min(kk, nr)
            v, j = torch.topk(sc, t, dim=1)
            cat_s = torch.cat([best_s, v], dim=1)
            cat_src = torch.cat([best_src, torch.full_like(j, s)], dim=1)
            cat_id = torch.cat([best_id, j + (r0 + ix.id_offset)], dim=1)
            best_s, o = torch.topk(cat_s, kk, dim=1)
            best_src, best_id = torch.gather(cat_src, 1, o), torch.gather(cat_id, 1, o)
        rows_checked += n
    best_src = torch.where(torch.isinf(best_s), torch.full_like(best_src, -1), best_src)
    if world > 1:
        gs = [torch.empty_like(best_s) for _ in range(world)]
        gsrc = [torch.empty_like(best_src) for _ in range(world)]
        gid = [torch.empty_like(best_id) for _ in range(world)]
        dist.all_gather(gs, best_s)
        dist.all_gather(gsrc, best_src)
        dist.all_gather(gid, best_id)
        cat_s, cat_src, cat_id = torch.cat(gs, 1), torch.cat(gsrc, 1), torch.cat(gid, 1)
        best_s, o = torch.topk(cat_s, kk, dim=1)
        best_src, best_id = torch.gather(cat_src, 1, o), torch.gather(cat_id, 1, o)
        t = torch.tensor([float(rows_checked)], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        rows_checked = int(t.item())
    if w.rank != 0:
        return None
    ref_s, ref_src, ref_id = best_s.cpu().numpy(), best_src.cpu().numpy(), best_id.cpu().numpy()
    got_s = out["scores"][qs].cpu().numpy()
    got_id = out["ids"][qs].cpu().numpy()
    got_src = out["source"][qs].cpu().numpy()
    mism = ties = bad_scores = 0
    max_rel = 0.0
    tol = 1e-5
    for i in range(nv):
        ref = {(int(ref_src[i, j]), int(ref_id[i, j])): float(ref_s[i, j]) for j in range(kk) if ref_src[i, j] >= 0}
        order = [(int(ref_src[i, j]), int(ref_id[i, j])) for j in range(kk) if ref_src[i, j] >= 0]
        want = order[:k]
        got = [(int(got_src[i, j]), int(got_id[i, j])) for j in range(k) if got_id[i, j] >= 0]
        if len(got) != len(want) or len(set(got)) != len(got):
            mism += 1
            continue
        if np.any(np.diff(got_s[i, :len(got)].astype(np.float64)) > 0):
            mism += 1
            continue
        kth = ref[want[-1]] if want else 0.0
        q_bad = False
        for key in set(got) ^ set(want):
            sc = ref.get(key)
            if sc is not None and abs(sc - kth) <= tol * max(abs(kth), 1e-30):
                ties += 1
            else:
                q_bad = True
        for j, key in enumerate(got):
            sc = ref.get(key)
            if sc is None:
                continue
            rel = abs(float(got_s[i, j]) - sc) / max(abs(sc), 1e-30)
            max_rel = max(max_rel, rel)
            if rel > tol:
                bad_scores += 1
                q_bad = True
        mism += 1 if q_bad else 0
    return {"queries": nv, "mismatches": mism, "ties": ties, "score_errors": bad_scores, "max_rel_score_err": max_rel,
            "rows_checked": rows_checked, "sources_checked_per_query": float(mask.float().sum(1).mean().item()),
            "method": "float64 exhaustive top-k over every rank's stored rows (rr_index_reconstruct blocks), merged "
                      "across ranks; tolerance 1e-5 relative (north_star)"}


def gpu_main(args, spec: Spec):
    import torch
    import torch.distributed as dist

    from ragroute_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the GPU arm has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # a collective that does not complete is a bug here, never a slow step: fail in minutes, not in ten
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    if world != args.gpus and rank == 0:
        log(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    lib = _lib.load()

    # does this rank's part fit?  (4 bytes per stored element + generation buffers)
    def need_bytes(sp: Spec) -> int:
        if sp.placement == "whole-source":
            per_rank = [0] * world
            for i, n in enumerate(sp.names):
                per_rank[i % world] += sp.rows[n] * sp.d_pad(n) * BYTES_PER_ELEM_STORED
            return max(per_rank)
        return sum(-(-sp.rows[n] // world) * sp.d_pad(n) for n in sp.names) * BYTES_PER_ELEM_STORED

    need = need_bytes(spec)
    free_b, _ = torch.cuda.mem_get_info(dev)
    if need + (6 << 30) > free_b and args.workload == "auto":
        log(f"[bench] federation needs {need / 1e9:.1f} GB per rank, {free_b / 1e9:.1f} GB free: "
            f"falling back to the largest single-GPU configuration (pubmed)")
        spec = make_spec("pubmed", args.rows_scale)
        need = need_bytes(spec)
    if need + (6 << 30) > free_b:
        raise SystemExit(f"bench.py: workload '{spec.workload}' needs {need / 1e9:.1f} GB per rank at world={world}, "
                         f"only {free_b / 1e9:.1f} GB free; pick --workload pubmed or more GPUs")
    K = spec.k

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def build(corpus: str) -> GpuWorkload:
        t_build = time.time()
        w = GpuWorkload(spec, rank, world, local_rank, args.routing, corpus)
        if rank == 0:
            log(f"[bench] built {spec.workload} ({corpus}) rank-0 rows {w.local_rows} in {time.time() - t_build:.1f}s "
                f"(free {torch.cuda.mem_get_info(dev)[0] / 1e9:.1f} GB)")
        for ix in w.fed.indexes:
            ix.set_param("profile", 1)
            if args.scan_path:
                ix.set_param("scan_path", args.scan_path)
            for kv in args.index_param:
                name, val = kv.split("=")
                ix.set_param(name, float(val))
        return w

    def measure(w: GpuWorkload, batch: int, steps: int, warmup: int, host: bool, sampler=None, n_verify: int = 0,
                in_flight: bool = False):
        """Time `steps` steps; returns dict(ms_total, per_step_ms, launches, scan stats, verify).
        in_flight: the step is a wave of `batch` concurrent single-query requests through the micro-batching
        collector (run_benchmark.py:90-114 --parallel), each request a host embedding per encoder."""
        fed = w.fed
        n_distinct = min(warmup + steps, 24)   # a cycle of distinct batches (each >> what any cache could replay)
        q_host = [w.pinned_queries(batch, 42 + i) for i in range(n_distinct)]
        q_dev = None if host else [w.to_device(q) for q in q_host]
        torch.cuda.synchronize(dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        wave = make_wave(w, batch) if in_flight else None

        def one(i):
            j = i % n_distinct
            if in_flight:
                return wave(q_host[j])
            return w.step_host(q_host[j]) if host else w.step_dev(q_dev[j])

        sink = None
        for i in range(warmup):
            sink = one(i)
        for ix in fed.indexes:
            ix.stats(reset=True)
        barrier()
        l0 = lib.rr_kernel_launches()
        if sampler:
            sampler.start()
        ev[0].record()
        for i in range(steps):
            sink = one(warmup + i)
            ev[i + 1].record()
        torch.cuda.synchronize(dev)
        if sampler:
            sampler.stop()
        barrier()
        l1 = lib.rr_kernel_launches()
        ms_total = max_over_ranks(ev[0].elapsed_time(ev[steps]))
        per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
        st = [ix.stats(reset=True) for ix in fed.indexes]
        res = {"ms_total": ms_total, "per_step": per_step, "launches": int(sum_over_ranks(l1 - l0)),
               "scan_ms": sum(s["scan_ms_total"] for s in st), "scan_bytes": sum(s["scan_bytes_total"] for s in st),
               "scan_flops": sum(s["scan_flops_total"] for s in st), "scan_launches": sum(s["scan_launches"] for s in st),
               "exact_rescans": int(sum_over_ranks(sum(s["exact_rescans"] for s in st))),
               "exact_scans": int(sum_over_ranks(sum(s["exact_scans"] for s in st))), "steps": steps, "verify": None}
        if n_verify and not host:
            j = (warmup + steps - 1) % n_distinct
            out = w.step_dev(q_dev[j])           # the last timed batch again, outside the timed region
            torch.cuda.synchronize(dev)
            res["verify"] = verify_step(w, q_dev[j], out, n_verify)
            for ix in fed.indexes:
                ix.stats(reset=True)
        del sink
        return res

    def make_wave(w: GpuWorkload, batch: int):
        """One wave of `batch` concurrent single-query requests (the client of run_benchmark.py:90-114) against a
        MicroBatcher whose batch function is the host entry point of the federation."""
        import asyncio
        from ragroute_b200.collector import MicroBatcher
        loop = asyncio.new_event_loop()

        pinned = {}  # per encoder slot: one pinned staging buffer [batch, d], reused by every wave

        def batch_fn(items):
            # items: per request the tuple of per-encoder embeddings [1, d] (host) -> one pinned batch per encoder
            cols = list(zip(*items))
            qh = []
            for e, c in enumerate(cols):
                buf = pinned.get(e)
                if buf is None:
                    buf = pinned[e] = torch.empty((batch, c[0].shape[1]), dtype=torch.float32).pin_memory()
                qh.append(torch.cat(c, dim=0, out=buf[:len(c)]))
            s, i, src = w.step_host(qh if w.multi else qh[0])
            return [(s[b], i[b], src[b]) for b in range(len(items))]

        mb = MicroBatcher(batch_fn, max_batch=batch, max_wait_s=200e-6)

        async def run(qh):
            slots = qh if w.multi else [qh]
            reqs = [tuple(t[b:b + 1] for t in slots) for b in range(batch)]
            return await asyncio.gather(*[mb.submit(r) for r in reqs])

        def wave(qh):
            return loop.run_until_complete(run(qh))

        wave.batcher = mb
        return wave

    peaks = {}
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # the scan runs inside a long step under the power cap: the sustained cuBLAS figure is the denominator
    tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    tc_src = ("MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16, seconds-long loop)" if "bf16_tflops_sustained" in peaks
              else "fallback 1400 TFLOP/s sustained (B200_PROFILING.md)")
    try:
        with open(os.path.join(REPO, "profiles", "scan_traffic.json")) as f:
            traffic_db = json.load(f)
    except Exception:  # noqa: BLE001
        traffic_db = {}

    def roofline_of(run: dict, batch: int) -> dict:
        """Roofline entry of the dominant kernel of a run (rank 0's launches, CUDA events on the
        launching stream).  batch <= 256 (at most ~128 routed queries per source): HBM-bound, the
        intensity over the bf16 plane is below the ridge; larger batches: bound by the bf16 tensor pipe."""
        n = run["scan_launches"]
        if not n:
            return {"bound": None, "achieved": None, "peak": None, "frac": None, "traffic": None}
        small = batch <= 256  # one 128-query tile per CTA or the swap-AB scan: HBM-bound
        # dispatch of rr_index.cu: swap-AB scan up to 24 queries, queries-in-TMEM scan above
        key = "rr_tc_scan_kernel" if (batch <= 24 and args.scan_path == 0) or args.scan_path == 1 else "rr_gemm_scan_kernel"
        tj = traffic_db.get(key + "_small_batch" if (key == "rr_gemm_scan_kernel" and small) else key) or traffic_db.get(key, {})
        traffic = tj.get("dram_over_algorithmic", 0) * run["scan_bytes"] / n if tj else None
        common = {"kernel": key, "launches_timed": n, "avg_launch_ms": run["scan_ms"] / n,
                  "kernel_share_of_step": run["scan_ms"] / max(run["ms_total"], 1e-9),
                  "algorithmic_bytes_per_launch": run["scan_bytes"] / n,
                  "algorithmic_flops_per_launch": run["scan_flops"] / n,
                  "traffic": traffic,
                  "traffic_source": ("ncu dram read+write / algorithmic bytes = %.5f (%s), scaled to this launch size"
                                     % (tj["dram_over_algorithmic"], tj["source"])) if tj else None}
        gbs = run["scan_bytes"] / max(run["scan_ms"], 1e-9) * 1e-6
        tfs = run["scan_flops"] / max(run["scan_ms"], 1e-9) * 1e-9
        if small:
            return {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                    "peak_source": hbm_src, "tflops": tfs, **common}
        return {"bound": "tensor", "achieved": tfs, "peak": tc_peak, "unit": "TFLOP/s", "frac": tfs / tc_peak,
                "peak_source": tc_src, "hbm_gbs": gbs, **common}

    def ideal_ms(batch: int, run: dict) -> float:
        """Roofline time of the whole step: the larger of the HBM time of the bytes the scans covered and the
        tensor time of the flops they scored, per step, on this rank."""
        st = max(run["steps"], 1)
        return max(run["scan_bytes"] / st / (hbm_peak * 1e6), run["scan_flops"] / st / (tc_peak * 1e9))

    sampler = ClockSampler(local_rank) if rank == 0 else None
    B = args.batch
    in_flight = spec.workload == "feb4rag16"
    w = build(args.corpus)
    dev_run = measure(w, B, args.steps, args.warmup, host=False, sampler=sampler, n_verify=args.verify_queries)
    e2e_run = measure(w, B, args.steps, args.warmup, host=True, in_flight=in_flight)
    qps = B * args.steps / (dev_run["ms_total"] * 1e-3)
    qps_e2e = B * args.steps / (e2e_run["ms_total"] * 1e-3)

    sweep = []
    roofline_small = None
    if args.sweep:
        for b in [int(x) for x in args.sweep.split(",") if x]:
            if b == B:
                r, st_ = dev_run, args.steps
            else:
                per_step = dev_run["ms_total"] / args.steps * max(b, 64) / max(B, 64)
                st_ = int(args.sweep_ms_budget / max(0.05, per_step))
                # latency percentiles of the small batches come from >= 200 timed steps
                st_ = max(200, min(1000, st_)) if b <= 64 else max(10, min(60, st_))
                r = measure(w, b, st_, 5, host=False, n_verify=min(4, args.verify_queries))
            rf = roofline_of(r, b)
            ms_step = r["ms_total"] / st_
            sweep.append({"batch": b, "steps": st_, "qps": b * st_ / (r["ms_total"] * 1e-3),
                          "p50_ms": pct(r["per_step"], 50), "p99_ms": pct(r["per_step"], 99),
                          "bound": rf["bound"], "achieved": rf["achieved"], "unit": rf.get("unit"), "frac": rf["frac"],
                          "whole_step_frac": ideal_ms(b, r) / max(ms_step, 1e-9),
                          "exact_rescans": r["exact_rescans"], "exact_scans": r["exact_scans"], "verify": r["verify"]})
            if b == 64:
                roofline_small = rf

    # (every rank runs this step: a search is a collective when the federation is sharded)
    qd = w.to_device(w.pinned_queries(B, 42))
    mask_density = float(w.step_dev(qd)["mask"].float().mean().item())
    del qd
    local_rows = w.local_rows

    # near-duplicate corpus leg: the same workload with documents stored as runs of
    # near-duplicate chunks, where first-pass certificates fail and the repair path does the work
    clustered = None
    if spec.dataset == "medrag" and args.corpus == "iid" and not args.no_clustered_leg:
        w.close()
        w = build("clustered")
        cs = max(3, min(args.steps, 10))
        r = measure(w, B, cs, 3, host=False, n_verify=args.verify_queries)
        clustered = {"corpus": f"runs of 128 near-duplicate rows (spread 0.03/sqrt(d)), queries next to run centres",
                     "batch": B, "steps": cs, "qps": B * cs / (r["ms_total"] * 1e-3), "ms_per_step": r["ms_total"] / cs,
                     "exact_rescans_per_step": r["exact_rescans"] / cs, "exact_scans_per_step": r["exact_scans"] / cs,
                     "verify": r["verify"]}
        if args.sweep and 64 in [int(x) for x in args.sweep.split(",") if x]:
            r64 = measure(w, 64, 30, 3, host=False, n_verify=min(4, args.verify_queries))
            clustered["batch64"] = {"qps": 64 * 30 / (r64["ms_total"] * 1e-3), "ms_per_step": r64["ms_total"] / 30,
                                    "exact_rescans_per_step": r64["exact_rescans"] / 30,
                                    "exact_scans_per_step": r64["exact_scans"] / 30, "verify": r64["verify"]}
    w.close()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_legs(spec, B, args.cpu_sample_rows, steps=3, warmup=1, routing=args.routing)

    if rank == 0:
        ms_step = dev_run["ms_total"] / args.steps
        line = {
            "metric": METRIC_FEB if spec.workload == "feb4rag16" else METRIC, "value": qps, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "dtype_detail": "bf16 tensor-core scan selects candidates; returned scores are fp32-exact (fp32 rows, fp64 accumulate)",
            "data": "synthetic",
            "config": {"workload": spec.workload, "corpus": args.corpus, "rows": spec.rows, "rows_rank0": local_rows,
                       "dim": spec.d_max if spec.dataset == "medrag" else spec.dims, "k": K, "batch": B,
                       "routing": args.routing, "mask_density": mask_density,
                       "parallelism": (f"{spec.placement} x{world}" if world > 1 else "single GPU"),
                       "in_flight": B if in_flight else None,
                       "l2": "inputs larger than L2 (bf16 scan plane per rank >> 126 MB); no flush needed"
                             if need / 2 > (512 << 20) else "corpus comparable to L2: re-scans may hit L2"},
            "p50_ms": pct(dev_run["per_step"], 50), "p99_ms": pct(dev_run["per_step"], 99),
            "e2e": {"value": qps_e2e, "unit": UNIT, "h2d_bytes_per_step": B * spec.query_bytes(),
                    "d2h_bytes_per_step": B * K * (4 + 8 + 4), "ms_per_step": e2e_run["ms_total"] / args.steps,
                    "p50_ms": pct(e2e_run["per_step"], 50), "p99_ms": pct(e2e_run["per_step"], 99),
                    "path": ("MicroBatcher wave of %d concurrent single-query requests -> Federation.search_host" % B)
                            if in_flight else "Federation.search_host (pinned host queries in, top-k out, sync per step)"},
            "gpu_launches": dev_run["launches"],
            "roofline": roofline_of(dev_run, B),
            "whole_step_frac": ideal_ms(B, dev_run) / max(ms_step, 1e-9),
            "roofline_small_batch": roofline_small,
            "exact_rescans": dev_run["exact_rescans"], "exact_scans": dev_run["exact_scans"],
            "verify": dev_run["verify"],
            "clocks": sampler.summary() if sampler else None,
            "cpu_baseline": cpu_baseline,
        }
        if clustered:
            line["clustered"] = clustered
        if sweep:
            line["sweep"] = sweep
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    # Rank 0 must print exactly ONE line on stdout: libraries (NCCL prints its version banner there)
    # are sent to stderr; the JSON line goes to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + list(WORKLOADS))
    ap.add_argument("--corpus", default="iid", choices=["iid", "clustered"],
                    help="clustered: documents stored as runs of near-duplicate chunks (MedRAG workloads)")
    ap.add_argument("--batch", type=int, default=0,
                    help="queries per step (default 2048; feb4rag16: 64 in flight); <= 64 is the HBM-bound regime, "
                         ">= 1024 the tensor-bound one")
    ap.add_argument("--routing", default=None, choices=["ragroute", "all"],
                    help="default: ragroute; 'all' (no router in the step, every query searches every source) for the "
                         "textbooks workload = BASELINE configs[0], one source alone")
    ap.add_argument("--rows-scale", type=float, default=1.0, help="scale every source's row count (debug)")
    ap.add_argument("--sweep", default=None, help="comma separated extra batch sizes ('' = none; default per workload)")
    ap.add_argument("--sweep-ms-budget", type=float, default=3000.0)
    ap.add_argument("--verify-queries", type=int, default=8, help="queries per measured run re-done exhaustively in float64")
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clustered-leg", action="store_true")
    ap.add_argument("--index-param", action="append", default=[], metavar="NAME=VALUE",
                    help="rr_index_set_param on every source (experiments), e.g. gemm_tmem_all=0")
    ap.add_argument("--scan-path", type=int, default=0,
                    help="debug: force a scan kernel (1 = 64-query passes, 2 = large-batch, 3 = large-batch CTA pairs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        log("[bench] warmup raised to 3 (timing rule)")
        args.warmup = 3
    workload = "federation" if args.workload == "auto" else args.workload
    if args.batch <= 0:
        args.batch = 64 if workload == "feb4rag16" else 2048
    if args.sweep is None:
        args.sweep = DEFAULT_SWEEP[workload]
    if args.routing is None:
        args.routing = "all" if workload == "textbooks" else "ragroute"
    if workload == "feb4rag16" and args.rows_scale != 1.0:
        raise SystemExit("bench.py: feb4rag16 has a fixed layout (--rows-scale is for the MedRAG workloads)")
    spec = make_spec(workload, args.rows_scale)
    if args.corpus == "clustered" and spec.dataset != "medrag":
        raise SystemExit("bench.py: --corpus clustered applies to the MedRAG workloads")
    if args.impl == "reference":
        return reference_main(args, spec)
    return gpu_main(args, spec)


if __name__ == "__main__":
    sys.exit(main())

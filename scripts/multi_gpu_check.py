"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL):
the row-sharded federation (per-rank shards -> one all-gather of the packed lists -> in-place merge)
must return bit for bit what the unsharded federation returns on one GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/multi_gpu_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ragroute_b200 import synthetic as syn  # noqa: E402
from ragroute_b200.workloads import build_medrag_federation  # noqa: E402


def tie_case(rank, world, local):
    """Small-integer corpora: scores are exact in bf16 and most of them are equal, documents are duplicated
    across the two sources and across the shards -- the merge's tie order (score, source, id) must make the
    sharded result equal the unsharded one bit for bit."""
    import numpy as np
    from ragroute_b200.faiss_compat import IndexFlatIP
    from ragroute_b200.federation import Federation, shard_rows
    rng = np.random.default_rng(7)
    d, k = 64, 32
    corpora = {"a": rng.integers(-2, 3, size=(5003, d)).astype(np.float32)}
    corpora["b"] = np.concatenate([corpora["a"][1000:3000], rng.integers(-2, 3, size=(1500, d)).astype(np.float32)])
    def build(w, r):
        src = {}
        for n, x in corpora.items():
            lo, hi = shard_rows(len(x), w, r)
            ix = IndexFlatIP(d, device=local)
            if hi > lo:
                ix.add(x[lo:hi])
            ix.set_id_offset(lo)
            src[n] = ix
        return Federation(src, k, routing="all", distributed=w > 1)
    sharded, whole = build(world, rank), build(1, 0)
    bad = 0
    for B in (3, 70, 400):
        q = torch.from_numpy(rng.integers(-2, 3, size=(B, d)).astype(np.float32)).cuda()
        a, b = sharded.search(q), whole.search(q)
        torch.cuda.synchronize()
        same = all(torch.equal(a[key], b[key]) for key in ("scores", "ids", "source"))
        n_tied = int((a["scores"][:, :-1] == a["scores"][:, 1:]).sum().item())
        if rank == 0:
            print(f"world={world} B={B}: tie-heavy federation sharded == unsharded: {same} ({n_tied} adjacent ties)", flush=True)
        bad += (not same) or n_tied == 0
    return bad


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    rows = {"pubmed": 239_001, "statpearls": 30_100, "textbooks": 12_500, "wikipedia": 299_003}
    sharded, local_rows = build_medrag_federation(rows, rank=rank, world=world, device=local)
    whole, _ = build_medrag_federation(rows, rank=0, world=1, device=local)
    whole.distributed = False
    sids = [syn.MEDRAG_SOURCE_TO_ID[n] for n in rows]
    bad = 0
    for B in (1, 7, 64, 300, 2100):
        q = torch.from_numpy(syn.queries(B, 768, sids, seed=B)).cuda()
        a = sharded.search(q)
        b = whole.search(q)
        torch.cuda.synchronize()
        same = all(torch.equal(a[key], b[key]) for key in ("scores", "ids", "source", "mask"))
        # the winner's rank must be the shard that owns the row
        ok_rank = True
        if same:
            ids, src, rk = a["ids"].cpu(), a["source"].cpu(), a["rank"].cpu()
            names = list(rows)
            for bb in range(min(B, 16)):
                for j in range(ids.shape[1]):
                    if ids[bb, j] >= 0:
                        per = -(-rows[names[src[bb, j]]] // world)
                        ok_rank &= int(ids[bb, j]) // per == int(rk[bb, j])
        if rank == 0:
            print(f"world={world} B={B}: sharded == unsharded: {same}; ranks consistent: {ok_rank}", flush=True)
        bad += (not same) or (not ok_rank)
    bad += tie_case(rank, world, local)
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    dist.barrier()
    dist.destroy_process_group()
    if int(t.item()):
        raise SystemExit(1)
    if rank == 0:
        print("multi-GPU parity ok", flush=True)


if __name__ == "__main__":
    main()

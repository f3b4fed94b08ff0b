"""Train a router for the synthetic MedRAG-shaped federation with the REFERENCE's recipe, write its artefacts in
the reference's file formats, and load them back through ragroute_b200.artifacts.

Recipe = /root/reference/scripts/train/train_medrag_router.py: CorpusRoutingNN (:66-84), features
[query | centroid | one-hot] as the SERVER builds them (router.py:253-262; the trainer's :145 omits the one-hot,
SURVEY 8a notes), StandardScaler fitted on the training rows (:349-351), BCEWithLogitsLoss, Adam(lr 1e-3,
weight_decay 1e-5), CyclicLR(1e-3 .. 5e-3, step_size_up 10, triangular2) then StepLR(50, 0.05), gradient clipping
at 1.0, batch 128, keep the state with the best validation accuracy (:385-466); labels = which corpora contribute
to the merged global top-32 of a query (medrag_relevance/*_relevant_top_32.json), here taken from this library's
own search with routing 'all'.

Artefacts (one directory): best_model.pth (state_dict), preprocessed_data.pkl (the 5-tuple whose 4th item is the
scaler, :240-245 / router.py:123-125), <corpus>_stats.json with "centroid" computed on the GPU by
rr_index_centroid (router.py:136-151), threshold.json.  Then: artifacts.load_router on those files, logits of the
fused kernel against the torch module that was trained (<= 1e-5), routing quality on held-out queries.

    python scripts/train_synthetic_router.py --out gpurun_out/trained_router [--rows-scale 1.0] [--epochs 40]
"""
import argparse
import json
import os
import pickle
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ragroute_b200 import artifacts, synthetic as syn  # noqa: E402
from ragroute_b200.workloads import build_medrag_federation  # noqa: E402

ROWS = {"pubmed": 23_900_000, "statpearls": 301_000, "textbooks": 125_000, "wikipedia": 29_900_000}
ENCODER = "ncbi/MedCPT-Query-Encoder"


class CorpusRoutingNN(nn.Module):  # train_medrag_router.py:66-84 == router.py:37-55
    def __init__(self, input_dim):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, 256)
        self.ln1 = nn.LayerNorm(256)
        self.dropout1 = nn.Dropout(0.4)
        self.fc2 = nn.Linear(256, 128)
        self.ln2 = nn.LayerNorm(128)
        self.dropout2 = nn.Dropout(0.4)
        self.fc3 = nn.Linear(128, 1)

    def forward(self, x):
        x = self.dropout1(F.relu(self.ln1(self.fc1(x))))
        x = self.dropout2(F.relu(self.ln2(self.fc2(x))))
        return self.fc3(x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/trained_router")
    ap.add_argument("--rows-scale", type=float, default=1.0)
    ap.add_argument("--queries", type=int, default=6144)
    ap.add_argument("--epochs", type=int, default=40)
    args = ap.parse_args()
    from sklearn.preprocessing import StandardScaler
    torch.manual_seed(42)
    np.random.seed(42)
    os.makedirs(args.out, exist_ok=True)
    dev = torch.device("cuda", 0)
    rows = {n: max(32, int(r * args.rows_scale)) for n, r in ROWS.items()}
    names = list(rows)
    d, k = syn.MEDRAG_DIM, syn.MEDRAG_K
    fed, _ = build_medrag_federation(rows, routing="all", device=0)
    sids = [syn.MEDRAG_SOURCE_TO_ID[n] for n in names]

    # ---- centroids on the GPU -> <corpus>_stats.json
    stats_files = {}
    centroids = {}
    for n, ix in zip(names, fed.indexes):
        c = ix.centroid()
        centroids[n] = c
        stats_files[n] = os.path.join(args.out, f"{n}_stats.json")
        with open(stats_files[n], "w") as f:
            json.dump({"centroid": [float(v) for v in c], "num_documents": int(ix.ntotal)}, f)

    # ---- labels: the corpora that contribute to the merged global top-32 (routing = all)
    q = syn.queries(args.queries, d, sids, seed=2024)
    labels = np.zeros((args.queries, len(names)), dtype=np.float32)
    t0 = time.time()
    for b0 in range(0, args.queries, 2048):
        out = fed.search(torch.from_numpy(q[b0:b0 + 2048]).to(dev))
        src = out["source"].cpu().numpy()
        for s in range(len(names)):
            labels[b0:b0 + 2048, s] = (src == s).any(axis=1)
    print(f"labels from {args.queries} routing-all searches in {time.time() - t0:.1f}s; relevant fraction per corpus: "
          f"{dict(zip(names, labels.mean(axis=0).round(3)))}", flush=True)

    # ---- features exactly as the server builds them (router.py:253-262)
    onehot = np.eye(len(syn.MEDRAG_SOURCE_TO_ID))
    feats = np.empty((args.queries, len(names), 2 * d + onehot.shape[0]))   # float64 rows (np.eye), as the reference builds them
    for s, n in enumerate(names):
        feats[:, s, :d] = q
        feats[:, s, d:2 * d] = centroids[n]
        feats[:, s, 2 * d:] = onehot[syn.MEDRAG_SOURCE_TO_ID[n]]
    n_train, n_val = int(args.queries * 0.6), int(args.queries * 0.1)
    tr, va, te = slice(0, n_train), slice(n_train, n_train + n_val), slice(n_train + n_val, args.queries)
    scaler = StandardScaler()
    x_tr = scaler.fit_transform(feats[tr].reshape(-1, feats.shape[-1]))
    x_va = scaler.transform(feats[va].reshape(-1, feats.shape[-1]))
    y_tr, y_va = labels[tr].reshape(-1), labels[va].reshape(-1)

    # ---- the reference's training loop (:385-466)
    model = CorpusRoutingNN(feats.shape[-1]).to(dev)
    criterion = nn.BCEWithLogitsLoss()
    opt = torch.optim.Adam(model.parameters(), lr=0.001, weight_decay=1e-5)
    cyc = torch.optim.lr_scheduler.CyclicLR(opt, base_lr=1e-3, max_lr=5e-3, step_size_up=10, mode="triangular2",
                                            cycle_momentum=False)
    fixed = torch.optim.lr_scheduler.StepLR(opt, step_size=50, gamma=0.05)
    xt, yt = torch.tensor(x_tr, dtype=torch.float32, device=dev), torch.tensor(y_tr, device=dev)
    xv, yv = torch.tensor(x_va, dtype=torch.float32, device=dev), torch.tensor(y_va, device=dev)
    best_acc, best_state = 0.0, None
    switch = int(args.epochs * 115 / 150)
    for epoch in range(args.epochs):
        model.train()
        perm = torch.randperm(len(xt), device=dev)
        for i in range(0, len(xt), 128):
            idx = perm[i:i + 128]
            opt.zero_grad()
            loss = criterion(model(xt[idx]).squeeze(-1), yt[idx])
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
            opt.step()
            (cyc if epoch < switch else fixed).step()
        model.eval()
        with torch.no_grad():
            acc = float(((model(xv).squeeze(-1) > 0.5).float() == yv).float().mean())   # the trainer thresholds the LOGIT at 0.5 (:443)
        if acc > best_acc:
            best_acc, best_state = acc, {k_: v.detach().cpu().clone() for k_, v in model.state_dict().items()}
    model.load_state_dict(best_state)
    model.eval()
    torch.save(best_state, os.path.join(args.out, "best_model.pth"))
    with open(os.path.join(args.out, "preprocessed_data.pkl"), "wb") as f:   # (train, val, test, scaler, val_qs): router.py:123-125
        pickle.dump(([], [], {}, scaler, []), f)

    # threshold on the validation set: best F1 over a grid (the reference's find_optimal_threshold picks from the ROC curve)
    with torch.no_grad():
        pv = torch.sigmoid(model(xv).squeeze(-1)).cpu().numpy()
    grid = np.linspace(0.05, 0.95, 91)
    f1 = []
    for th in grid:
        pr = pv > th
        tp, fp, fn = (pr & (y_va > 0)).sum(), (pr & (y_va == 0)).sum(), (~pr & (y_va > 0)).sum()
        f1.append(2 * tp / max(2 * tp + fp + fn, 1))
    thr = float(grid[int(np.argmax(f1))])
    with open(os.path.join(args.out, "threshold.json"), "w") as f:
        json.dump({"threshold": thr, "val_accuracy_at_logit_0.5": best_acc}, f)

    # ---- load the artefacts back the way Router.load_router does (router.py:106-151) and compare with the torch module
    router, _ = artifacts.load_router("medrag", names, {n: ENCODER for n in names}, os.path.join(args.out, "best_model.pth"),
                                      stats_files, os.path.join(args.out, "preprocessed_data.pkl"), device=0)
    q_te = torch.from_numpy(q[te]).to(dev)
    mask, logits = router.select(q_te, thr, return_logits=True)
    with torch.no_grad():
        x_te = torch.tensor(scaler.transform(feats[te].reshape(-1, feats.shape[-1])), dtype=torch.float32, device=dev)
        ref = model(x_te).reshape(-1, len(names))
    max_diff = float((ref - logits).abs().max())
    ref_mask = torch.sigmoid(ref) > thr
    decided = (ref - float(np.log(thr / (1 - thr)))).abs() > 1e-5
    same = bool((ref_mask == mask.bool())[decided].all())
    y_te = labels[te] > 0
    m = mask.bool().cpu().numpy()
    report = {"rows": rows, "queries": args.queries, "epochs": args.epochs, "threshold": thr, "val_accuracy": best_acc,
              "fused_vs_torch_max_logit_diff": max_diff, "decisions_identical_outside_1e-5_band": same,
              "test": {"queries": int(m.shape[0]), "relevant_fraction": float(y_te.mean()), "selected_fraction": float(m.mean()),
                       "recall_of_relevant_pairs": float((m & y_te).sum() / max(y_te.sum(), 1)),
                       "queries_with_every_relevant_source_selected": float((~y_te | m).all(axis=1).mean()),
                       "searches_saved_vs_all": float(1 - m.mean())}}
    with open(os.path.join(args.out, "report.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report))
    assert max_diff <= 1e-4 and same, "fused router disagrees with the trained torch module"


if __name__ == "__main__":
    main()

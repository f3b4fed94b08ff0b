"""Scan rate on a 1024-dimensional source (FeB4RAG's e5-large / UAE-Large encoders, config.py:44-58):
queries-in-TMEM scan (default dispatch) against the swap-AB scan in 64-query passes."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ragroute_b200.faiss_compat import IndexFlatIP

d, n, k = 1024, 3_000_000, 10
dev = torch.device("cuda")
ix = IndexFlatIP(d, reserve=n)
g = torch.Generator(device=dev); g.manual_seed(1)
mu = torch.randn(d, device=dev, generator=g) * (0.5 / d ** 0.5)
for r0 in range(0, n, 1 << 19):
    m = min(1 << 19, n - r0)
    ix.add(torch.randn((m, d), device=dev, generator=g) / d ** 0.5 + mu)
torch.cuda.synchronize()
bytes_plane = n * d * 2
for B in (8, 32, 64, 128, 256, 1024, 2048):
    q = (torch.randn((B, d), device=dev, generator=g) / d ** 0.5 + mu).contiguous()
    res = {}
    for path in (0, 1):
        ix.set_param("scan_path", path)
        for _ in range(2):
            D, I = ix.search_tensor(q, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if B <= 256 or path == 0 else 2
        e0.record()
        for _ in range(reps):
            D, I = ix.search_tensor(q, k)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[path] = (ms, I.clone())
    same = torch.equal(res[0][1], res[1][1])
    ms0, ms1 = res[0][0], res[1][0]
    print(f"d=1024 n={n} B={B}: default {ms0:.3f} ms ({bytes_plane / ms0 * 1e-9:.2f} TB/s eq., "
          f"{2 * B * n * d / ms0 * 1e-9:.0f} TFLOP/s) | swap-AB passes {ms1:.3f} ms | same ids {same} | "
          f"rescans {ix.stats()['exact_rescans']}")

"""How often does the certificate fail on a corpus with long runs of near-duplicate rows (documents split into
similar chunks, stored contiguously)?  Prints exact_rescans / queries and the search time per batch."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from ragroute_b200.faiss_compat import IndexFlatIP

d, docs, per = 768, 20000, int(sys.argv[1]) if len(sys.argv) > 1 else 100
spread = float(sys.argv[2]) if len(sys.argv) > 2 else 0.03
g = torch.Generator(device="cuda").manual_seed(1)
centers = torch.randn(docs, d, device="cuda", generator=g) / d ** 0.5 + 0.02
ix = IndexFlatIP(d, reserve=docs * per)
for c0 in range(0, docs, 2000):
    c = centers[c0:c0 + 2000]
    x = c.repeat_interleave(per, dim=0) + spread * torch.randn(2000 * per, d, device="cuda", generator=g) / d ** 0.5
    ix.add(x.contiguous())
torch.cuda.synchronize()
print("rows", ix.ntotal, "chunks/doc", per, "spread", spread)
for B in (1, 64, 1024):
    pick = torch.randint(0, docs, (B,), device="cuda", generator=g)
    q = (centers[pick] + 0.05 * torch.randn(B, d, device="cuda", generator=g) / d ** 0.5).contiguous()
    ix.stats(reset=True)
    r0 = ix.stats()["exact_rescans"]
    for _ in range(2):
        D, I = ix.search_tensor(q, 32)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    D, I = ix.search_tensor(q, 32)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    r1 = ix.stats()["exact_rescans"]
    same_doc = ((I // per) == pick[:, None]).float().mean().item()
    print(f"B={B}: {dt*1e3:.2f} ms, exact rescans per search {(r1 - r0) / 3:.1f} of {B}, top-32 from the query's own document: {same_doc:.2f}")

"""Timeline of the queries-in-TMEM scan's block 0 (RR_TRACE=1): where a small-corpus launch spends its time.
usage: RR_TRACE=1 python scripts/gpu_trace.py [rows] [nq]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ragroute_b200 import _lib, synthetic as syn
from ragroute_b200.faiss_compat import IndexFlatIP

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 64
lib = _lib.load()
lib.rr_debug_read_trace.restype = ctypes.c_int
lib.rr_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
ix = IndexFlatIP(768)
ix.add(syn.corpus_rows(2, rows, 768))
ix.set_param("scan_path", 2 if nq <= 128 else 3)
q = torch.from_numpy(syn.queries(nq, 768, [2], seed=1)).cuda()
buf = (ctypes.c_ulonglong * 8192)()
TAGS = {0: "start", 1: "prod: tile's first stage free", 2: "mma: query tile in place", 3: "mma: acc free", 4: "mma: stage landed",
        5: "epi: acc complete", 6: "epi: tile in regs", 7: "epi: all tiles done", 8: "epi: query tile stored", 9: "end"}
for it in range(4):
    ix.search_tensor(q, 32)
    torch.cuda.synchronize()
    n = lib.rr_debug_read_trace(buf, 4096)
ev = sorted(((buf[2 * i + 1], buf[2 * i] >> 32, buf[2 * i] & 0xFFFFFFFF) for i in range(n)))
t0 = ev[0][0]
print(f"rows={rows} nq={nq}: {n} events of the last launch")
for t, tag, tile in ev:
    print(f"{(t - t0) / 1000:9.2f} us  {TAGS.get(tag, tag):32s} tile {tile}")

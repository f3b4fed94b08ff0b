"""Rebuild the inputs of the golden router cases from the seeds stored in
tests/golden/router_*.npz (the generating script is tests/golden/gen_golden.py;
it ran the reference's own code for the recorded outputs)."""
import os
import subprocess

import numpy as np

from ragroute_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class RouterCase:
    def __init__(self, dataset: str):
        g = np.load(os.path.join(GOLDEN, f"router_{dataset}.npz"))
        self.dataset = dataset
        self.seed = int(g["seed"])
        self.in_dim = int(g["in_dim"])
        self.d_max = int(g["d_max"])
        self.n_queries = int(g["n_queries"])
        self.data_sources = [str(s) for s in g["data_sources"]]
        self.models = [str(s) for s in g["models"]]
        self.model_dims = {m: int(d) for m, d in zip(self.models, g["model_dims"])}
        self.model_of_source = {c: str(m) for c, m in zip(self.data_sources, g["model_of_source"])}
        self.source_dims = [int(d) for d in g["source_dims"]]
        self.onehot_ids = {c: int(i) for c, i in zip(self.data_sources, g["onehot_ids"])}
        self.n_onehot = self.in_dim - 2 * self.d_max
        self.threshold = float(g["threshold"])
        self.ref_logits = g["logits"]
        self.ref_masks = g["masks"]
        self.sd = syn.router_state_dict(self.in_dim, self.seed)
        self.sd["fc3.bias"] = (self.sd["fc3.bias"] + np.float32(g["fc3_bias_shift"])).astype(np.float32)
        self.scaler = None
        if dataset in ("medrag", "wikipedia"):
            self.scaler = syn.router_scaler(self.in_dim, self.seed)
        self.centroids = {}
        for i, c in enumerate(self.data_sources):
            cen = syn.hash_uniform(self.seed * 1000 + 500 + i, (self.source_dims[i],), -0.3, 0.3)
            self.centroids[c] = np.pad(cen, (0, self.d_max - self.source_dims[i]))

    def query(self, qi: int):
        return {m: syn.hash_uniform(self.seed * 100000 + 7000 + qi * 16 + mi, (self.model_dims[m],), -0.4, 0.4)
                for mi, m in enumerate(self.models)}


def build_abi_check(tmp_path):
    """gcc tests/c/abi_check.c against include/ragroute_b200.h and the in-tree library -> executable path."""
    from ragroute_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as entry
        entry.build_cuda()
    exe = str(tmp_path / "abi_check")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"),
                    os.path.join(REPO, "tests", "c", "abi_check.c"), "-o", exe, "-L", libdir, "-lragroute_b200",
                    "-Wl,-rpath," + libdir], check=True)
    return exe

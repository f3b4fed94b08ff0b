"""Boundary B2: the router's relevance network and source selection.

``FusedCorpusRouter`` stands where ``CorpusRoutingNN`` stands in
/root/reference/ragroute/router.py:37-55 (constructed :117, ``load_state_dict``
:118, ``.eval()`` :119, ``.fc1.in_features`` :174, ``self.router(x)`` :274) and adds
the fused batch form of ``select_relevant_sources_ragroute`` (:241-283): feature
build + scaler + MLP + sigmoid + threshold for every (query, source) pair in one
kernel launch.  All arithmetic runs in libragroute_b200.so; there is no CPU path.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib

MEDRAG_THRESHOLD = 0.4924  # router.py:278
DEFAULT_THRESHOLD = 0.5    # router.py:280
H1, H2 = 256, 128          # router.py:40,44

_KEYS = ("fc1.weight", "fc1.bias", "ln1.weight", "ln1.bias", "fc2.weight", "fc2.bias",
         "ln2.weight", "ln2.bias", "fc3.weight", "fc3.bias")


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _np32(v) -> np.ndarray:
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().numpy()
    return np.ascontiguousarray(v, dtype=np.float32)


class FusedCorpusRouter:
    """CorpusRoutingNN-shaped module backed by one fused CUDA kernel."""

    def __init__(self, input_dim: int, device: Optional[int] = None):
        _lib.require_cuda()
        self._lib = _lib.load()
        self.input_dim = int(input_dim)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.fc1 = SimpleNamespace(in_features=self.input_dim, out_features=H1)
        self._h = None
        self.data_sources: List[str] = []
        self._slot_names: List[str] = []
        self.d_max = 0

    # -- nn.Module-shaped surface ------------------------------------------------
    def load_state_dict(self, sd: Dict[str, object]) -> None:
        missing = [k for k in _KEYS if k not in sd]
        if missing:
            raise KeyError(f"load_state_dict: missing keys {missing}")
        a = {k: _np32(sd[k]) for k in _KEYS}
        if a["fc1.weight"].shape != (H1, self.input_dim) or a["fc2.weight"].shape != (H2, H1) \
                or a["fc3.weight"].reshape(-1).shape != (H2,):
            raise ValueError("load_state_dict: shapes do not match CorpusRoutingNN(%d)" % self.input_dim)
        self._destroy()
        h = _lib.c_void_p()
        ptrs = [a[k].ctypes.data for k in _KEYS]
        _lib.check(self._lib.rr_router_create(self.input_dim, H1, H2, *ptrs, self.device, _lib.ctypes.byref(h)))
        self._h = h

    @classmethod
    def from_state_dict(cls, sd, device: Optional[int] = None) -> "FusedCorpusRouter":
        r = cls(int(_np32(sd["fc1.weight"]).shape[1]), device=device)
        r.load_state_dict(sd)
        return r

    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """self.router(input_tensor): fp32 [P, in_dim] -> logits [P, 1] (CUDA)."""
        self._need_handle()
        with torch.cuda.device(self.device):
            x = x.to(device=f"cuda:{self.device}", dtype=torch.float32).contiguous()
            assert x.dim() == 2 and x.shape[1] == self.input_dim, "router input must be [P, in_dim]"
            out = torch.empty((x.shape[0], 1), dtype=torch.float32, device=x.device)
            _lib.check(self._lib.rr_router_logits(self._h, x.data_ptr(), x.shape[0], out.data_ptr(), _stream_ptr()))
        return out

    # -- fused selection -----------------------------------------------------------
    def set_tables(self, data_sources: Sequence[str], d_max: int, model_of_source: Dict[str, str],
                   centroids: Dict[str, np.ndarray], onehot_id: Dict[str, int], n_onehot: int,
                   scaler=None) -> None:
        """What Router.load_router gathers (router.py:121-151): per-source centroid
        (zero padded to d_max), one-hot id, encoder of each source, optional
        StandardScaler (object with mean_/scale_, or a (mean, scale) pair)."""
        self._need_handle()
        self.data_sources = list(data_sources)
        self.d_max = int(d_max)
        self._slot_names = sorted({model_of_source[c] for c in self.data_sources})
        slots = np.array([self._slot_names.index(model_of_source[c]) for c in self.data_sources], dtype=np.int32)
        cen = np.zeros((len(self.data_sources), self.d_max), dtype=np.float32)
        for i, c in enumerate(self.data_sources):
            v = _np32(centroids[c]).reshape(-1)
            cen[i, :len(v)] = v
        oh = np.array([onehot_id[c] for c in self.data_sources], dtype=np.int32)
        mean_p = scale_p = None
        if scaler is not None:
            mean, scale = (scaler.mean_, scaler.scale_) if hasattr(scaler, "mean_") else scaler
            self._mean = np.ascontiguousarray(mean, dtype=np.float64)
            self._scale = np.ascontiguousarray(scale, dtype=np.float64)
            assert self._mean.shape == (self.input_dim,) and self._scale.shape == (self.input_dim,)
            mean_p, scale_p = self._mean.ctypes.data, self._scale.ctypes.data
        _lib.check(self._lib.rr_router_set_tables(self._h, len(self.data_sources), self.d_max, len(self._slot_names),
                                                  slots.ctypes.data, cen.ctypes.data, oh.ctypes.data, int(n_onehot),
                                                  mean_p, scale_p))

    @property
    def slot_names(self) -> List[str]:
        """Encoder names in the order `select` expects along its slot axis."""
        return list(self._slot_names)

    def select(self, q: torch.Tensor, threshold: float = DEFAULT_THRESHOLD, return_logits: bool = False):
        """q fp32 CUDA [B, d_max] (one encoder) or [B, n_slots, d_max] -> uint8 mask [B, S]
        (1 = source selected); rows follow ``self.data_sources`` order (router.py:282)."""
        self._need_handle()
        if q.dim() == 2:
            q = q.unsqueeze(1)
        assert q.is_cuda and q.dtype == torch.float32 and q.is_contiguous()
        B, n_slots, dm = q.shape
        assert n_slots == len(self._slot_names) and dm == self.d_max, "select: q must be [B, n_slots, d_max]"
        S = len(self.data_sources)
        with torch.cuda.device(self.device):
            mask = torch.empty((B, S), dtype=torch.uint8, device=q.device)
            logits = torch.empty((B, S), dtype=torch.float32, device=q.device) if return_logits else None
            _lib.check(self._lib.rr_router_select(self._h, q.data_ptr(), B, float(threshold), mask.data_ptr(),
                                                  logits.data_ptr() if return_logits else None, _stream_ptr()))
        return (mask, logits) if return_logits else mask

    def select_relevant_sources(self, query_embeddings: Dict[str, np.ndarray],
                                threshold: float = DEFAULT_THRESHOLD) -> List[str]:
        """One query, the reference's signature (router.py:241): dict encoder-name ->
        embedding; returns the selected source names."""
        q = np.zeros((1, len(self._slot_names), self.d_max), dtype=np.float32)
        for i, m in enumerate(self._slot_names):
            e = _np32(query_embeddings[m]).reshape(-1)
            q[0, i, :len(e)] = e  # np.pad(query_embed, (0, d_max - len)), router.py:248
        mask = self.select(torch.from_numpy(q).cuda(self.device), threshold)
        picked = mask[0].cpu().numpy().astype(bool)
        return [c for p, c in zip(picked, self.data_sources) if p]

    # -- plumbing --------------------------------------------------------------------
    def _need_handle(self):
        if self._h is None:
            raise _lib.RRError("FusedCorpusRouter: load_state_dict() has not been called")

    def _destroy(self):
        h, self._h = self._h, None
        if h is not None:
            self._lib.rr_router_destroy(h)

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

"""Boundary B1: the slice of the ``faiss`` module the reference's data sources use.

    import ragroute_b200.faiss_compat as faiss

replaces ``import faiss`` in /root/reference/ragroute/data_source.py:12 for the calls
it makes: ``faiss.read_index(path)`` (:71), ``index.search(x, k)`` positional
(:158, :203) and keyword ``k=k`` (:186), ``faiss.normalize_L2(x)`` in place (:199).
Same argument meaning and error behaviour as FAISS' Python wrapper: ``x`` must be a
2-D float32 array with ``x.shape[1] == index.d``, ``k > 0``; results are
``(D float32 [nq, k] descending, I int64 [nq, k])`` padded with ``(-FLT_MAX, -1)``.

All arithmetic runs in libragroute_b200.so on the GPU; there is no CPU path.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class IndexFlatIP:
    """Exact inner-product index (faiss.IndexFlatIP).

    Rows live on one GPU as two 16-bit planes (bf16 scan operand + remainder, 4 bytes
    per element in total); ``search`` scores on tcgen05 tensor cores and rescoring
    makes the returned ids / scores those of an fp32-exact scan.
    """

    metric_type = METRIC_INNER_PRODUCT
    is_trained = True

    def __init__(self, d: int, device: Optional[int] = None, reserve: int = 0):
        _lib.require_cuda()
        self._lib = _lib.load()
        self.d = int(d)
        self.device = torch.cuda.current_device() if device is None else int(device)
        h = _lib.c_void_p()
        _lib.check(self._lib.rr_index_create(self.d, int(reserve), self.device, _lib.ctypes.byref(h)))
        self._h = h
        self.id_offset = 0

    # -- faiss surface ---------------------------------------------------------
    @property
    def ntotal(self) -> int:
        return int(self._lib.rr_index_ntotal(self._h))

    def add(self, x) -> None:
        """index.add(x): x float32 [n, d], numpy (host) or a CUDA tensor."""
        if isinstance(x, torch.Tensor):
            assert x.dim() == 2 and x.shape[1] == self.d, "add: x must be [n, d]"
            if x.is_cuda:
                assert x.device.index == self.device, "add: tensor on another device"
                x = x.contiguous().float()
                with torch.cuda.device(self.device):
                    _lib.check(self._lib.rr_index_add_f32(self._h, x.data_ptr(), x.shape[0], 1, _stream_ptr()))
                    # the rows are consumed asynchronously on the current stream
                    x.record_stream(torch.cuda.current_stream())
                return
            x = x.numpy()
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d, "add: x must be [n, d]"
        _lib.check(self._lib.rr_index_add_f32(self._h, x.ctypes.data, x.shape[0], 0, None))

    def search(self, x, k: int, normalize: bool = False) -> Tuple[np.ndarray, np.ndarray]:
        """index.search(x, k): numpy in, numpy out (the reference's convention).

        normalize=True folds ``faiss.normalize_L2(x)`` (data_source.py:199) into the same host round trip:
        the queries are L2-normalised on the device before the scan and x is updated in place, exactly
        what the reference's two calls leave behind."""
        xin = x
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2, "search: x must be 2-D"
        nq, d = x.shape
        assert d == self.d, f"search: x.shape[1]={d} but index.d={self.d}"
        assert k > 0, "search: k must be > 0"
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        if normalize:
            _lib.check(self._lib.rr_index_search_host_ex(self._h, x.ctypes.data, nq, int(k), _lib.RR_SEARCH_NORMALIZE_L2,
                                                         D.ctypes.data, I.ctypes.data, x.ctypes.data))
            if x is not xin and isinstance(xin, np.ndarray):
                xin[...] = x
        else:
            _lib.check(self._lib.rr_index_search_host(self._h, x.ctypes.data, nq, int(k), D.ctypes.data, I.ctypes.data))
        return D, I

    def reconstruct_tensor(self, i0: int, n: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Rows [i0, i0 + n) as an fp32 CUDA tensor [n, d], bit-exact (asynchronous on the current stream)."""
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty((n, self.d), dtype=torch.float32, device="cuda")
            assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.shape == (n, self.d)
            _lib.check(self._lib.rr_index_reconstruct(self._h, int(i0), int(n), out.data_ptr(), _stream_ptr()))
        return out

    def reconstruct_n(self, i0: int = 0, n: int = -1) -> np.ndarray:
        n = self.ntotal - i0 if n < 0 else n
        return self.reconstruct_tensor(i0, n).cpu().numpy()

    def centroid(self) -> np.ndarray:
        """Column mean of the stored rows, fp32 [d] (what the router's <corpus>_stats.json holds, router.py:136-151)."""
        out = np.empty(self.d, dtype=np.float32)
        _lib.check(self._lib.rr_index_centroid(self._h, out.ctypes.data))
        return out

    def reconstruct(self, i: int) -> np.ndarray:
        return self.reconstruct_n(i, 1)[0]

    def reset(self) -> None:
        _lib.check(self._lib.rr_index_reset(self._h))

    # -- tensor fast path (batched callers, no host copies, no sync) -------------
    def search_tensor(self, x: torch.Tensor, k: int, mask: Optional[torch.Tensor] = None,
                      mask_col: int = 0, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                      out_col: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """x float32 CUDA [nq, d] -> (D [nq, k], I [nq, k]) CUDA tensors.

        mask: optional uint8 CUDA [nq, S]; only queries with mask[:, mask_col] != 0 are
        searched (routed search), the others return padding rows.
        out: optional pre-allocated (D_all [nq, W], I_all [nq, W]); the k results are
        written at columns [out_col, out_col + k) of each row."""
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
        nq, d = x.shape
        assert d == self.d, f"search: x.shape[1]={d} but index.d={self.d}"
        assert k > 0, "search: k must be > 0"
        with torch.cuda.device(self.device):
            if out is None:
                D = torch.empty((nq, k), dtype=torch.float32, device=x.device)
                I = torch.empty((nq, k), dtype=torch.int64, device=x.device)
                d_ptr, i_ptr, stride = D.data_ptr(), I.data_ptr(), k
            else:
                D, I = out
                assert D.is_contiguous() and I.is_contiguous() and D.shape == I.shape and D.shape[0] == nq
                stride = D.shape[1]
                assert out_col + k <= stride
                d_ptr = D.data_ptr() + 4 * out_col
                i_ptr = I.data_ptr() + 8 * out_col
            m_ptr, m_stride = None, 0
            if mask is not None:
                assert mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous() and mask.shape[0] == nq
                m_stride = mask.shape[1] if mask.dim() == 2 else 1
                m_ptr = mask.data_ptr() + mask_col
            _lib.check(self._lib.rr_index_search_ex(self._h, x.data_ptr(), nq, int(k), m_ptr, m_stride,
                                                    d_ptr, i_ptr, stride, _stream_ptr()))
        return D, I

    # -- knobs -------------------------------------------------------------------
    def set_mode(self, mode: str) -> None:
        """'auto' (bf16 tensor-core scan + exact rescoring + certificate) or 'exact'."""
        val = {"auto": _lib.RR_MODE_AUTO, "exact": _lib.RR_MODE_EXACT}[mode]
        _lib.check(self._lib.rr_index_set_param(self._h, b"mode", float(val)))

    def set_param(self, name: str, value: float) -> None:
        _lib.check(self._lib.rr_index_set_param(self._h, name.encode(), float(value)))

    def set_id_offset(self, offset: int) -> None:
        _lib.check(self._lib.rr_index_set_id_offset(self._h, int(offset)))
        self.id_offset = int(offset)

    def reserve(self, rows: int) -> None:
        _lib.check(self._lib.rr_index_reserve(self._h, int(rows)))

    def stats(self, reset: bool = False) -> dict:
        st = _lib.IndexStats()
        _lib.check(self._lib.rr_index_get_stats(self._h, _lib.ctypes.byref(st), 1 if reset else 0))
        return {f: getattr(st, f) for f, _ in st._fields_}

    def debug_tc_candidates(self, x: torch.Tensor):
        """Raw candidate lists of the bf16 tensor-core scan (testing hook)."""
        nq = x.shape[0]
        cap = 1 << 16
        with torch.cuda.device(self.device):
            cs = torch.empty((nq, cap), dtype=torch.float32, device=x.device)
            cr = torch.empty((nq, cap), dtype=torch.int32, device=x.device)
            nl, dp = _lib.c_int(0), _lib.c_int(0)
            _lib.check(self._lib.rr_index_debug_tc_scan(self._h, x.data_ptr(), nq, cs.data_ptr(), cr.data_ptr(),
                                                        cap, _lib.ctypes.byref(nl), _lib.ctypes.byref(dp),
                                                        _stream_ptr()))
            n = nl.value * dp.value
            return cs[:, :n].clone(), cr[:, :n].clone(), nl.value, dp.value

    def close(self) -> None:
        """Free the device memory now (the handle is unusable afterwards)."""
        h, self._h = getattr(self, "_h", None), None
        if h is not None and _lib is not None:
            try:
                self._lib.rr_index_destroy(h)
            except Exception:
                pass

    def __del__(self):
        self.close()


def normalize_L2(x: np.ndarray) -> None:
    """faiss.normalize_L2(x): in place on a C-contiguous float32 [n, d] array
    (data_source.py:199); rows with zero norm are left untouched."""
    assert isinstance(x, np.ndarray) and x.dtype == np.float32 and x.ndim == 2 and x.flags.c_contiguous, \
        "normalize_L2 needs a C-contiguous float32 [n, d] array"
    _lib.require_cuda()
    t = torch.from_numpy(x).cuda()
    _lib.check(_lib.load().rr_normalize_l2(t.data_ptr(), x.shape[0], x.shape[1], _stream_ptr()))
    x[...] = t.cpu().numpy()


# ---------------------------------------------------------------------------
# faiss.read_index / write_index for flat indexes (upstream faiss 1.7.4 layout, see csrc/rr_io.cu:
# fourcc "IxFI" | "IxFl", header d:int32 ntotal:int64 dummy:int64 dummy:int64 is_trained:uint8
# metric:int32, then uint64 count-of-floats and the raw fp32 rows).  Both run natively: the rows stream
# disk <-> pinned host <-> device with double buffering.
# ---------------------------------------------------------------------------
def write_index(index: IndexFlatIP, path: str) -> None:
    _lib.check(index._lib.rr_index_write(index._h, os.fsencode(path)))


def read_index(path: str, device: Optional[int] = None) -> IndexFlatIP:
    """faiss.read_index(path) (data_source.py:71) for a flat inner-product index file; anything else
    raises (the reference's data sources are IndexFlatIP, north_star)."""
    _lib.require_cuda()
    lib = _lib.load()
    dev = torch.cuda.current_device() if device is None else int(device)
    h = _lib.c_void_p()
    _lib.check(lib.rr_index_read(os.fsencode(path), dev, _lib.ctypes.byref(h)))
    index = IndexFlatIP.__new__(IndexFlatIP)
    index._lib, index._h, index.device, index.id_offset = lib, h, dev, 0
    index.d = int(lib.rr_index_dim(h))
    return index

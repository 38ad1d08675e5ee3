"""``PriorityTree`` -- drop-in for the reference's ``priority_tree.PriorityTree``
(priority_tree.py:4-45) backed by the HBM-resident float64 sum tree (K3).

Same constructor, same ``update(idxes, td_error)`` / ``sample(num_samples) ->
(idxes int64, is_weights float64)`` surface with NumPy arrays in and out (what
the reference's ``ReplayBuffer`` passes), plus ``*_device`` variants that take
and return CUDA tensors without any host round trip (what our HBM-resident
replay uses).  ``ptree`` returns the node array (host copy), ``num_layers`` is
the reference attribute.

``sample`` draws its uniforms from NumPy's global legacy generator exactly like
the reference (``np.random.uniform(0, interval, n)`` == ``interval *
random_sample(n)``), so ``np.random.seed(s)`` reproduces the reference's sampled
indices bit for bit; ``sample_device`` uses the on-device Philox stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class PriorityTree:
    def __init__(self, capacity, prio_exponent, is_exponent, device=None, seed: int = 0):
        _lib.require_device()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.capacity = int(capacity)
        self.prio_exponent = prio_exponent
        self.is_exponent = is_exponent
        self.seed = int(seed)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().r2d2_tree_create(self.capacity, float(prio_exponent), float(is_exponent),
                                                   C.byref(h)))
        self._h = h
        self.num_layers = _lib.lib().r2d2_tree_num_layers(h)
        self.num_nodes = _lib.lib().r2d2_tree_num_nodes(h)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().r2d2_tree_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------ device API
    def update_device(self, idxes: torch.Tensor, td_error: torch.Tensor, old_ptr: int = -1, cur_ptr: int = 0,
                      seq_per_block: int = 1) -> None:
        assert idxes.dtype == torch.int64 and td_error.dtype == torch.float32
        assert idxes.numel() == td_error.numel()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().r2d2_tree_update(self._h, _lib.ptr(idxes), _lib.ptr(td_error), idxes.numel(),
                                                   int(old_ptr), int(cur_ptr), int(seq_per_block),
                                                   _lib.stream_ptr()))

    def set_leaves_device(self, idxes: torch.Tensor, leaves: torch.Tensor) -> None:
        assert idxes.dtype == torch.int64 and leaves.dtype == torch.float64
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().r2d2_tree_set_leaves(self._h, _lib.ptr(idxes), _lib.ptr(leaves), idxes.numel(),
                                                       _lib.stream_ptr()))

    def sample_device(self, num_samples: int, unit_uniforms: torch.Tensor | None = None, want_f64: bool = False):
        n = int(num_samples)
        idx = torch.empty(n, dtype=torch.int64, device=self.device)
        w32 = torch.empty(n, dtype=torch.float32, device=self.device)
        w64 = torch.empty(n, dtype=torch.float64, device=self.device) if want_f64 else None
        if unit_uniforms is not None:
            assert unit_uniforms.dtype == torch.float64 and unit_uniforms.numel() == n
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().r2d2_tree_sample(self._h, n, _lib.ptr(unit_uniforms), self.seed, _lib.ptr(idx),
                                                   _lib.ptr(w32), _lib.ptr(w64), _lib.stream_ptr()))
        return (idx, w32, w64) if want_f64 else (idx, w32)

    def nodes_device(self) -> torch.Tensor:
        """Zero-copy float64 view of the node array in HBM."""
        addr = _lib.lib().r2d2_tree_nodes(self._h)
        return _from_device_ptr(addr, self.num_nodes, torch.float64, self.device)

    # ------------------------------------------------------------------ reference (NumPy) surface
    def update(self, idxes: np.ndarray, td_error: np.ndarray) -> None:
        idx = torch.from_numpy(np.ascontiguousarray(idxes, dtype=np.int64)).to(self.device)
        td = torch.from_numpy(np.ascontiguousarray(td_error, dtype=np.float32)).to(self.device)
        self.update_device(idx, td)

    def sample(self, num_samples: int):
        r = torch.from_numpy(np.random.random_sample(num_samples)).to(self.device)
        idx, _, w64 = self.sample_device(num_samples, r, want_f64=True)
        return idx.cpu().numpy(), w64.cpu().numpy()

    @property
    def ptree(self) -> np.ndarray:
        return self.nodes_device().cpu().numpy()


class _CudaArrayView:
    def __init__(self, addr, nbytes, typestr, shape):
        self.__cuda_array_interface__ = {"data": (addr, False), "shape": shape, "typestr": typestr, "version": 3,
                                         "strides": None}


def _from_device_ptr(addr: int, numel: int, dtype: torch.dtype, device) -> torch.Tensor:
    typestr = {torch.float64: "<f8", torch.float32: "<f4", torch.int64: "<i8", torch.int32: "<i4",
               torch.uint8: "|u1", torch.bfloat16: "<i2"}[dtype]
    t = torch.as_tensor(_CudaArrayView(addr, 0, typestr, (int(numel),)), device=device)
    return t.view(torch.bfloat16) if dtype == torch.bfloat16 else t

"""ctypes binding of libr2d2_b200.so -- the C-ABI boundary (include/r2d2_b200.h).

There is NO fallback: if the shared object is missing or the device is not a
Blackwell (sm_100) part, importing a product module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libr2d2_b200.so")

_lib = None


class R2D2Error(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise R2D2Error(
                f"{LIB_PATH} is missing: build it with `python -m r2d2_b200.build` "
                "(or __graft_entry__.build()); there is no CPU/PyTorch fallback")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise R2D2Error(f"r2d2_b200 error {rc}: {lib().r2d2_last_error().decode()}")


def require_device() -> None:
    check(lib().r2d2_device_ok())


p = C.c_void_p
i64, i32, u64, f64, f32 = C.c_int64, C.c_int32, C.c_uint64, C.c_double, C.c_float

# name -> (restype, argtypes); kept in one table so tests can verify that every
# symbol declared in include/r2d2_b200.h is exported.
SIGNATURES = {
    "r2d2_last_error": (C.c_char_p, []),
    "r2d2_abi_version": (C.c_int, []),
    "r2d2_device_ok": (C.c_int, []),
    "r2d2_tree_create": (C.c_int, [i64, f64, f64, C.POINTER(p)]),
    "r2d2_tree_destroy": (C.c_int, [p]),
    "r2d2_tree_num_layers": (C.c_int, [p]),
    "r2d2_tree_num_nodes": (i64, [p]),
    "r2d2_tree_nodes": (p, [p]),
    "r2d2_tree_update": (C.c_int, [p, p, p, i64, i64, i64, i64, p]),
    "r2d2_tree_set_leaves": (C.c_int, [p, p, p, i64, p]),
    "r2d2_tree_sample": (C.c_int, [p, i64, p, u64, p, p, p, p]),
    "r2d2_td_loss": (C.c_int, [p] * 8 + [C.c_int, C.c_int] + [p] * 6),
    "r2d2_net_param_layout": (C.c_int, [C.c_int, C.c_int, C.POINTER(i64)]),
    "r2d2_net_create": (C.c_int, [C.c_int] * 6 + [C.POINTER(p)]),
    "r2d2_net_destroy": (C.c_int, [p]),
    "r2d2_net_rows_capacity": (C.c_int, [p]),
    "r2d2_net_ku": (C.c_int, [p]),
    "r2d2_net_pack": (C.c_int, [p, C.c_int, p, p]),
    "r2d2_net_forward": (C.c_int, [p, C.c_int] + [p] * 11),
    "r2d2_net_forward_pair": (C.c_int, [p] * 14),
    "r2d2_net_backward": (C.c_int, [p, p, p, p, p]),
    "r2d2_set_persistent_recurrence": (C.c_int, [C.c_int]),
    "r2d2_set_pair_gemm": (C.c_int, [C.c_int]),
    "r2d2_set_cluster_recurrence": (C.c_int, [C.c_int]),
    "r2d2_debug_cluster_capacity": (C.c_int, []),
    "r2d2_config_epoch": (C.c_int, []),
    "r2d2_debug_rec_trace": (C.c_int, [p]),
    "r2d2_debug_rec_trace_bwd": (C.c_int, [p]),
    "r2d2_net_debug_ptr": (p, [p, C.c_int, C.c_char_p]),
    "r2d2_replay_create": (C.c_int, [C.c_int] * 8 + [C.POINTER(p)]),
    "r2d2_replay_destroy": (C.c_int, [p]),
    "r2d2_replay_layout": (C.c_int, [p, C.POINTER(i64)]),
    "r2d2_replay_ingest": (C.c_int, [p, C.c_int, p, i64, p]),
    "r2d2_replay_gather": (C.c_int, [p, p, p, C.c_int, C.c_int] + [p] * 13),
    "r2d2_replay_gather_s2d": (C.c_int, [p, p, p, C.c_int, C.c_int] + [p] * 13),
    "r2d2_net_s2d_buffer": (p, [p]),
    "r2d2_net_s2d_buffer_at": (p, [p, C.c_int]),
    "r2d2_net_select_s2d": (C.c_int, [p, C.c_int]),
    "r2d2_event_record": (C.c_int, [p, p]),
    "r2d2_net_shadow_gate": (C.c_int, [p, p]),
    "r2d2_net_shadow_gate_reset": (C.c_int, [p, p]),
    "r2d2_replay_set_copy_smem": (C.c_int, [p, C.c_int]),
    "r2d2_net_state_after": (C.c_int, [p, C.c_int, C.c_int, p, p]),
    "r2d2_net_set_dense_grads_event": (C.c_int, [p, p]),
    "r2d2_set_fast_math": (C.c_int, [C.c_int]),
    "r2d2_debug_gemm2": (C.c_int, [C.c_int] * 6 + [p] * 5 + [C.c_int, p]),
    "r2d2_debug_gemm3": (C.c_int, [C.c_int] * 5 + [p] * 5 + [C.c_int, p]),
    "r2d2_debug_shift_probe": (C.c_int, [p, p, p, C.c_int, C.c_int, p]),
    "r2d2_debug_ts_probe": (C.c_int, [p, p, p, p]),
    "r2d2_debug_mma_rate": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, p, p]),
    "r2d2_clip_adam": (C.c_int, [p, p, p, p, i64, p, p, f32, f32, f32, f32, f32, i64, p, p]),
    "r2d2_clip_adam_dev": (C.c_int, [p, p, p, p, i64, p, p, p, f32, f32, f32, f32, f32, p, p, p]),
    "r2d2_clip_adam_step": (C.c_int, [p, p, p, p, i64, p, p, p, f32, f32, f32, f32, f32, p, p, p]),
    "r2d2_dp_ctl_bytes": (C.c_size_t, []),
    "r2d2_dp_create": (C.c_int, [C.c_int, C.c_int, p, C.c_ulonglong, p, p]),
    "r2d2_dp_destroy": (None, [p]),
    "r2d2_dp_allreduce": (C.c_int, [p, C.c_longlong, C.c_longlong, C.c_int, p, C.c_longlong, p, C.c_int, C.c_int, C.c_int, p]),
    "r2d2_dp_is_post": (C.c_int, [p, p, C.c_longlong, p, C.c_int, p]),
    "r2d2_dp_is_apply": (C.c_int, [p, C.c_double, p, C.c_int, p, p]),
    "r2d2_dp_error": (C.c_int, [p, p]),
}


def _declare(l: C.CDLL) -> None:
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args


def ptr(t) -> int:
    """data_ptr of a (contiguous, CUDA) torch tensor or None."""
    if t is None:
        return None
    assert t.is_contiguous(), "r2d2_b200 kernels take contiguous tensors"
    return t.data_ptr()


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream on the CURRENT device (callers run inside `on_device`)."""
    import torch
    return torch.cuda.current_stream().cuda_stream


def on_device(method):
    """Decorator for methods of objects with a `.device`: run the body with that CUDA device current, so that the stream
    handed to the C ABI, the kernels it launches and the buffers it touches all belong to the same device even when the
    caller never called torch.cuda.set_device (e.g. Learner(device='cuda:1'))."""
    import functools

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        import torch
        idx = self.device.index
        if idx is None or torch.cuda.current_device() == idx:
            return method(self, *args, **kwargs)
        with torch.cuda.device(idx):
            return method(self, *args, **kwargs)
    return wrapper

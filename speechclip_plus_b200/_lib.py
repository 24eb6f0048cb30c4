"""ctypes binding of libscp_b200.so (include/scp_b200.h).

The CUDA library is the only backend: importing this module on a machine where the shared object is missing (and
cannot be built) raises, and every wrapper raises :class:`ScpError` on a non-zero status -- nothing falls back to
PyTorch or the CPU.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libscp_b200.so")

SCP_F32, SCP_F16, SCP_BF16 = 0, 1, 2
SCP_NORM_NONE, SCP_NORM_LAYERNORM, SCP_NORM_L2_FRAME, SCP_NORM_UTT_MEAN = 0, 1, 2, 3
SCP_MAX_LAYERS = 32
SCP_MAX_MASKED = 8
SCP_MAX_PACKED = 16


class ScpError(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None

# name -> (restype, argtypes); mirrors include/scp_b200.h one to one (checked by tests/test_abi.py)
SIGNATURES = {
    "scp_version": (c_int, []),
    "scp_last_error_string": (c_char_p, [c_int]),
    "scp_num_launches": (c_int, []),
    "scp_wsum_fwd": (c_int, [POINTER(c_void_p), c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_int,
                             c_void_p, c_int, c_float, c_void_p, c_void_p, c_int, c_void_p]),
    "scp_wsum_utt_scale": (c_int, [POINTER(c_void_p), c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_int,
                                   c_void_p, c_void_p]),
    "scp_wsum_bwd_workspace_bytes": (c_size_t, [c_int, c_int64, c_int64, c_int64]),
    "scp_wsum_bwd": (c_int, [POINTER(c_void_p), c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_int,
                             c_void_p, c_int, c_float, c_void_p, c_void_p, c_int, c_void_p, POINTER(c_void_p),
                             c_void_p, c_size_t, c_void_p]),
    "scp_kwbn_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "scp_kwbn_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                             c_void_p]),
    "scp_kwbn_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "scp_vq_padded_vocab": (c_int64, [c_int64]),
    "scp_vq_prepare_table": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "scp_vq_fwd_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "scp_vq_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                           POINTER(c_int32), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "scp_vq_saved_probs_bytes": (c_size_t, [c_int64, c_int64]),
    "scp_vq_fwd_save_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "scp_vq_bwd_saved_available": (c_int, [c_int64, c_int64, c_int64]),
    "scp_vq_bwd_saved_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int]),
    "scp_vq_fwd_save": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                POINTER(c_int32), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "scp_vq_bwd_saved": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, POINTER(c_int32), c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_size_t, c_void_p]),
    "scp_vq_bwd_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "scp_vq_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p, POINTER(c_int32), c_int, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_size_t, c_void_p]),
    "scp_vq_dense_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "scp_vq_dense_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, POINTER(c_int32), c_int, c_void_p,
                                 c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_size_t, c_void_p]),
    "scp_vq_dense_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    "scp_kw_splice_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64,
                                  c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "scp_kw_splice_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                  c_void_p]),
    "scp_keypadding_mask": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "scp_cif_plan": (c_int, [c_void_p, c_int64, c_int64, c_float, c_int, c_void_p, c_void_p, c_void_p]),
    "scp_cif_fire_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float, c_int64,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    "scp_cif_tail": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_float, c_float, c_int, c_void_p,
                             c_void_p]),
    "scp_cif_fire_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float,
                                 c_int64, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_size_t,
                                 c_void_p]),
    "scp_pack_bytes": (c_size_t, [c_int, c_int64, c_int64]),
    "scp_l2norm_pack": (c_int, [POINTER(c_void_p), c_int, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "scp_l2norm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "scp_grad_pack": (c_int, [POINTER(c_void_p), POINTER(c_int64), c_int, c_float, c_void_p, c_void_p]),
    "scp_adam_packed": (c_int, [POINTER(c_void_p), POINTER(c_int64), c_int, c_void_p, c_int, c_int64, c_float, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_float, c_void_p]),
    "scp_p2p_buffer_bytes": (c_size_t, [c_int, c_size_t]),
    "scp_p2p_allgather": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_int, c_int, c_size_t, c_void_p, c_void_p,
                                  c_void_p]),
    "scp_p2p_allgather_segments": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_int, c_int, c_size_t, c_void_p,
                                           POINTER(c_int64), c_int, c_void_p, c_void_p]),
    "scp_nce_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "scp_nce_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_float, c_float, c_int, c_int,
                            c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "scp_nce_fwd_local": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_float, c_float, c_int, c_int64,
                                  c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "scp_nce_loss_from_stats": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_float, c_float, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p, c_void_p]),
    "scp_nce_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_float, c_float, c_int, c_int,
                            c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_size_t, c_void_p]),
}


def load(build_if_missing: bool = True):
    """Load (building first if the .so is absent or stale and nvcc is available) and return the ctypes library."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if build_if_missing:
            try:
                from . import build as _build
                if _build.needs_build():
                    _build.build()
            except Exception as exc:  # stale-but-present library is still usable; a missing one is fatal below
                if not os.path.exists(LIB_PATH):
                    raise ScpError(f"libscp_b200.so is missing and could not be built: {exc}") from exc
        if not os.path.exists(LIB_PATH):
            raise ScpError(f"{LIB_PATH} not found: run `python -m speechclip_plus_b200.build` (needs nvcc); "
                           "there is no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().scp_last_error_string(status)
        raise ScpError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")


def num_launches() -> int:
    return int(load().scp_num_launches())


def dtype_code(dtype) -> int:
    import torch
    if dtype == torch.float32:
        return SCP_F32
    if dtype == torch.float16:
        return SCP_F16
    if dtype == torch.bfloat16:
        return SCP_BF16
    raise ScpError(f"unsupported dtype {dtype} (float32, float16, bfloat16)")


def ptr(t) -> c_void_p:
    """Device pointer of a tensor (None -> NULL)."""
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())


def ptr_array(tensors) -> ctypes.Array:
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def stream_ptr(device=None) -> c_void_p:
    import torch
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, what: str) -> None:
    if not t.is_cuda:
        raise ScpError(f"{what}: expected a CUDA tensor, got device {t.device}; speechclip_plus_b200 has no CPU path")

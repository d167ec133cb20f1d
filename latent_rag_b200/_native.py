"""ctypes binding of liblatentknn.so (include/latentknn.h).

The library is the product: if it is missing, or if there is no CUDA device, every
compute entry point raises -- there is no CPU or PyTorch fallback on this path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_void_p, byref

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "liblatentknn.so")

LK_ABI_VERSION = 2
LK_F32, LK_BF16 = 0, 1
LK_HOST, LK_DEVICE = 0, 1
LK_COSINE, LK_EUCLIDEAN, LK_MAHALANOBIS = 0, 1, 2
LK_KERNEL_AUTO, LK_KERNEL_SIMT, LK_KERNEL_UMMA = 0, 1, 2
LK_AE_DAE, LK_AE_CAE, LK_AE_VAE_MU = 0, 1, 2
LK_MAX_K = 4096       # lk_index_search / lk_merge_topk / lk_maxsim_rerank
LK_MAX_K_FUSED = 128  # one fused pass; above it lk_index_search runs the slab search.  Also the peer exchange's limit
LK_MAX_WORLD = 16
LK_IPC_HANDLE_BYTES = 64

METRICS = {"cosine": LK_COSINE, "euclidean": LK_EUCLIDEAN, "mahalanobis": LK_MAHALANOBIS}
KERNELS = {"auto": LK_KERNEL_AUTO, "simt": LK_KERNEL_SIMT, "umma": LK_KERNEL_UMMA}
STORAGE = {"bf16": LK_BF16, "fp32": LK_F32}

# every symbol include/latentknn.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("lk_abi_version", c_int, []),
    ("lk_last_error", c_char_p, []),
    ("lk_device_count", c_int, [POINTER(c_int)]),
    ("lk_launch_count", c_int64, []),
    ("lk_index_create", c_int, [POINTER(c_void_p), c_int, c_int64, c_int, c_int, c_int, POINTER(c_double)]),
    ("lk_index_add", c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p]),
    ("lk_index_reserve", c_int, [c_void_p, c_int64, c_void_p]),
    ("lk_index_size", c_int, [c_void_p, POINTER(c_int64), POINTER(c_int)]),
    ("lk_index_destroy", c_int, [c_void_p]),
    ("lk_index_search", c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_void_p, c_void_p, c_int,
                                c_int64, c_int, c_void_p]),
    ("lk_index_check", c_int, [c_void_p]),
    ("lk_index_last_timing", c_int, [c_void_p, POINTER(c_float), POINTER(c_float)]),
    ("lk_index_set_timing", c_int, [c_void_p, c_int]),
    ("lk_index_storage_bytes", c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64)]),
    ("lk_index_export", c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    ("lk_index_import", c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p]),
    ("lk_merge_topk", c_int, [c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                              c_void_p]),
    ("lk_ae_create", c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                             c_void_p]),
    ("lk_ae_set_kernel", c_int, [c_void_p, c_int]),
    ("lk_ae_set_precision", c_int, [c_void_p, c_int]),
    ("lk_ae_encode", c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_int, c_void_p]),
    ("lk_ae_destroy", c_int, [c_void_p]),
    ("lk_maxsim_rerank", c_int, [c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p,
                                 c_void_p, c_void_p]),
    ("lk_retrieval_metrics", c_int, [c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                     c_void_p, c_void_p, c_void_p]),
    ("lk_rank_positive", c_int, [c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    ("lk_comm_create", c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int64, c_int]),
    ("lk_comm_ipc_handle", c_int, [c_void_p, c_void_p]),
    ("lk_comm_open_peers", c_int, [c_void_p, c_void_p]),
    ("lk_comm_attach_local", c_int, [c_void_p, c_int, c_void_p]),
    ("lk_comm_exchange_merge", c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    ("lk_comm_begin", c_int, [c_void_p]),
    ("lk_comm_publish", c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    ("lk_comm_collect", c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    ("lk_comm_check", c_int, [c_void_p]),
    ("lk_comm_destroy", c_int, [c_void_p]),
    ("lk_bert_create", c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    ("lk_bert_set_precision", c_int, [c_void_p, c_int]),
    ("lk_bert_encode", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_int, c_void_p]),
    ("lk_bert_check", c_int, [c_void_p]),
    ("lk_bert_destroy", c_int, [c_void_p]),
    ("lk_linear_forward", c_int, [c_int, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                                  c_void_p]),
]


class NativeError(RuntimeError):
    """A liblatentknn call failed (message from lk_last_error)."""


_lib = None


def load() -> ctypes.CDLL:
    """dlopen liblatentknn.so once; raise (never fall back) when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} is missing. Build it with `python -m latent_rag_b200.build` "
            "(nvcc, sm_100a). latent_rag_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.lk_abi_version() != LK_ABI_VERSION:
        raise NativeError(f"liblatentknn ABI {lib.lk_abi_version()} != {LK_ABI_VERSION} expected by this package")
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().lk_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise NativeError(f"{what} failed (status {rc}): {last_error()}")


def device_count() -> int:
    n = c_int(0)
    rc = load().lk_device_count(byref(n))
    return n.value if rc == 0 else 0


def require_device() -> int:
    n = device_count()
    if n < 1:
        raise NativeError(
            "no CUDA device visible: latent_rag_b200 runs only on B200 (sm_100a) GPUs and has no CPU fallback "
            f"({last_error()})"
        )
    return n


def launch_count() -> int:
    return int(load().lk_launch_count())

"""ctypes binding of ``libcmwdense.so`` -- the C ABI declared in ``include/cmw_dense.h``.

There is no CPU fallback: if the shared library is missing (or no B200 is visible when a
compute entry point is called) this module raises.  Build the library with
``python -m cmw_rag_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libcmwdense.so")

METRIC_COSINE = 0
METRIC_IP = 1
MODE_F32_EXACT = 0
MODE_BF16 = 1
ALGO_AUTO = 0 << 8
ALGO_SCAN = 1 << 8
ALGO_GEMM = 2 << 8
ALGO_GEMM_TF32 = 3 << 8
SLABS_SAFE = 1 << 16
KPRIME_MAX = 1 << 17
STORE_F32 = 1
STORE_BF16 = 2
STORE_F16 = 4
FLAG_UNCERTIFIED = 1
FLAG_PEER_TIMEOUT = 2
HOST_SLOTS = 4
MAX_K = 1024

METRICS = {"cosine": METRIC_COSINE, "ip": METRIC_IP, METRIC_COSINE: METRIC_COSINE, METRIC_IP: METRIC_IP}
MODES = {"f32": MODE_F32_EXACT, "exact": MODE_F32_EXACT, "bf16": MODE_BF16,
         MODE_F32_EXACT: MODE_F32_EXACT, MODE_BF16: MODE_BF16}
ALGOS = {"auto": ALGO_AUTO, "scan": ALGO_SCAN, "gemm": ALGO_GEMM, None: ALGO_AUTO,
         "gemm_tf32": ALGO_GEMM_TF32, "scan_safe": ALGO_SCAN | SLABS_SAFE, "gemm_safe": ALGO_GEMM | SLABS_SAFE}


class StoreInfo(ctypes.Structure):
    _fields_ = [
        ("device", c_int32),
        ("dim", c_int32),
        ("flags", c_uint32),
        ("sm_count", c_int32),
        ("capacity_rows", c_int64),
        ("rows", c_int64),
        ("live_rows", c_int64),
        ("id_offset", c_int64),
        ("hbm_bytes", c_int64),
    ]


# name -> (restype, argtypes); every symbol include/cmw_dense.h declares
SIGNATURES = {
    "cmw_last_error": (c_char_p, []),
    "cmw_abi_version": (c_int, []),
    "cmw_store_create": (c_int, [c_int, c_int, c_int64, c_uint32, c_int64, POINTER(c_void_p)]),
    "cmw_store_destroy": (c_int, [c_void_p]),
    "cmw_store_get_info": (c_int, [c_void_p, POINTER(StoreInfo)]),
    "cmw_store_append_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "cmw_store_append_host_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64]),
    "cmw_store_copy_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "cmw_store_tombstone": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "cmw_store_tombstone_host": (c_int, [c_void_p, c_void_p, c_int64]),
    "cmw_store_kb_gid_dev": (c_void_p, [c_void_p]),
    "cmw_store_read_rows_f32": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "cmw_search_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
    "cmw_search": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cmw_search_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                c_void_p]),
    "cmw_search_host_submit": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p, POINTER(c_int)]),
    "cmw_search_host_wait": (c_int, [c_void_p, c_int]),
    "cmw_multivector": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_int, c_int] + [c_void_p] * 11 + [c_void_p]),
    "cmw_merge_topk": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "cmw_peer_buffer_bytes": (c_size_t, [c_int, c_int, c_int]),
    "cmw_peer_alloc": (c_int, [c_int, c_size_t, POINTER(c_void_p), c_void_p]),
    "cmw_peer_open": (c_int, [c_int, c_void_p, POINTER(c_void_p)]),
    "cmw_peer_close": (c_int, [c_void_p]),
    "cmw_peer_free": (c_int, [c_void_p]),
    "cmw_exchange_merge": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_uint32,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cmw_exchange_merge_ex": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_uint32,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                      c_void_p]),
    "cmw_search_filter": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                  c_void_p]),
    "cmw_shard_kth": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cmw_shard_block_bytes": (c_size_t, [c_int, c_int]),
    "cmw_search_finish": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_void_p]),
    "cmw_shard_merge": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "cmw_shard_merge_ex": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "cmw_peer_gather_bytes": (c_size_t, [c_int, c_size_t]),
    "cmw_peer_gather": (c_int, [POINTER(c_void_p), c_int, c_int, c_size_t, c_void_p, c_size_t, c_uint32, c_int,
                                c_void_p, POINTER(c_void_p), c_void_p]),
    "cmw_kernel_launches": (c_int64, []),
    "cmw_profile_enable": (c_int, [c_int]),
    "cmw_profile_read": (c_int, [POINTER(c_double), POINTER(c_int64), c_int]),
    "cmw_set_option": (c_int, [c_char_p, c_double]),
    "cmw_get_option": (c_double, [c_char_p]),
}

_lib = None


class NativeError(RuntimeError):
    """An entry point of libcmwdense.so returned a negative status."""


def lib() -> ctypes.CDLL:
    """Load libcmwdense.so (once).  Raises if it has not been built -- no fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m cmw_rag_b200.build` "
                "(nvcc, sm_100a).  cmw_rag_b200 has no CPU or PyTorch fallback."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError = a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().cmw_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise NativeError(f"{what} failed (status {rc}): {last_error()}")


def kernel_launches() -> int:
    return int(lib().cmw_kernel_launches())


PHASES = ("filter", "compact", "finalize", "prep", "shard_kth", "shard_merge")


def profile_enable(on: bool = True) -> None:
    lib().cmw_profile_enable(1 if on else 0)


def profile_read() -> dict:
    """{phase: (milliseconds, kernel launches)} accumulated since the last enable/read."""
    ms = (c_double * len(PHASES))()
    cnt = (c_int64 * len(PHASES))()
    lib().cmw_profile_read(ms, cnt, len(PHASES))
    return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(PHASES)}


def set_option(name: str, value: float) -> None:
    check(lib().cmw_set_option(name.encode(), float(value)), f"cmw_set_option({name})")


def get_option(name: str) -> float:
    return float(lib().cmw_get_option(name.encode()))

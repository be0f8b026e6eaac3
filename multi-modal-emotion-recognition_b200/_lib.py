"""ctypes binding of libmmer_sm100.so (include/mmer.h).  No CPU fallback: if the library is
missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch  # noqa: F401  (loads libcudart.so.12 before our library resolves it)

HERE = os.path.dirname(os.path.abspath(__file__))
# MMER_B200_LIB: alternative build of the same library (kernel A/B experiments under tools/)
LIB_PATH = os.environ.get("MMER_B200_LIB") or os.path.join(HERE, "libmmer_sm100.so")

F32, BF16 = 0, 1
MAJOR_K, MAJOR_MN = 0, 1
LOSS_FOCAL, LOSS_WCE = 0, 1
REDUCE_MEAN, REDUCE_SUM, REDUCE_NONE = 0, 1, 2
DEBUG_MN_SWAP, DEBUG_FORCE_BN, DEBUG_FORCE_SIMT, DEBUG_DIRECT_STORE, DEBUG_ATT_SIMT, DEBUG_NO_PAIR, DEBUG_GENERIC_EPI, DEBUG_ATT_ROWS, DEBUG_NO_PDL, DEBUG_RESERVE_SMS = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9
DEBUG_FORCE_SPLITS, DEBUG_NO_LN_FUSE = 10, 11
DEBUG_SERVE_GLOBAL, DEBUG_SERVE_STAMPS, DEBUG_EMBED_GENERIC = 13, 14, 15
MAX_LAYERS = 16
NORM_FUSION_IDENTITY, NORM_HEAD_BATCHNORM = 1, 2   # mmer_model.norms (use_layernorm=False variants of train2.py)
G_NAMES = ["POS", "WV", "BV", "WA", "BA", "NV_W", "NV_B", "NA_W", "NA_B", "ON_W", "ON_B", "C0_W", "C0_B", "C1_W",
           "C1_B", "C4_W", "C4_B", "C5_W", "C5_B", "C8_W", "C8_B"]
L_NAMES = ["IN_W", "IN_B", "OUT_W", "OUT_B", "FF1_W", "FF1_B", "FF2_W", "FF2_B", "N1_W", "N1_B", "N2_W", "N2_B"]
G_COUNT, L_COUNT = len(G_NAMES), len(L_NAMES)
G = {n: i for i, n in enumerate(G_NAMES)}
L = {n: i for i, n in enumerate(L_NAMES)}


class MmerError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("D", C.c_void_p), ("bias", C.c_void_p),
                ("residual", C.c_void_p), ("gate", C.c_void_p),
                ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
                ("lda", C.c_int64), ("ldb", C.c_int64), ("ldd", C.c_int64),
                ("a_major", C.c_int32), ("b_major", C.c_int32), ("in_dtype", C.c_int32), ("out_dtype", C.c_int32),
                ("accumulate", C.c_int32), ("relu", C.c_int32), ("drop_p", C.c_float), ("gate_scale", C.c_float),
                ("seed", C.c_uint64), ("drop_site", C.c_uint32), ("reserved", C.c_uint32),
                ("a_rowsum", C.c_void_p), ("relu_mask_out", C.c_void_p), ("gate_bits", C.c_void_p),
                ("d_colsum", C.c_void_p)]


class Model(C.Structure):
    _fields_ = [("variant", C.c_int32), ("dtype", C.c_int32), ("B", C.c_int32), ("T", C.c_int32),
                ("video_dim", C.c_int32), ("audio_dim", C.c_int32), ("fused", C.c_int32), ("heads", C.c_int32),
                ("layers", C.c_int32), ("ffn", C.c_int32), ("hidden", C.c_int32), ("classes", C.c_int32),
                ("training", C.c_int32), ("has_mask", C.c_int32), ("p_fusion", C.c_float), ("p_classifier", C.c_float),
                ("seed", C.c_uint64), ("n_params", C.c_int64),
                ("off_g", C.c_int64 * G_COUNT), ("off_l", (C.c_int64 * L_COUNT) * MAX_LAYERS),
                ("params", C.c_void_p), ("shadow", C.c_void_p), ("grads", C.c_void_p), ("bn_state", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
                ("video", C.c_void_p), ("audio", C.c_void_p), ("mask", C.c_void_p),
                ("logits", C.c_void_p), ("probs", C.c_void_p), ("fused_out", C.c_void_p), ("attn_probs", C.c_void_p),
                ("dlogits", C.c_void_p), ("dvideo", C.c_void_p), ("daudio", C.c_void_p),
                ("stage", C.c_int32), ("input_grads_only", C.c_int32),
                ("fused_in", C.c_void_p), ("dfused_in", C.c_void_p), ("dfused_out", C.c_void_p),
                ("grad_events", C.POINTER(C.c_void_p)), ("n_grad_events", C.c_int32), ("bn_world", C.c_int32),
                ("bn_sync", C.c_void_p), ("bn_sync_user", C.c_void_p), ("norms", C.c_int32)]


# int (*bn_sync)(void* user, float* buf, int64_t n, void* stream): SyncBatchNorm hook of mmer_model
BN_SYNC_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


_P, _I64, _I, _F, _U64, _U32 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint64, C.c_uint32

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/mmer.h one to one
SIGNATURES = {
    "mmer_version": [],
    "mmer_last_error": [],
    "mmer_debug_set": [_I, _I],
    "mmer_feature_stats": [_P, _I64, _I64, _F, _P, _P, _P, _P],
    "mmer_normalize_rows": [_P, _P, _P, _P, _I64, _I64, _P],
    "mmer_collate_bf16": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _P],
    "mmer_collate": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I, _P],
    "mmer_eval_accumulate": [_P, _P, _P, _P, _I64, _I64, _P],
    "mmer_ig_expand": [_P, _P, _P, _P, _I64, _I64, _I, _I, _P],
    "mmer_ig_reduce": [_P, _P, _P, _P, _P, _I64, _I64, _I, _I, _P],
    "mmer_event_create": [C.POINTER(C.c_void_p)],
    "mmer_event_destroy": [_P],
    "mmer_stream_wait_event": [_P, _P],
    "mmer_debug_get": [_I],
    "mmer_launch_count": [],
    "mmer_gemm": [C.POINTER(GemmArgs), _P],
    "mmer_embed_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I, _F, _U64, _U32, _P],
    "mmer_embed_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I, _F, _U64, _U32, _P],
    "mmer_add_ln_fwd": [_P, _P, _P, _P, _P, _P, _I64, _I64, _I, _I, _F, _U32, _F, _U32, _U64, _P],
    "mmer_add_ln_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I, _I, _F, _U32, _F, _U32, _U64, _P],
    "mmer_add_ln_bwd_z": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I, _F, _U32, _U64, _P],
    "mmer_gemm_ln_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _F, _U32, _U64, _P],
    "mmer_pool_ln_fwd": [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I, _P],
    "mmer_pool_ln_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I, _P],
    "mmer_colsum": [_P, _P, _I64, _I64, _I64, _I, _P],
    "mmer_mha_fwd": [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I, _F, _U64, _U32, _P],
    "mmer_mha_bwd": [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I, _F, _U64, _U32, _P],
    "mmer_head_out_fwd": [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I, _P],
    "mmer_head_out_bwd": [_P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I, _P],
    "mmer_loss_fwd_bwd": [_P, _P, _P, _I, _F, _I, _P, _P, _P, _P, _I64, _I64, _F, _P],
    "mmer_adam_step": [_P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _I64, _F, _P, _F, _P],
    "mmer_adam_step_multicast": [_P, _P, _P, _P, _P, _P, _I64, _I64, _F, _F, _F, _F, _F, _I64, _F, _P, _I, _F, _P],
    "mmer_grad_sumsq_multicast": [_P, _I64, _I64, _P, _P, _I, _P],
    "mmer_grad_sumsq": [_P, _I64, _P, _P],
    "mmer_cast_bf16": [_P, _P, _I64, _P],
    "mmer_cast_f32": [_P, _P, _I64, _P],
    "mmer_bn_fwd": [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I, _I, _I, _F, _F, _U64, _U32, _P],
    "mmer_bn_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I, _I, _I, _F, _U64, _U32, _P],
    "mmer_serve_scratch_bytes": [],
    "mmer_serve_forward": [C.POINTER(Model), _P, _P, _P],
    "mmer_serve_pack": [C.POINTER(Model), _P, _P],
    "mmer_workspace_bytes": [C.POINTER(Model)],
    "mmer_model_forward": [C.POINTER(Model), _P],
    "mmer_model_backward": [C.POINTER(Model), _P],
}
_RESTYPES = {"mmer_last_error": C.c_char_p, "mmer_workspace_bytes": C.c_int64, "mmer_launch_count": C.c_int64,
             "mmer_serve_scratch_bytes": C.c_int64}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (building is the job of __graft_entry__.build / build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MmerError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    # A/B timing knobs without code changes: MMER_DEBUG="8=1,6=1" sets mmer_debug_set(8, 1) and (6, 1) at load
    for item in filter(None, os.environ.get("MMER_DEBUG", "").split(",")):
        k, v = item.split("=")
        lib.mmer_debug_set(int(k), int(v))
    _lib = lib
    return lib


_bound_device: Optional[int] = None


def bind_device(index: Optional[int]) -> None:
    """One GPU per process (the torchrun model): the library caches per-kernel attributes (opt-in shared-memory sizes,
    occupancy) that CUDA keeps per DEVICE, so the first device a process uses stays its only one.  A tensor on another
    GPU raises instead of failing later inside a launch."""
    global _bound_device
    if index is None:
        return
    if _bound_device is None:
        _bound_device = index
    elif _bound_device != index:
        raise MmerError(f"mmer_b200 is bound to cuda:{_bound_device} in this process (one GPU per process); got a tensor "
                        f"on cuda:{index}")


def last_error() -> str:
    return (load().mmer_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise MmerError(f"{what} failed (code {rc}): {last_error()}")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)

"""ctypes binding of ``libb200clip.so`` (the C ABI declared in ``include/b200clip.h``).

There is no fallback: if the shared library has not been built, or no sm_100 device is
present, the calls raise ``RuntimeError``.
"""
from __future__ import annotations

import os

import ctypes as C
import threading
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libb200clip.so"

# enums of include/b200clip.h
MAJOR_K, MAJOR_MN = 0, 1
EPI_NONE, EPI_QUICKGELU, EPI_RESIDUAL, EPI_QUICKGELU_BWD, EPI_QUICKGELU_D8, EPI_QUICKGELU_BWD_D8 = 0, 1, 2, 3, 4, 5
DT_BF16, DT_F32, DT_U8 = 0, 1, 2
ABI_VERSION = 1

_p, _i, _l, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes (every function returns int unless listed in _RESTYPES)
SIGNATURES = {
    "b200clip_abi_version": [],
    "b200clip_last_error": [],
    "b200clip_launch_count": [],
    "b200clip_ctx_create": [C.POINTER(_p), _i],
    "b200clip_ctx_destroy": [_p],
    "b200clip_gemm_bf16": [_p, _p, _l, _i, _p, _l, _i, _p, _l, _i, _p, _p, _l, _p, _p, _p, _l, _l, _l, _i, _i, _i, _p],
    "b200clip_layernorm_fwd": [_p, _p, _l, _p, _p, _p, _l, _p, _p, _p, _l, _p, _p, _p, _l, _l, _f, _i, _i, _p],
    "b200clip_layernorm_bwd": [_p, _p, _l, _p, _l, _p, _p, _p, _p, _p, _l, _p, _l, _p, _p, _p, _l, _l, _i, _p],
    "b200clip_attn_fwd": [_p, _p, _p, _p, _l, _l, _l, _i, _p],
    "b200clip_attn_bwd": [_p, _p, _p, _p, _p, _p, _p, _l, _l, _l, _l, _i, _p],
    "b200clip_attn_fwd_varlen": [_p, _p, _p, _p, _p, _l, _l, _l, _l, _i, _p],
    "b200clip_attn_bwd_varlen": [_p, _p, _p, _p, _p, _p, _p, _l, _l, _l, _l, _i, _p],
    "b200clip_embed_tokens_fwd": [_p, _p, _p, _p, _p, _i, _p, _l, _l, _l, _l, _p],
    "b200clip_embed_tokens_bwd": [_p, _p, _p, _p, _p, _l, _l, _l, _l, _p],
    "b200clip_text_pack_plan": [_p, _p, _p, _p, _l, _l, _l, _p],
    "b200clip_embed_tokens_packed_fwd": [_p, _p, _p, _p, _p, _p, _i, _l, _l, _l, _l, _l, _p],
    "b200clip_embed_tokens_packed_bwd": [_p, _p, _p, _p, _p, _p, _l, _l, _l, _l, _p],
    "b200clip_gather_rows": [_p, _p, _l, _p, _p, _l, _l, _p],
    "b200clip_scatter_rows": [_p, _p, _p, _p, _l, _l, _l, _i, _p],
    "b200clip_im2col_patch": [_p, _p, _i, _p, _l, _l, _l, _l, _p],
    "b200clip_colsum": [_p, _p, _l, _p, _l, _l, _p],
    "b200clip_vision_assemble_bwd": [_p, _p, _p, _p, _p, _l, _l, _l, _p],
    "b200clip_l2norm_fwd": [_p, _p, _p, _p, _l, _l, _p],
    "b200clip_l2norm_bwd": [_p, _p, _p, _p, _p, _l, _l, _p],
    "b200clip_cast_f32_to_bf16": [_p, _p, _p, _l, _p],
    "b200clip_split_f32_to_bf16": [_p, _p, _p, _p, _l, _p],
    "b200clip_cast_bf16_to_f32": [_p, _p, _p, _l, _p],
    "b200clip_logits": [_p, _p, _p, _p, _p, _l, _l, _l, _p],
    "b200clip_clip_loss_workspace_bytes": [_p, _l, _l, _l],
    "b200clip_clip_loss_fwd": [_p, _p, _p, _p, _l, _l, _l, _l, _p, _p, _p, _p, _p, _l, _p],
    "b200clip_clip_loss_bwd": [_p, _p, _p, _p, _p, _p, _p, _l, _l, _l, _l, _p, _p, _p, _p, _l, _p],
    "b200clip_check_gemm_f32": [_p, _p, _l, _p, _l, _i, _p, _p, _l, _p, _l, _l, _l, _l, _i, _p],
    "b200clip_check_attn_fwd_f32": [_p, _p, _p, _l, _l, _l, _i, _p],
    "b200clip_check_im2col_f32": [_p, _p, _p, _l, _l, _l, _l, _p],
    "b200clip_adamw": [_p, _p, _p, _p, _p, _p, _l, _f, _f, _f, _f, _f, _f, _l, _p, _p],
    "b200clip_adamw_g16": [_p, _p, _p, _p, _p, _p, _l, _f, _f, _f, _f, _f, _f, _l, _p, _p],
    "b200clip_resize_crop_u8": [_p, _p, _l, _l, _l, _p, _p, _l, _p, _p, _l, _l, _l, _p, _p, _l, _p],
    "b200clip_peer_buffer_bytes": [_i, _l],
    "b200clip_peer_alloc": [_p, _l, C.POINTER(_p), _p],
    "b200clip_peer_open": [_p, _p, C.POINTER(_p)],
    "b200clip_peer_close": [_p, _p],
    "b200clip_peer_free": [_p, _p],
    "b200clip_peer_allgather": [_p, _p, _i, _i, _l, _i, _p, _p, _p, _p],
    "b200clip_peer_status": [_p, _p, _p],
}
_RESTYPES = {"b200clip_last_error": C.c_char_p, "b200clip_launch_count": C.c_uint64, "b200clip_clip_loss_workspace_bytes": C.c_int64,
             "b200clip_peer_buffer_bytes": C.c_int64}

_lib = None
_lock = threading.RLock()
_ctx: dict[int, int] = {}


def load() -> C.CDLL:
    """Loads the shared library (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            # B200CLIP_LIB: an alternative build of the same library (kernel A/B experiments, tools/build_variants.py)
            path = Path(os.environ["B200CLIP_LIB"]).resolve() if os.environ.get("B200CLIP_LIB") else LIB_PATH
            if not path.exists():
                raise RuntimeError(
                    f"{path} is missing: build it with `python -m construction_clip_b200.build` "
                    "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for the CLIP hot path.")
            lib = C.CDLL(str(path))
            for name, argtypes in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the symbol is not exported
                fn.argtypes = argtypes
                fn.restype = _RESTYPES.get(name, C.c_int)
            if lib.b200clip_abi_version() != ABI_VERSION:
                raise RuntimeError("libb200clip.so ABI version mismatch; rebuild")
            _lib = lib
    return _lib


def launch_count() -> int:
    return int(load().b200clip_launch_count())


def last_error() -> str:
    msg = load().b200clip_last_error()
    return msg.decode() if msg else ""


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"libb200clip {what} failed (code {rc}): {last_error()}")


def ctx(device_index: int) -> int:
    """Per-device library context (created on first use)."""
    h = _ctx.get(device_index)
    if h is None:
        with _lock:
            h = _ctx.get(device_index)
            if h is None:
                out = _p()
                check(load().b200clip_ctx_create(C.byref(out), int(device_index)), "ctx_create")
                h = out.value
                _ctx[device_index] = h
    return h

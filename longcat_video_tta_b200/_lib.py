"""ctypes binding of libb200tta.so (the C ABI declared in include/b200tta.h).

The library is loaded from the package directory (built in-tree by ``_build.py``).  If it is missing the
import fails loudly -- there is no eager/PyTorch fallback for any of these operators."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

PKG = Path(__file__).resolve().parent
# B200TTA_LIB selects another build of the SAME library (A/B runs of kernel variants, debug-instrumented builds)
LIB_PATH = Path(os.environ["B200TTA_LIB"]).resolve() if os.environ.get("B200TTA_LIB") else PKG / "libb200tta.so"

OK, EINVAL, EARCH, ECUDA = 0, -1, -2, -3

i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class GemmSeg(C.Structure):
    _fields_ = [("a", vp), ("lda", i64), ("b", vp), ("ldb", i64), ("b_hi", vp), ("k", i64),
                ("b_mn_major", i32), ("a_mn_major", i32)]


class GemmEpi(C.Structure):
    _fields_ = [("mode", i32), ("tokens_per_frame", i32), ("d", vp), ("ldd", i64), ("d2", vp), ("ldd2", i64),
                ("d3", vp), ("ldd3", i64), ("bias", vp), ("bias_is_f32", i32), ("reserved", i32),
                ("resid", vp), ("ldr", i64), ("gate", vp), ("ldg", i64), ("aux1", vp), ("ldaux1", i64),
                ("aux2", vp), ("ldaux2", i64)]


class AttnSeg(C.Structure):
    _fields_ = [("q_begin", i32), ("q_end", i32), ("kv_len", i32)]


class TensorDesc(C.Structure):
    _fields_ = [("param", vp), ("master", vp), ("grad", vp), ("exp_avg", vp), ("exp_avg_sq", vp),
                ("numel", i64), ("is_bf16", i32), ("t_rows", i32), ("t_cols", i32), ("reserved", i32)]


EPI_STORE, EPI_STORE_F32, EPI_GELU, EPI_GATE_RESID, EPI_SWIGLU, EPI_SWIGLU_BWD = 0, 1, 2, 3, 4, 5
EPI_GEGLU = 7

# name -> argtypes (restype is int32 unless noted); mirrors include/b200tta.h one to one
SIGNATURES = {
    "b200tta_version": [],
    "b200tta_selfcheck": [],
    "b200tta_gemm": [i64, i64, C.POINTER(GemmSeg), i32, C.POINTER(GemmEpi), vp],
    "b200tta_lora_linear_fwd": [vp, i64, vp, vp, vp, vp, vp, i64, i64, i64, i32, f32, C.POINTER(GemmEpi), vp],
    "b200tta_lora_linear_bwd": [vp, i64, vp, i64, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, f32,
                                C.POINTER(GemmEpi), vp],
    "b200tta_attn_fwd": [vp, i64, vp, vp, i64, vp, i64, vp, i64, i32, i32, i32, f32, C.POINTER(AttnSeg), i32, vp],
    "b200tta_attn_bwd": [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, i64, vp, i64, vp, i64, i32, i32,
                         i32, f32, C.POINTER(AttnSeg), i32, vp],
    "b200tta_attn_bwd_fused": [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, i64, vp, i64, vp, i64, i32, i32,
                               i32, f32, C.POINTER(AttnSeg), i32, vp, vp],
    "b200tta_attn_bsa_fwd": [vp, i64, vp, vp, i64, vp, i64, vp, i64, i32, i32, f32, vp, vp, vp],
    "b200tta_attn_bsa_bwd": [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, i64, vp, i64, vp, i64, i32, i32, f32,
                             vp, vp, vp, vp, vp],
    "b200tta_ln_mod_fwd": [vp, i64, vp, i64, vp, vp, i64, i32, f32, i64, i32, i32, f32, vp],
    "b200tta_ln_mod_bwd": [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, i32, f32, vp, vp, i64, i64, i32, i32, f32, vp],
    "b200tta_qk_rmsnorm_rope_fwd": [vp, i64, vp, i64, vp, vp, i32, i32, i64, i64, i32, i32, i32, f32, f32, vp],
    "b200tta_qk_rmsnorm_rope_bwd": [vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, i32, i32, i64, i64, i32, i32, i32,
                                    f32, f32, vp],
    "b200tta_gate_mul": [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, i64, i32, i32, vp],
    "b200tta_noise_patchify": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp],
    "b200tta_patchify": [vp, vp, i32, i32, i32, vp],
    "b200tta_unpatchify": [vp, vp, i32, i32, i32, vp],
    "b200tta_latent_to_tokens": [vp, vp, i32, i32, i32, i32, vp],
    "b200tta_swiglu_fwd": [vp, vp, vp, i64, vp],
    "b200tta_swiglu_bwd": [vp, vp, vp, vp, vp, i64, vp],
    "b200tta_mse_fwd_bwd": [vp, vp, vp, vp, i64, f32, vp],
    "b200tta_timestep_sinusoid": [vp, vp, i32, i32, vp],
    "b200tta_skinny_linear": [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "b200tta_skinny_linear_bwd": [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "b200tta_lora_down": [vp, i64, vp, i64, vp, i32, i64, i64, i32, f32, vp],
    "b200tta_lora_grad": [vp, vp, i64, vp, i64, i64, i64, i32, vp],
    "b200tta_mt_sumsq": [vp, i32, i64, vp, vp],
    "b200tta_clip_coef": [vp, vp, vp, i32, f32, i32, f32, vp],
    "b200tta_mt_adamw": [vp, i32, i64, vp, f32, f32, f32, f32, f32, f32, i32, i32, vp],
    "b200tta_mt_sgd": [vp, i32, i64, vp, f32, f32, f32, vp],
    "b200tta_colsum": [vp, vp, i64, i64, i32, vp],
    "b200tta_gather_rows": [vp, i64, vp, i64, vp, i64, i32, vp],
    "b200tta_t5_rmsnorm": [vp, i64, vp, i64, vp, i64, i32, f32, vp],
    "b200tta_latent_affine": [vp, vp, vp, vp, i64, i64, i32, i32, i32, vp],
    "b200tta_t5_attn": [vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, i32, i32, i32, vp],
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m longcat_video_tta_b200._build` "
                "(or __graft_entry__.build()).  This package has no CPU / eager fallback.")
        lib = C.CDLL(str(LIB_PATH))
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header / library mismatch
            fn.argtypes = argtypes
            fn.restype = i32
        lib.b200tta_launch_count.argtypes = []
        lib.b200tta_launch_count.restype = i64
        lib.b200tta_last_error.argtypes = []
        lib.b200tta_last_error.restype = C.c_char_p
        _lib = lib
    return _lib


class B200TTAError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != OK:
        msg = load().b200tta_last_error().decode(errors="replace")
        kind = {EINVAL: "EINVAL", EARCH: "EARCH", ECUDA: "ECUDA"}.get(rc, str(rc))
        raise B200TTAError(f"{what} failed ({kind}): {msg}")

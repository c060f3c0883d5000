"""ctypes binding of `libvitssl_b200.so` (C ABI declared in `include/vitssl_b200.h`).

There is no CPU fallback: if the library is missing or the device is not sm_100, calls raise.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_ROOT = os.path.dirname(os.path.dirname(_HERE))  # .../vit-ssl_b200
# VITSSL_LIB points at an alternative build of the same library (A/B timing of two builds on one box)
LIB_PATH = os.environ.get("VITSSL_LIB") or os.path.join(_PKG_ROOT, "lib", "libvitssl_b200.so")

_T = {
    "p": ctypes.c_void_p,
    "i": ctypes.c_int,
    "l": ctypes.c_int64,
    "f": ctypes.c_float,
    "u": ctypes.c_uint64,
    "s": ctypes.c_void_p,  # cudaStream_t
}

# name -> argument signature (one letter per argument, see _T). Must mirror include/vitssl_b200.h.
SIGNATURES = {}
# the gemm signature spelled out (A,B,C | M,N,K,lda,ldb,ldc | a_mn,b_mn,epilogue | bias,aux |
# ld_aux | alpha | out_fp32, split_k | dropout_p | seed, offset | stream)
SIGNATURES["vitssl_gemm_bf16"] = "ppp" + "llllll" + "iii" + "pp" + "l" + "f" + "ii" + "f" + "uu" + "s"

SIGNATURES["vitssl_gemm_bf16_rowsum"] = "pppp" + "llllll" + "ii" + "f" + "i" + "s"
SIGNATURES["vitssl_add_layernorm_fwd"] = "plpp" + "pp" + "ppp" + "ll" + "ff" + "uu" + "s"
SIGNATURES["vitssl_add_layernorm_bwd"] = "ppl" + "ppp" + "pl" + "pl" + "p" + "pp" + "ll" + "f" + "uu" + "s"
SIGNATURES["vitssl_add_layernorm_bwd_acc"] = SIGNATURES["vitssl_add_layernorm_bwd"]
SIGNATURES["vitssl_attention_supported"] = "lll"
SIGNATURES["vitssl_attention_fwd"] = "ppp" + "lll" + "ppl" + "p" + "llll" + "f" + "s"
SIGNATURES["vitssl_attention_bwd"] = "ppp" + "lll" + "pppl" + "pp" + "pl" + "pl" + "pl" + "llll" + "f" + "s"
SIGNATURES["vitssl_attention_generic_fwd"] = "ppp" + "p" + "ppp" + "lllll" + "f" + "s"
SIGNATURES["vitssl_attention_generic_bwd"] = "ppp" + "p" + "pp" + "p" + "ppp" + "lllll" + "f" + "s"
SIGNATURES["vitssl_encoder_stack_fwd"] = "ps"
SIGNATURES["vitssl_encoder_stack_bwd"] = "ps"
SIGNATURES["vitssl_multi_cast_bf16"] = "pppis"
SIGNATURES["vitssl_multi_ema"] = "pppifs"
SIGNATURES["vitssl_colsum_bf16"] = "plllps"
SIGNATURES["vitssl_colsum_bf16_acc"] = "plllps"
SIGNATURES["vitssl_im2col_bf16"] = "pp" + "lllll" + "s"
SIGNATURES["vitssl_gather_patches_f32"] = "ppp" + "lllll" + "s"
SIGNATURES["vitssl_embed_tokens_fwd"] = "pppppp" + "lll" + "s"
SIGNATURES["vitssl_embed_tokens_bwd"] = "pllp" + "ppp" + "lll" + "i" + "s"
SIGNATURES["vitssl_gather_rows_bf16"] = "plppll" + "s"
SIGNATURES["vitssl_scatter_rows_f32"] = "pppll" + "s"
SIGNATURES["vitssl_simmim_mask"] = "pppp" + "lll" + "uu" + "s"
SIGNATURES["vitssl_l1_loss_fwd"] = "ppppl" + "s"
SIGNATURES["vitssl_l2norm_fwd"] = "pppll" + "s"
SIGNATURES["vitssl_l2norm_bwd"] = "ppppll" + "s"
SIGNATURES["vitssl_weight_norm_fwd"] = "ppppll" + "s"
SIGNATURES["vitssl_weight_norm_bwd"] = "pppppp" + "ll" + "s"
SIGNATURES["vitssl_center_ema"] = "ppplff" + "s"
SIGNATURES["vitssl_dino_loss_fwd"] = "pppppp" + "llll" + "ff" + "s"
SIGNATURES["vitssl_dino_loss_bwd"] = "ppppppp" + "llll" + "ff" + "s"

SIGNATURES["vitssl_im2col_u8_bf16"] = SIGNATURES["vitssl_im2col_bf16"]
SIGNATURES["vitssl_gather_patches_u8_f32"] = SIGNATURES["vitssl_gather_patches_f32"]
SIGNATURES["vitssl_interp_rows_fwd"] = "pppp" + "lll" + "s"
SIGNATURES["vitssl_interp_rows_bwd"] = "pppp" + "llll" + "s"
SIGNATURES["vitssl_l1_loss_bwd"] = "pppl" + "s"
SIGNATURES["vitssl_adamw_step"] = "ppppppp" + "i" + "fffff" + "pp" + "s"
SIGNATURES["vitssl_profile_read"] = "lplpp"
SIGNATURES["vitssl_graph_stats"] = "pp"
SIGNATURES["vitssl_graph_enable"] = "i"
SIGNATURES["vitssl_multi_ema_shadow"] = "ppppifs"
SIGNATURES["vitssl_mean_tokens_f32"] = "pp" + "lll" + "s"
SIGNATURES["vitssl_knn_cosine"] = "pppppppp" + "lllll" + "s"

_lib = None


class VitsslError(RuntimeError):
    pass


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"vitssl_b200: native library not found at {LIB_PATH}; build it with "
            f"`python {os.path.join(_PKG_ROOT, 'build.py')}` (there is no CPU fallback)."
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.vitssl_version.restype = ctypes.c_int
    lib.vitssl_last_error.restype = ctypes.c_char_p
    lib.vitssl_device_check.restype = ctypes.c_int
    lib.vitssl_num_sms.restype = ctypes.c_int
    lib.vitssl_launch_count.restype = ctypes.c_int64
    lib.vitssl_launch_count.argtypes = [ctypes.c_int]
    lib.vitssl_profile_begin.restype = ctypes.c_int
    lib.vitssl_profile_end.restype = ctypes.c_int64
    lib.vitssl_randperm_bits.restype = ctypes.c_int
    lib.vitssl_randperm_bits.argtypes = [ctypes.c_int64]
    lib.vitssl_randperm_offset_per_call.restype = ctypes.c_int64
    lib.vitssl_randperm_offset_per_call.argtypes = [ctypes.c_int64]
    for name, sig in SIGNATURES.items():
        if os.environ.get("VITSSL_LIB") and not hasattr(lib, name):
            continue  # an older build given for A/B timing may predate an entry point
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = [_T[c] for c in sig]
    _lib = lib
    return lib


def lib():
    return _load()


_device_ok = False


def ensure_device():
    """Raise unless the current CUDA device can run the sm_100a kernels."""
    global _device_ok
    if _device_ok:
        return
    if not torch.cuda.is_available():
        raise VitsslError("vitssl_b200 needs a CUDA device (sm_100a / B200); no CPU fallback exists")
    l = _load()
    rc = l.vitssl_device_check()
    if rc != 0:
        raise VitsslError(l.vitssl_last_error().decode())
    _device_ok = True


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def call(name: str, *args):
    l = _load()
    rc = getattr(l, name)(*args)
    if rc != 0:
        raise VitsslError(f"{name} failed ({rc}): {l.vitssl_last_error().decode()}")


def profile_begin() -> None:
    """Open a per-launch profile of the C-sequenced paths (csrc/api.cu, ProfScope)."""
    _load().vitssl_profile_begin()


def profile_collect():
    """Close the profile; returns [(kind, ms, algorithmic work)] — waits for the device."""
    l = _load()
    n = int(l.vitssl_profile_end())
    out = []
    kind = ctypes.create_string_buffer(64)
    work, ms = ctypes.c_double(), ctypes.c_float()
    for i in range(n):
        call("vitssl_profile_read", i, ctypes.addressof(kind), 64, ctypes.addressof(work), ctypes.addressof(ms))
        out.append((kind.value.decode(), float(ms.value), float(work.value)))
    return out


def graph_stats():
    """(graphs captured, stack calls served by a graph replay) since the library was loaded (csrc/encoder.cu)."""
    c, r = ctypes.c_int64(), ctypes.c_int64()
    call("vitssl_graph_stats", ctypes.addressof(c), ctypes.addressof(r))
    return int(c.value), int(r.value)


def graph_enable(on: bool) -> None:
    call("vitssl_graph_enable", 1 if on else 0)


def launch_count(reset: bool = False) -> int:
    return int(_load().vitssl_launch_count(1 if reset else 0))

"""ctypes binding of libovk.so (the C ABI declared in include/ovk.h).

There is no CPU or PyTorch fallback: if the shared library is missing or the device is not sm_100 the
product path raises.  Building is `make -C openvision_b200/csrc` (see __graft_entry__.build()).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libovk.so")

# flags (mirror include/ovk.h)
EPI_NONE = 0
EPI_GELU_ERF = 1
EPI_GELU_TANH = 2
EPI_GELU_QUICK = 3
EPI_BIAS = 4
EPI_RESIDUAL = 8
EPI_SAVE_PREACT = 16


class OvkError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); every symbol include/ovk.h declares must be listed here (tests check both ways)
_PROTOTYPES = {
    "ovk_version": (c_int, []),
    "ovk_last_error": (c_char_p, []),
    "ovk_device_supported": (c_int, []),
    "ovk_gemm_bf16": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    "ovk_gemm_bf16_ex": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_void_p]),
    "ovk_gemm_bf16_ln": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_int, c_float, c_void_p, c_longlong, c_void_p, c_longlong,
                                 c_void_p, c_int, c_void_p]),
    "ovk_sumsq": (c_int, [c_void_p, c_int, c_longlong, c_void_p, c_void_p]),
    "ovk_adamw_step": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_longlong, c_float, c_float, c_float,
                               c_float, c_float, c_int, c_float, c_void_p, c_float, c_void_p]),
    "ovk_gemm_bf16_rowadd": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int,
                                     c_void_p, c_int, c_void_p]),
    "ovk_pack_ln_linear": (c_int, [c_void_p, c_int, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong,
                                   c_void_p, c_int, c_int, c_void_p]),
    "ovk_row_stats": (c_int, [c_void_p, c_longlong, c_void_p, c_int, c_int, c_void_p]),
    "ovk_gemm_bf16_nn": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int,
                                 c_int, c_float, c_void_p, c_longlong, c_int, c_void_p]),
    "ovk_gemm_colsum_rows": (c_longlong, [c_int]),
    "ovk_gemm_bf16_nn_dact_colsum": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int,
                                             c_float, c_void_p, c_longlong, c_int, c_void_p, c_void_p]),
    "ovk_colsum_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ovk_gemm_bf16_tn": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int,
                                 c_int, c_float, c_void_p]),
    "ovk_gemm_bf16_scaled": (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_longlong, c_int, c_void_p, c_longlong, c_int,
                                     c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "ovk_add_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "ovk_act_fwd": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    "ovk_act_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    "ovk_layernorm_fwd": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int, c_int, c_float, c_void_p]),
    "ovk_layernorm_bwd": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_longlong, c_void_p, c_longlong, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "ovk_im2col_patches": (c_int, [c_void_p, c_int, c_void_p, c_longlong, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ovk_patch_embed_kdim": (c_int, [c_int]),
    "ovk_patch_embed_supported": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "ovk_patch_embed": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ovk_pool_head": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_float,
                              c_void_p, c_int, c_void_p]),
    "ovk_embed_assemble": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ovk_col2im_patches": (c_int, [c_void_p, c_longlong, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ovk_colsum_bf16": (c_int, [c_void_p, c_longlong, c_int, c_int, c_void_p, c_void_p]),
    "ovk_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                  c_float, c_void_p]),
    "ovk_attention_fwd_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "ovk_attention_bwd_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                     c_int, c_float, c_int, c_void_p]),
    "ovk_attention_bwd_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_int, c_float, c_int, c_void_p]),
    "ovk_attention_bwd_fused_workspace_floats": (c_longlong, [c_int, c_int, c_int, c_int, c_int]),
    "ovk_attention_bwd_workspace_floats": (c_longlong, [c_int, c_int, c_int, c_int]),
    "ovk_pool_tokens_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ovk_l2_normalize_bwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_float, c_void_p]),
    "ovk_attention_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "ovk_pool_tokens": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ovk_clip_loss_workspace_floats": (c_longlong, [c_int, c_int]),
    "ovk_clip_loss_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "ovk_clip_loss_finalize": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ovk_clip_loss_combine": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ovk_clip_loss_value": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ovk_clip_loss_grad_logits": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                          c_float, c_float, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p]),
    "ovk_l2_normalize": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_float, c_void_p]),
}


def load(path: str | None = None):
    """Load libovk.so and attach prototypes. Raises OvkError when the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise OvkError(
            f"{p} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = ctypes.CDLL(p)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_PROTOTYPES)


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().ovk_last_error()
        raise OvkError(f"{what or 'libovk'} failed (code {rc}): {msg.decode() if msg else ''}")


def call(name: str, *args):
    lib = load()
    check(getattr(lib, name)(*args), name)

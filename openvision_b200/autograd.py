"""Functional layer between the drop-in modules and libovk: packs parameters, sequences kernels, and (when gradients
are requested) records torch.autograd.Function nodes whose backward is again libovk kernels."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import OvkError


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _no_backward(what: str):
    raise OvkError(f"{what}: backward kernels are not part of this build yet; call under torch.no_grad()")


def _w_bf16(owner, key, w, transpose=False):
    from .transformer import _packed
    return _packed(owner, key, w, torch.bfloat16, (lambda t: t.t()) if transpose else None)


def _v_f32(owner, key, v):
    from .transformer import _packed
    return _packed(owner, key, v, torch.float32)


def layer_norm_fn(x2, weight, bias, eps, owner):
    if _needs_grad(x2, weight, bias):
        _no_backward("layer_norm")
    return ops.layernorm(x2, _v_f32(owner, "ln_w", weight), _v_f32(owner, "ln_b", bias), eps)


def linear_fn(x2, weight, bias, residual, act, owner, transpose_weight=False):
    """x2 @ W^T (+bias)(act)(+residual); `transpose_weight` for raw [in, out] projection matrices (visual.proj)."""
    if _needs_grad(x2, weight, bias, residual):
        _no_backward("linear")
    w = _w_bf16(owner, "w_t" if transpose_weight else "w", weight, transpose=transpose_weight)
    b = _v_f32(owner, "b", bias) if bias is not None else None
    return ops.gemm(x2, w, bias=b, residual=residual, act=act)


def patch_embed_fn(images, conv_weight, conv):
    if _needs_grad(images, conv_weight):
        _no_backward("patch_embed")
    w, kpad = conv.packed_weight()
    cols = ops.im2col_patches(images, conv.kernel_size[0], kpad)
    return ops.gemm(cols, w)


def embed_assemble_fn(tok, cls, pos, B, N, owner):
    if _needs_grad(tok, cls, pos):
        _no_backward("embed_assemble")
    return ops.embed_assemble(tok, _v_f32(owner, "cls", cls), _v_f32(owner, "pos", pos), B, N).view(B * (N + 1), -1)


def pool_fn(x2, B, L, pool_type):
    if _needs_grad(x2):
        _no_backward("pool")
    return ops.pool_tokens(x2.view(B, L, -1), pool_type)


def attention_block_fn(x2, attn, B, L, residual: Optional[torch.Tensor], out: Optional[torch.Tensor] = None):
    """in_proj GEMM (+bias) -> flash attention -> out_proj GEMM (+bias, +residual)."""
    if _needs_grad(x2, attn.in_proj_weight, attn.out_proj.weight, residual):
        _no_backward("attention")
    H = attn.num_heads
    hd = attn.embed_dim // H
    wqkv = _w_bf16(attn, "in_w", attn.in_proj_weight)
    bqkv = _v_f32(attn, "in_b", attn.in_proj_bias) if attn.in_proj_bias is not None else None
    wo = _w_bf16(attn, "out_w", attn.out_proj.weight)
    bo = _v_f32(attn, "out_b", attn.out_proj.bias) if attn.out_proj.bias is not None else None
    qkv = ops.gemm(x2, wqkv, bias=bqkv)
    a = ops.attention(qkv, B, L, H, hd)
    return ops.gemm(a, wo, bias=bo, residual=residual, out=out)


def block_fn(x2, blk, B, L, inplace):
    """One ResidualAttentionBlock on a bf16 [B*L, D] residual stream: 7 kernels
    (LN, QKV GEMM, attention, out-proj GEMM+residual, LN, fc1 GEMM+GELU, fc2 GEMM+residual)."""
    params = [p for p in blk.parameters()]
    if _needs_grad(x2, *params):
        _no_backward("ResidualAttentionBlock")
    h = ops.layernorm(x2, _v_f32(blk.ln_1, "ln_w", blk.ln_1.weight), _v_f32(blk.ln_1, "ln_b", blk.ln_1.bias), blk.ln_1.eps)
    x_mid = attention_block_fn(h, blk.attn, B, L, residual=x2, out=x2 if inplace else None)
    h = ops.layernorm(x_mid, _v_f32(blk.ln_2, "ln_w", blk.ln_2.weight), _v_f32(blk.ln_2, "ln_b", blk.ln_2.bias), blk.ln_2.eps, out=h)
    fc, pj = blk.mlp.c_fc, blk.mlp.c_proj
    f = ops.gemm(h, _w_bf16(fc, "w", fc.weight), bias=_v_f32(fc, "b", fc.bias) if fc.bias is not None else None,
                 act=blk._act_kind())
    return ops.gemm(f, _w_bf16(pj, "w", pj.weight), bias=_v_f32(pj, "b", pj.bias) if pj.bias is not None else None,
                    residual=x_mid, out=x_mid)


def normalize_fn(features: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """F.normalize(features, dim=-1): bf16 rows in, unit rows out in the caller's dtype (fp32 stays fp32)."""
    if _needs_grad(features):
        _no_backward("l2_normalize")
    if not features.is_cuda:
        raise OvkError("openvision_b200 runs on CUDA (sm_100a) only; got a CPU tensor")
    out_dtype = torch.float32 if features.dtype == torch.float32 else torch.bfloat16
    x = features if features.dtype == torch.bfloat16 else features.to(torch.bfloat16)
    y = ops.l2_normalize(x.contiguous(), out_dtype=out_dtype, eps=eps)
    return y if y.dtype == features.dtype else y.to(features.dtype)

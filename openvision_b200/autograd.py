"""Functional layer between the drop-in modules and libovk: packs parameters, sequences kernels, and records
torch.autograd.Function nodes whose backward is again libovk kernels (torch is the tape, never the arithmetic).

Backward of one ResidualAttentionBlock (transformer.py:254-265), given dY (bf16 [B*L, D]):
    dU   = (dY W2) . act'(u)          gemm_nn + fused GELU'          dW2 = dY^T f   gemm_tn    db2 = colsum(dY)
    dH2  = dU W1                      gemm_nn                        dW1 = dU^T h2  gemm_tn    db1 = colsum(dU)
    dXm  = LN2'(dH2) + dY             layernorm_bwd (+ residual grad)
    dA   = dXm Wo                     gemm_nn                        dWo = dXm^T a  gemm_tn    dbo = colsum(dXm)
    dQKV = attention_bwd(qkv, a, dA, lse)
    dH1  = dQKV Wqkv                  gemm_nn                        dWqkv = dQKV^T h1         dbqkv = colsum(dQKV)
    dX   = LN1'(dH1) + dXm            layernorm_bwd (+ residual grad)
h1 / h2 (the LayerNorm outputs) are recomputed from the saved statistics rather than stored.
"""
from __future__ import annotations

import math
from typing import Optional

import os

import torch

from . import ops
from ._lib import OvkError


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors)


def _w_bf16(owner, key, w, transpose=False):
    from .transformer import _packed
    return _packed(owner, key, w, torch.bfloat16, (lambda t: t.t()) if transpose else None)


def _v_f32(owner, key, v):
    from .transformer import _packed
    return _packed(owner, key, v, torch.float32)


def _grad_dtype(p: Optional[torch.Tensor]):
    return torch.float32 if p is None or p.dtype == torch.float32 else torch.bfloat16


def _like_param(g: Optional[torch.Tensor], p: Optional[torch.Tensor]):
    if g is None or p is None:
        return None
    g = g.reshape(p.shape)
    return g if g.dtype == p.dtype else g.to(p.dtype)


def _c(t: torch.Tensor) -> torch.Tensor:
    """upstream gradients arrive in whatever layout autograd built; the kernels want contiguous bf16 rows"""
    if t.dtype != torch.bfloat16:
        t = t.to(torch.bfloat16)
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------------------------------------
# LayerNorm
# ----------------------------------------------------------------------------------------------------------------
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, weight, bias, eps, owner):
        g, b = _v_f32(owner, "ln_w", weight), _v_f32(owner, "ln_b", bias)
        if any(ctx.needs_input_grad):
            y, mean, rstd = ops.layernorm(x2, g, b, eps, save_stats=True)
            ctx.save_for_backward(x2, g, mean, rstd)
            ctx.params = (weight, bias)
            return y
        return ops.layernorm(x2, g, b, eps)

    @staticmethod
    def backward(ctx, dy):
        x2, g, mean, rstd = ctx.saved_tensors
        weight, bias = ctx.params
        dg = torch.zeros_like(g)
        db = torch.zeros_like(g)
        dx = ops.layernorm_bwd(_c(dy), x2, g, mean, rstd, dg, db)
        return dx, _like_param(dg, weight), _like_param(db, bias), None, None


def layer_norm_fn(x2, weight, bias, eps, owner):
    if not _needs_grad(x2, weight, bias):
        return ops.layernorm(x2, _v_f32(owner, "ln_w", weight), _v_f32(owner, "ln_b", bias), eps)
    return _LayerNorm.apply(x2, weight, bias, eps, owner)


# ----------------------------------------------------------------------------------------------------------------
# Linear (+ bias, + residual) and raw projection matrices
# ----------------------------------------------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, weight, bias, residual, owner, transpose_weight):
        w = _w_bf16(owner, "w_t" if transpose_weight else "w", weight, transpose=transpose_weight)   # [N, K]
        b = _v_f32(owner, "b", bias) if bias is not None else None
        y = ops.gemm(x2, w, bias=b, residual=residual)
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(x2, w)
            ctx.meta = (weight, bias, residual is not None, transpose_weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        weight, bias, has_res, transposed = ctx.meta
        dy = _c(dy)
        dx = ops.gemm_nn(dy, w) if ctx.needs_input_grad[0] else None                      # dX = dY W
        dw = None
        if ctx.needs_input_grad[1]:
            if transposed:   # parameter stored [in, out] (visual.proj, text_projection): d(proj) = X^T dY
                dw = ops.gemm_tn(x2, dy, out_dtype=_grad_dtype(weight))
            else:            # nn.Linear layout [out, in]: dW = dY^T X
                dw = ops.gemm_tn(dy, x2, out_dtype=_grad_dtype(weight))
        db = ops.colsum(dy) if bias is not None and ctx.needs_input_grad[2] else None
        return dx, _like_param(dw, weight), _like_param(db, bias), (dy if has_res else None), None, None


def linear_fn(x2, weight, bias, residual, act, owner, transpose_weight=False):
    """x2 @ W^T (+bias)(+residual); `transpose_weight` for raw [in, out] projection matrices (visual.proj).
    Activations are fused only on the block path (block_fn) and are not accepted here."""
    if act is not None:
        raise OvkError("linear_fn: fused activations belong to the block path")
    N = weight.shape[1] if transpose_weight else weight.shape[0]
    K = weight.shape[0] if transpose_weight else weight.shape[1]
    if (N % 8) or (K % 8):
        raise OvkError(f"linear: in/out features must be multiples of 8 (got {K} -> {N})")
    if not _needs_grad(x2, weight, bias, residual):
        w = _w_bf16(owner, "w_t" if transpose_weight else "w", weight, transpose=transpose_weight)
        return ops.gemm(x2, w, bias=_v_f32(owner, "b", bias) if bias is not None else None, residual=residual)
    return _Linear.apply(x2, weight, bias, residual, owner, transpose_weight)


# ----------------------------------------------------------------------------------------------------------------
# activation as a module (hooked nn.GELU / QuickGELU)
# ----------------------------------------------------------------------------------------------------------------
class _Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kind):
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(x)
            ctx.kind = kind
        return ops.act_fwd(x, kind)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.act_bwd(x, _c(dy).reshape(x.shape), ctx.kind), None


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return ops.add(a, b)

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


def add_fn(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a + b on add_kernel (the stand-alone residual add of the hooked, module-by-module block path)."""
    if a.numel() % 8:
        raise OvkError("add: element count must be a multiple of 8")
    a, b = a.contiguous(), b.contiguous()
    return _Add.apply(a, b) if _needs_grad(a, b) else ops.add(a, b)


def act_fn(x: torch.Tensor, kind: str) -> torch.Tensor:
    return _Act.apply(x, kind) if _needs_grad(x) else ops.act_fwd(x, kind)


# ----------------------------------------------------------------------------------------------------------------
# patch embedding, cls / pos assembly
# ----------------------------------------------------------------------------------------------------------------
def _pos_table(owner, cls, pos):
    """bf16 [L, D] table the patch-embedding epilogue adds: positional_embedding with class_embedding folded into row 0
    (transformer.py:615-617); cached on the versions of the two parameters."""
    cache = owner.__dict__.setdefault("_ovk_cache", {})
    tag = tuple((t.data_ptr(), t._version, t.dtype, t.device) for t in (pos, cls))
    hit = cache.get("pos_cls")
    if hit is None or hit[0] != tag:
        with torch.no_grad():
            table = pos.detach().float().clone()
            table[0] += cls.detach().float()
            hit = (tag, table.to(torch.bfloat16).contiguous())
        cache["pos_cls"] = hit
    return hit[1]


def _patch_embed_backward(ctx, dtok):
    """shared by the two patch-embedding nodes: dtok bf16 [B*(N+1), D] (class-token rows included; the im2col matrix has a
    zero row there, so they contribute nothing to dW).  The im2col matrix is rebuilt here (the forward never made one)."""
    (images,) = ctx.saved_tensors
    conv, conv_weight, P = ctx.conv, ctx.conv_weight, ctx.P
    w, kpad = conv.packed_weight()
    dw = dimg = None
    if ctx.needs_input_grad[1]:
        k = conv_weight[0].numel()
        cols = ops.im2col_patches(images, P, kpad, lead_rows=1)
        dwp = ops.gemm_tn(dtok, cols, out_dtype=_grad_dtype(conv_weight))          # [D, Kpad]
        dw = _like_param(dwp[:, :k].contiguous(), conv_weight)
        del cols
    if ctx.needs_input_grad[0]:
        dcols = ops.gemm_nn(dtok, w)                                                # [B*(N+1), Kpad]
        B, _, H, W = images.shape
        dimg = ops.col2im_patches(dcols, B, H, W, P, 1, torch.float32 if images.dtype == torch.float32 else torch.bfloat16)
    return dimg, dw


class _PatchEmbed(torch.autograd.Function):
    """conv1 alone: tokens with a zero row in every image's class-token slot."""

    @staticmethod
    def forward(ctx, images, conv_weight, conv):
        P = conv.kernel_size[0]
        if ops.patch_embed_supported(images, P, conv.out_channels):
            tok = ops.patch_embed(images, conv.packed_weight_fused(), P, None)
            tok = tok.view(-1, tok.shape[-1])
        else:
            w, kpad = conv.packed_weight()
            tok = ops.gemm(ops.im2col_patches(images, P, kpad, lead_rows=1), w)
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(images)
            ctx.conv, ctx.conv_weight, ctx.P = conv, conv_weight, P
        return tok

    @staticmethod
    def backward(ctx, dtok):
        dimg, dw = _patch_embed_backward(ctx, _c(dtok))
        return dimg, dw, None


def patch_embed_fn(images, conv_weight, conv):
    """images [B,3,H,W] -> bf16 [B*(N+1), D] patch tokens with a zero row in every image's cls slot."""
    if not _needs_grad(images, conv_weight):
        P = conv.kernel_size[0]
        if ops.patch_embed_supported(images, P, conv.out_channels):
            tok = ops.patch_embed(images, conv.packed_weight_fused(), P, None)
            return tok.view(-1, tok.shape[-1])
        w, kpad = conv.packed_weight()
        return ops.gemm(ops.im2col_patches(images, P, kpad, lead_rows=1), w)
    return _PatchEmbed.apply(images, conv_weight, conv)


class _PatchEmbedTokens(torch.autograd.Function):
    """transformer.py:610-617 as one kernel: conv1, reshape / permute, cat(class token), + positional embedding."""

    @staticmethod
    def forward(ctx, images, conv_weight, cls, pos, conv, owner):
        P = conv.kernel_size[0]
        tok = ops.patch_embed(images, conv.packed_weight_fused(), P, _pos_table(owner, cls, pos))
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(images)
            ctx.conv, ctx.conv_weight, ctx.P = conv, conv_weight, P
            ctx.cls, ctx.pos = cls, pos
        return tok.view(-1, tok.shape[-1])

    @staticmethod
    def backward(ctx, dx):
        dx = _c(dx)
        dimg, dw = _patch_embed_backward(ctx, dx)
        dcls = dpos = None
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            L, D = ctx.pos.shape
            s = ops.colsum(dx.view(-1, L * D)).view(L, D)       # sum over images
            dpos = _like_param(s, ctx.pos) if ctx.needs_input_grad[3] else None
            dcls = _like_param(s[0].clone(), ctx.cls) if ctx.needs_input_grad[2] else None
        return dimg, dw, dcls, dpos, None, None


def patch_embed_tokens_fn(images, conv, cls, pos, owner):
    """-> bf16 [B*(N+1), D] finished token buffer (class token and positional embedding added in the GEMM epilogue)."""
    if not _needs_grad(images, conv.weight, cls, pos):
        tok = ops.patch_embed(images, conv.packed_weight_fused(), conv.kernel_size[0], _pos_table(owner, cls, pos))
        return tok.view(-1, tok.shape[-1])
    return _PatchEmbedTokens.apply(images, conv.weight, cls, pos, conv, owner)


class _EmbedAssemble(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tok, cls, pos, B, N, owner):
        x = ops.embed_assemble(tok, _v_f32(owner, "cls", cls), _v_f32(owner, "pos", pos), B, N, inplace=False)
        ctx.meta = (cls, pos, B, N, tok.shape[0] == B * (N + 1))
        return x.view(B * (N + 1), -1)

    @staticmethod
    def backward(ctx, dx):
        cls, pos, B, N, lead = ctx.meta
        dx = _c(dx)
        D = dx.shape[-1]
        dpos = dcls = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            s = ops.colsum(dx.view(B, (N + 1) * D)).view(N + 1, D)      # sum over images
            dpos = _like_param(s, pos) if ctx.needs_input_grad[2] else None
            dcls = _like_param(s[0].clone(), cls) if ctx.needs_input_grad[1] else None
        if not lead:      # tokens came without the cls slot ([B*N, D]): their gradient is rows 1.. of every image
            dx = dx.view(B, N + 1, D)[:, 1:].reshape(B * N, D)
        return dx, dcls, dpos, None, None, None


def embed_assemble_fn(tok, cls, pos, B, N, owner):
    if not _needs_grad(tok, cls, pos):
        lead = tok.shape[0] == B * (N + 1)      # GEMM output with the cls slot in place: assemble in place
        x = ops.embed_assemble(tok, _v_f32(owner, "cls", cls), _v_f32(owner, "pos", pos), B, N, inplace=lead)
        return x.view(B * (N + 1), -1)
    return _EmbedAssemble.apply(tok, cls, pos, B, N, owner)


# ----------------------------------------------------------------------------------------------------------------
# pooling, L2 normalisation
# ----------------------------------------------------------------------------------------------------------------
class _Pool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, B, L, pool_type):
        ctx.meta = (B, L, pool_type)
        return ops.pool_tokens(x2.view(B, L, -1), pool_type)

    @staticmethod
    def backward(ctx, dp):
        B, L, pool_type = ctx.meta
        return ops.pool_tokens_bwd(_c(dp), B, L, pool_type).view(B * L, -1), None, None, None


def pool_fn(x2, B, L, pool_type):
    if not _needs_grad(x2):
        return ops.pool_tokens(x2.view(B, L, -1), pool_type)
    return _Pool.apply(x2, B, L, pool_type)


class _Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, out_dtype, eps):
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(x)
            ctx.eps = eps
        return ops.l2_normalize(x, out_dtype=out_dtype, eps=eps)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        if dy.dtype not in (torch.float32, torch.bfloat16):
            dy = dy.float()
        return ops.l2_normalize_bwd(x, dy.contiguous(), ctx.eps), None, None


def normalize_fn(features: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """F.normalize(features, dim=-1): bf16 rows in, unit rows out in the caller's dtype (fp32 stays fp32)."""
    if not features.is_cuda:
        raise OvkError("openvision_b200 runs on CUDA (sm_100a) only; got a CPU tensor")
    out_dtype = torch.float32 if features.dtype == torch.float32 else torch.bfloat16
    x = features if features.dtype == torch.bfloat16 else features.to(torch.bfloat16)
    x = x.contiguous()
    y = _Normalize.apply(x, out_dtype, eps) if _needs_grad(x) else ops.l2_normalize(x, out_dtype=out_dtype, eps=eps)
    return y if y.dtype == features.dtype else y.to(features.dtype)


# ----------------------------------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------------------------------
class _AttentionCore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, B, L, H, hd, causal):
        if any(ctx.needs_input_grad):
            out, lse = ops.attention(qkv, B, L, H, hd, save_lse=True, causal=causal)
            ctx.save_for_backward(qkv, out, lse)
            ctx.meta = (B, L, H, hd, causal)
            return out
        return ops.attention(qkv, B, L, H, hd, causal=causal)

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        B, L, H, hd, causal = ctx.meta
        return ops.attention_bwd(qkv, out, _c(dout), lse, B, L, H, hd, causal=causal), None, None, None, None, None


def attention_block_fn(x2, attn, B, L, residual: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
                       causal: bool = False):
    """in_proj GEMM (+bias) -> flash attention -> out_proj GEMM (+bias, +residual)  (module-by-module path)."""
    H = attn.num_heads
    hd = attn.embed_dim // H
    if _needs_grad(x2, attn.in_proj_weight, attn.out_proj.weight, residual):
        qkv = _Linear.apply(x2, attn.in_proj_weight, attn.in_proj_bias, None, _Sub(attn, "in"), False)
        a = _AttentionCore.apply(qkv, B, L, H, hd, causal)
        return _Linear.apply(a, attn.out_proj.weight, attn.out_proj.bias, residual, _Sub(attn, "out"), False)
    wqkv = _w_bf16(attn, "in_w", attn.in_proj_weight)
    bqkv = _v_f32(attn, "in_b", attn.in_proj_bias) if attn.in_proj_bias is not None else None
    wo = _w_bf16(attn, "out_w", attn.out_proj.weight)
    bo = _v_f32(attn, "out_b", attn.out_proj.bias) if attn.out_proj.bias is not None else None
    qkv = ops.gemm(x2, wqkv, bias=bqkv)
    a = ops.attention(qkv, B, L, H, hd, causal=causal)
    return ops.gemm(a, wo, bias=bo, residual=residual, out=out)


class _Sub:
    """distinct packing-cache namespace on one owner module (in_proj / out_proj of the same nn.MultiheadAttention)"""

    def __init__(self, owner, prefix):
        self.__dict__["_ovk_cache"] = owner.__dict__.setdefault("_ovk_cache_" + prefix, {})


# ----------------------------------------------------------------------------------------------------------------
# fused ResidualAttentionBlock
# ----------------------------------------------------------------------------------------------------------------
def _block_params(blk):
    a = blk.attn
    return (blk.ln_1.weight, blk.ln_1.bias, a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias,
            blk.ln_2.weight, blk.ln_2.bias, blk.mlp.c_fc.weight, blk.mlp.c_fc.bias, blk.mlp.c_proj.weight,
            blk.mlp.c_proj.bias)


def _block_packed(blk):
    a, fc, pj = blk.attn, blk.mlp.c_fc, blk.mlp.c_proj
    return dict(
        g1=_v_f32(blk.ln_1, "ln_w", blk.ln_1.weight), b1=_v_f32(blk.ln_1, "ln_b", blk.ln_1.bias),
        wqkv=_w_bf16(a, "in_w", a.in_proj_weight),
        bqkv=_v_f32(a, "in_b", a.in_proj_bias) if a.in_proj_bias is not None else None,
        wo=_w_bf16(a, "out_w", a.out_proj.weight),
        bo=_v_f32(a, "out_b", a.out_proj.bias) if a.out_proj.bias is not None else None,
        g2=_v_f32(blk.ln_2, "ln_w", blk.ln_2.weight), b2=_v_f32(blk.ln_2, "ln_b", blk.ln_2.bias),
        w1=_w_bf16(fc, "w", fc.weight), c1=_v_f32(fc, "b", fc.bias) if fc.bias is not None else None,
        w2=_w_bf16(pj, "w", pj.weight), c2=_v_f32(pj, "b", pj.bias) if pj.bias is not None else None)


def _ln_fold_enabled(D: int) -> bool:
    """LayerNorm is folded into the QKV / fc1 GEMMs when the width fits the statistics epilogue (include/ovk.h,
    ovk_gemm_bf16_ln); OVK_LN_FOLD=0 selects the stand-alone LayerNorm kernels (A/B measurements)."""
    return 128 < D <= 8192 and D % 64 == 0 and os.environ.get("OVK_LN_FOLD", "1") != "0"


def _ln_folded(owner, key, w, b, ln, gamma_f32, beta_f32, b_f32):
    """(Wc bf16 [N,K], d fp32 [N]) from ops.pack_ln_linear, cached on the versions of the parameters involved:
    ln(x) W^T + b = rstd (x Wc^T) + d with Wc = W.gamma, rows centred to zero sum."""
    cache = owner.__dict__.setdefault("_ovk_cache", {})
    tag = tuple(None if t is None else (t.data_ptr(), t._version, t.dtype, t.device) for t in (w, b, ln.weight, ln.bias))
    hit = cache.get(key)
    if hit is not None and hit[0] == tag:
        return hit[1]
    with torch.no_grad():
        wd = w.detach()
        if wd.dtype not in (torch.float32, torch.bfloat16):
            wd = wd.float()
        packed = ops.pack_ln_linear(wd.contiguous(), gamma_f32, beta_f32, b_f32)
    cache[key] = (tag, packed)
    return packed


def _block_forward(x2, p, blk, B, L, inplace, save, stats=None, causal=False):
    """-> (y, saved tensors or None, row statistics of y or None).  `stats` are the partial (sum, sum sq) of the rows of
    x2 when the producer of x2 already had them (the previous block's last GEMM)."""
    H = blk.attn.num_heads
    hd = blk.attn.embed_dim // H
    act = blk._act_kind()
    M, D = x2.shape
    if _ln_fold_enabled(D) and blk.ln_1.eps > 0:
        a_, fc = blk.attn, blk.mlp.c_fc
        wq, dq = _ln_folded(a_, "in_ln", a_.in_proj_weight, a_.in_proj_bias, blk.ln_1, p["g1"], p["b1"], p["bqkv"])
        w1, dd1 = _ln_folded(fc, "w_ln", fc.weight, fc.bias, blk.ln_2, p["g2"], p["b2"], p["c1"])
        if stats is None:
            stats = ops.row_stats(x2)
        parts = (D + 127) // 128
        qkv = ops.gemm_ln(x2, wq, bias=dq, row_stats=stats, eps=blk.ln_1.eps)
        if save:
            a, lse = ops.attention(qkv, B, L, H, hd, save_lse=True, causal=causal)
        else:
            a = ops.attention(qkv, B, L, H, hd, causal=causal)
        st_mid = torch.empty((parts, M, 2), dtype=torch.float32, device=x2.device)
        x_mid = ops.gemm_ln(a, p["wo"], bias=p["bo"], residual=x2, out=x2 if inplace else None, stats_out=st_mid)
        remlp = getattr(blk, "_recompute_mlp_hidden", False)   # selective recompute: u / f are rebuilt in backward
        u = torch.empty((M, w1.shape[0]), dtype=torch.bfloat16, device=x2.device) if save and not remlp else None
        f = ops.gemm_ln(x_mid, w1, bias=dd1, row_stats=st_mid, eps=blk.ln_2.eps, act=act, preact_out=u)
        st_y = torch.empty((parts, M, 2), dtype=torch.float32, device=x2.device)
        y = ops.gemm_ln(f, p["w2"], bias=p["c2"], residual=x_mid, out=None if save else x_mid, stats_out=st_y)
        if save:   # LayerNorm statistics are recomputed in backward together with h1 / h2
            if remlp:
                f = None
            return y, (x2, None, None, qkv, a, lse, x_mid, None, None, u, f), st_y
        return y, None, st_y
    if save:
        h, mean1, rstd1 = ops.layernorm(x2, p["g1"], p["b1"], blk.ln_1.eps, save_stats=True)
    else:
        h = ops.layernorm(x2, p["g1"], p["b1"], blk.ln_1.eps)
    qkv = ops.gemm(h, p["wqkv"], bias=p["bqkv"])
    if save:
        a, lse = ops.attention(qkv, B, L, H, hd, save_lse=True, causal=causal)
    else:
        a = ops.attention(qkv, B, L, H, hd, causal=causal)
    x_mid = ops.gemm(a, p["wo"], bias=p["bo"], residual=x2, out=x2 if inplace else None)
    if save:
        h, mean2, rstd2 = ops.layernorm(x_mid, p["g2"], p["b2"], blk.ln_2.eps, save_stats=True, out=h)
        remlp = getattr(blk, "_recompute_mlp_hidden", False)
        u = None if remlp else torch.empty((x2.shape[0], p["w1"].shape[0]), dtype=torch.bfloat16, device=x2.device)
        f = ops.gemm(h, p["w1"], bias=p["c1"], act=act, preact_out=u)
        y = ops.gemm(f, p["w2"], bias=p["c2"], residual=x_mid)
        if remlp:
            f = None
        return y, (x2, mean1, rstd1, qkv, a, lse, x_mid, mean2, rstd2, u, f), None
    h = ops.layernorm(x_mid, p["g2"], p["b2"], blk.ln_2.eps, out=h)
    f = ops.gemm(h, p["w1"], bias=p["c1"], act=act)
    return ops.gemm(f, p["w2"], bias=p["c2"], residual=x_mid, out=x_mid), None, None


class _Block(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, stats, blk, B, L, causal, *params):
        p = _block_packed(blk)
        y, saved, st_y = _block_forward(x2, p, blk, B, L, inplace=False, save=True, stats=stats, causal=causal)
        ctx.save_for_backward(*saved)
        ctx.blk, ctx.B, ctx.L, ctx.p, ctx.params, ctx.causal = blk, B, L, p, params, causal
        if st_y is None:
            return y, None
        ctx.mark_non_differentiable(st_y)
        return y, st_y

    @staticmethod
    def backward(ctx, dy, _dstats=None):
        x_in, mean1, rstd1, qkv, a, lse, x_mid, mean2, rstd2, u, f = ctx.saved_tensors
        blk, B, L, p, params = ctx.blk, ctx.B, ctx.L, ctx.p, ctx.params
        (ln1_w, ln1_b, in_w, in_b, out_w, out_b, ln2_w, ln2_b, fc_w, fc_b, pj_w, pj_b) = params
        H = blk.attn.num_heads
        hd = blk.attn.embed_dim // H
        act = blk._act_kind()
        dy = _c(dy)
        need_w = any(ctx.needs_input_grad[6:])
        # ---- MLP branch
        h2 = None
        if u is None:   # selective recompute: ln_2 output, then fc1 + GELU with the pre-activation saved
            if mean2 is None:
                h2, mean2, rstd2 = ops.layernorm(x_mid, p["g2"], p["b2"], blk.ln_2.eps, save_stats=True)
            else:
                h2 = ops.layernorm(x_mid, p["g2"], p["b2"], blk.ln_2.eps)
            u = torch.empty((x_mid.shape[0], p["w1"].shape[0]), dtype=torch.bfloat16, device=x_mid.device)
            f = ops.gemm(h2, p["w1"], bias=p["c1"], act=act, preact_out=u)
        fc_b_fused = need_w and fc_b is not None and os.environ.get("OVK_COLSUM_FUSE", "1") != "0"
        if fc_b_fused:   # the same GEMM also leaves db1 = dU.sum(0): no stand-alone pass over the [tokens, 4 D] tensor
            du, g_fc_b = ops.gemm_nn_dact_colsum(dy, p["w2"], u, act)
        else:
            du = ops.gemm_nn(dy, p["w2"], preact=u, act=act)                                # (dY W2) . act'(u)
        g_pj_w = ops.gemm_tn(dy, f, out_dtype=_grad_dtype(pj_w)) if need_w else None
        g_pj_b = ops.colsum(dy) if need_w and pj_b is not None else None
        del f, u
        if h2 is not None:
            pass                # already recomputed above
        elif mean2 is None:   # LayerNorm was folded into the forward GEMMs: statistics come with the recomputed h2
            h2, mean2, rstd2 = ops.layernorm(x_mid, p["g2"], p["b2"], blk.ln_2.eps, save_stats=True)
        else:
            h2 = ops.layernorm(x_mid, p["g2"], p["b2"], blk.ln_2.eps) if need_w else None   # recomputed, not stored
        dh = ops.gemm_nn(du, p["w1"])
        g_fc_w = ops.gemm_tn(du, h2, out_dtype=_grad_dtype(fc_w)) if need_w else None
        if not fc_b_fused:
            g_fc_b = ops.colsum(du) if need_w and fc_b is not None else None
        del du, h2
        dg2, db2 = torch.zeros_like(p["g2"]), torch.zeros_like(p["g2"])
        dxm = ops.layernorm_bwd(dh, x_mid, p["g2"], mean2, rstd2, dg2, db2, dx=dh, dres=dy)  # + residual gradient
        # ---- attention branch
        da = ops.gemm_nn(dxm, p["wo"])
        g_out_w = ops.gemm_tn(dxm, a, out_dtype=_grad_dtype(out_w)) if need_w else None
        g_out_b = ops.colsum(dxm) if need_w and out_b is not None else None
        dqkv = ops.attention_bwd(qkv, a, da, lse, B, L, H, hd, causal=ctx.causal)
        del da
        if mean1 is None:
            h1, mean1, rstd1 = ops.layernorm(x_in, p["g1"], p["b1"], blk.ln_1.eps, save_stats=True)
        else:
            h1 = ops.layernorm(x_in, p["g1"], p["b1"], blk.ln_1.eps) if need_w else None
        dh1 = ops.gemm_nn(dqkv, p["wqkv"])
        g_in_w = ops.gemm_tn(dqkv, h1, out_dtype=_grad_dtype(in_w)) if need_w else None
        g_in_b = ops.colsum(dqkv) if need_w and in_b is not None else None
        del dqkv, h1
        dg1, db1 = torch.zeros_like(p["g1"]), torch.zeros_like(p["g1"])
        dx = ops.layernorm_bwd(dh1, x_in, p["g1"], mean1, rstd1, dg1, db1, dx=dh1, dres=dxm)
        grads = (_like_param(dg1, ln1_w), _like_param(db1, ln1_b), _like_param(g_in_w, in_w), _like_param(g_in_b, in_b),
                 _like_param(g_out_w, out_w), _like_param(g_out_b, out_b), _like_param(dg2, ln2_w),
                 _like_param(db2, ln2_b), _like_param(g_fc_w, fc_w), _like_param(g_fc_b, fc_b),
                 _like_param(g_pj_w, pj_w), _like_param(g_pj_b, pj_b))
        return (dx, None, None, None, None, None) + grads


def block_fn(x2, blk, B, L, inplace, stats=None, causal=False):
    """One ResidualAttentionBlock on a bf16 [B*L, D] residual stream -> (y, row statistics of y or None).
    5 kernels forward with LayerNorm folded into the GEMMs (QKV GEMM, attention, out-proj GEMM+residual, fc1 GEMM+GELU,
    fc2 GEMM+residual; +1 row-statistics kernel in the first block), 7 with the stand-alone LayerNorm kernels.
    `stats`: the statistics the previous block returned for x2, if any."""
    params = _block_params(blk)
    if _needs_grad(x2, *params):
        return _Block.apply(x2, stats, blk, B, L, causal, *params)
    y, _, st_y = _block_forward(x2, _block_packed(blk), blk, B, L, inplace=inplace, save=False, stats=stats, causal=causal)
    return y, st_y

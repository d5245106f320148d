"""Tensor-level wrappers over the libovk C ABI (torch is used for memory and streams only).

Every function takes CUDA tensors, validates dtype / layout, and enqueues exactly the kernels named in its
docstring on torch's current stream.  Nothing here computes on the CPU or falls back to ATen math.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib
from ._lib import (EPI_BIAS, EPI_GELU_ERF, EPI_GELU_QUICK, EPI_GELU_TANH, EPI_NONE, EPI_RESIDUAL, EPI_SAVE_PREACT,
                   OvkError)

_ACT = {None: EPI_NONE, "none": EPI_NONE, "gelu": EPI_GELU_ERF, "gelu_erf": EPI_GELU_ERF, "gelu_tanh": EPI_GELU_TANH,
        "quick_gelu": EPI_GELU_QUICK}

launch_count = 0  # number of libovk kernels enqueued since import (bench.py reports it as gpu_launches)

# Optional per-kernel timing hook (bench.py): callable(kind, work) -> context manager that brackets ONE launch with
# CUDA events on the launching stream.  `work` is the launch's algorithmic FLOPs (tensor kernels) or bytes (bandwidth
# kernels) as defined in DESIGN.md.  None in normal operation: zero overhead.
recorder = None


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL = _Null()


def _timed(kind: str, work: float):
    return recorder(kind, work) if recorder is not None else _NULL


def _count(n=1):
    global launch_count
    launch_count += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _require(t: torch.Tensor, dtype, name: str, ndim: Optional[int] = None):
    if not t.is_cuda:
        raise OvkError(f"{name}: expected a CUDA tensor (libovk has no CPU path)")
    if t.dtype != dtype:
        raise OvkError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise OvkError(f"{name}: expected {ndim} dims, got {tuple(t.shape)}")
    if t.stride(-1) != 1:
        raise OvkError(f"{name}: innermost dimension must be contiguous")
    if t.device.index != torch.cuda.current_device():
        # kernels are enqueued on the CURRENT device's current stream: a tensor of another GPU would be touched from the
        # wrong device / stream (wrap the call in `with torch.cuda.device(t.device):`)
        raise OvkError(f"{name}: tensor lives on cuda:{t.device.index} but the current device is "
                       f"cuda:{torch.cuda.current_device()}")


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, act: Optional[str] = None,
         out: Optional[torch.Tensor] = None, preact_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = act(a[M,K] @ w[N,K]^T + bias)  or  a @ w^T + bias + residual   (kernel: gemm_bf16_kernel, tcgen05).

    `w` has the nn.Linear layout [out_features, in_features]; bias is fp32 [N]; residual may alias `out`.
    With an activation, `preact_out` (bf16 [M,N]) additionally receives a @ w^T + bias (saved for backward)."""
    _require(a, torch.bfloat16, "gemm.a", 2)
    _require(w, torch.bfloat16, "gemm.w", 2)
    M, K = a.shape
    N, K2 = w.shape
    if K2 != K:
        raise OvkError(f"gemm: inner dimensions differ ({K} vs {K2})")
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    else:
        _require(out, torch.bfloat16, "gemm.out", 2)
        if tuple(out.shape) != (M, N):
            raise OvkError("gemm.out has the wrong shape")
    flags = _ACT[act]
    if bias is not None:
        _require(bias, torch.float32, "gemm.bias", 1)
        if bias.numel() != N:
            raise OvkError("gemm.bias has the wrong length")
        flags |= EPI_BIAS
    ldr = 0
    if residual is not None:
        _require(residual, torch.bfloat16, "gemm.residual", 2)
        if tuple(residual.shape) != (M, N):
            raise OvkError("gemm.residual has the wrong shape")
        flags |= EPI_RESIDUAL
        ldr = residual.stride(0)
    ldp = 0
    if preact_out is not None:
        _require(preact_out, torch.bfloat16, "gemm.preact_out", 2)
        if tuple(preact_out.shape) != (M, N):
            raise OvkError("gemm.preact_out has the wrong shape")
        flags |= EPI_SAVE_PREACT
        ldp = preact_out.stride(0)
    with _timed("gemm", 2.0 * M * N * K):
        _lib.call("ovk_gemm_bf16_ex", _p(a), a.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), M, N, K, _p(bias),
                  _p(residual), ldr, _p(preact_out), ldp, flags, _stream())
    _count()
    return out


def gemm_ln(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
            row_stats: Optional[torch.Tensor] = None, eps: float = 1e-6, residual: Optional[torch.Tensor] = None,
            act: Optional[str] = None, out: Optional[torch.Tensor] = None, preact_out: Optional[torch.Tensor] = None,
            stats_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GEMM with LayerNorm folded in and / or row statistics out (kernel: gemm_bf16_kernel<..., FUSE>):
       out = act(rstd * (a @ w^T) + bias) (+ residual), w and bias from pack_ln_linear, rstd from row_stats [P,M,2], the
       partial (sum, sum sq) of the rows of a.  stats_out [ceil(N/128),M,2] fp32 receives the statistics of the output."""
    _require(a, torch.bfloat16, "gemm_ln.a", 2)
    _require(w, torch.bfloat16, "gemm_ln.w", 2)
    M, K = a.shape
    N, K2 = w.shape
    if K2 != K:
        raise OvkError(f"gemm_ln: inner dimensions differ ({K} vs {K2})")
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    flags = _ACT[act]
    if bias is not None:
        _require(bias, torch.float32, "gemm_ln.bias", 1)
        flags |= EPI_BIAS
    parts = 0
    if row_stats is not None:
        _require(row_stats, torch.float32, "gemm_ln.row_stats", 3)
        if tuple(row_stats.shape[1:]) != (M, 2) or not row_stats.is_contiguous():
            raise OvkError("gemm_ln: row_stats must be contiguous [P, M, 2]")
        parts = row_stats.shape[0]
    ldr = 0
    if residual is not None:
        _require(residual, torch.bfloat16, "gemm_ln.residual", 2)
        flags |= EPI_RESIDUAL
        ldr = residual.stride(0)
    ldp = 0
    if preact_out is not None:
        _require(preact_out, torch.bfloat16, "gemm_ln.preact_out", 2)
        flags |= EPI_SAVE_PREACT
        ldp = preact_out.stride(0)
    if stats_out is not None:
        _require(stats_out, torch.float32, "gemm_ln.stats_out", 3)
        if tuple(stats_out.shape) != ((N + 127) // 128, M, 2) or not stats_out.is_contiguous():
            raise OvkError("gemm_ln: stats_out must be contiguous [ceil(N/128), M, 2]")
    with _timed("gemm", 2.0 * M * N * K):
        _lib.call("ovk_gemm_bf16_ln", _p(a), a.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), M, N, K, _p(bias),
                  _p(row_stats), parts, float(eps), _p(residual), ldr, _p(preact_out), ldp, _p(stats_out), flags,
                  _stream())
    _count()
    return out


def pack_ln_linear(w: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bias: Optional[torch.Tensor] = None):
    """(Wc bf16 [N,K], d fp32 [N]) with ln(x) w^T + bias = rstd * (x Wc^T) + d: Wc = w.gamma with every row centred to zero
    sum (also after bf16 rounding), d = w beta + bias.  Kernel: ln_pack_kernel."""
    if w.dtype not in (torch.float32, torch.bfloat16) or w.dim() != 2 or not w.is_cuda or w.stride(1) != 1:
        raise OvkError("pack_ln_linear: w must be a CUDA fp32 / bf16 matrix with unit inner stride")
    N, K = w.shape
    _require(gamma, torch.float32, "pack_ln_linear.gamma", 1)
    _require(beta, torch.float32, "pack_ln_linear.beta", 1)
    if bias is not None:
        _require(bias, torch.float32, "pack_ln_linear.bias", 1)
    if gamma.numel() != K or beta.numel() != K or (bias is not None and bias.numel() != N):
        raise OvkError("pack_ln_linear: gamma / beta must have K entries, bias N")
    wc = torch.empty((N, K), dtype=torch.bfloat16, device=w.device)
    d = torch.empty((N,), dtype=torch.float32, device=w.device)
    with _timed("pack", 4.0 * N * K):
        _lib.call("ovk_pack_ln_linear", _p(w), 1 if w.dtype == torch.float32 else 0, w.stride(0), _p(gamma), _p(beta),
                  _p(bias), _p(wc), wc.stride(0), _p(d), N, K, _stream())
    _count()
    return wc, d


def gemm_rowadd(a: torch.Tensor, w: torch.Tensor, row_add: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[row] = a[row] @ w^T + row_add[row % period]: the patch-embedding GEMM with the class-token / positional-embedding
    add in its epilogue (transformer.py:610-617). row_add: bf16 [period, N]."""
    _require(a, torch.bfloat16, "gemm_rowadd.a", 2)
    _require(w, torch.bfloat16, "gemm_rowadd.w", 2)
    _require(row_add, torch.bfloat16, "gemm_rowadd.row_add", 2)
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K or row_add.shape[1] != N or not row_add.is_contiguous():
        raise OvkError("gemm_rowadd: shapes do not match (a [M,K], w [N,K], row_add contiguous [period,N])")
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    with _timed("gemm", 2.0 * M * N * K):
        _lib.call("ovk_gemm_bf16_rowadd", _p(a), a.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), M, N, K,
                  _p(row_add), row_add.shape[0], _stream())
    _count()
    return out


def row_stats(x: torch.Tensor) -> torch.Tensor:
    """(sum, sum of squares) per row of x bf16 [rows, D] -> fp32 [1, rows, 2] (one partial slot). Kernel: row_stats_kernel."""
    _require(x, torch.bfloat16, "row_stats.x", 2)
    rows, D = x.shape
    st = torch.empty((1, rows, 2), dtype=torch.float32, device=x.device)
    with _timed("layernorm", 2.0 * rows * D):
        _lib.call("ovk_row_stats", _p(x), x.stride(0), _p(st), rows, D, _stream())
    _count()
    return st


def gemm_nn(a: torch.Tensor, b: torch.Tensor, alpha: float = 1.0, out_dtype=torch.bfloat16,
            preact: Optional[torch.Tensor] = None, act: Optional[str] = None,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = alpha * a[M,K] @ b[K,N]  (b row-major: the dgrad dX = dY @ W with W = nn.Linear weight [out, in]).
    With `preact` + `act`: out = alpha * (a @ b) * act'(preact)  (GELU backward fused, bf16 out)."""
    _require(a, torch.bfloat16, "gemm_nn.a", 2)
    _require(b, torch.bfloat16, "gemm_nn.b", 2)
    M, K = a.shape
    K2, N = b.shape
    if K2 != K:
        raise OvkError(f"gemm_nn: inner dimensions differ ({K} vs {K2})")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    _require(out, out_dtype, "gemm_nn.out", 2)
    ldp = 0
    if preact is not None:
        _require(preact, torch.bfloat16, "gemm_nn.preact", 2)
        if tuple(preact.shape) != (M, N):
            raise OvkError("gemm_nn.preact has the wrong shape")
        ldp = preact.stride(0)
    with _timed("gemm", 2.0 * M * N * K):
        _lib.call("ovk_gemm_bf16_nn", _p(a), a.stride(0), _p(b), b.stride(0), _p(out), out.stride(0),
                  int(out_dtype == torch.float32), M, N, K, float(alpha), _p(preact), ldp, _ACT[act], _stream())
    _count()
    return out


def gemm_nn_dact_colsum(a: torch.Tensor, b: torch.Tensor, preact: torch.Tensor, act: str):
    """(out, colsum): out[M,N] = (a[M,K] @ b[K,N]) * act'(preact) in bf16, and colsum[N] = out.sum(0) in fp32 — the bias
    gradient of the layer whose pre-activation gradient `out` is — taken from the GEMM epilogue's staged tiles instead of
    a second pass over out.  Kernels: gemm_bf16_kernel<EPI_DACT> + colsum_f32_kernel."""
    _require(a, torch.bfloat16, "gemm_nn_dact_colsum.a", 2)
    _require(b, torch.bfloat16, "gemm_nn_dact_colsum.b", 2)
    _require(preact, torch.bfloat16, "gemm_nn_dact_colsum.preact", 2)
    M, K = a.shape
    K2, N = b.shape
    if K2 != K or tuple(preact.shape) != (M, N):
        raise OvkError("gemm_nn_dact_colsum: shape mismatch")
    out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    rows = _lib.load().ovk_gemm_colsum_rows(M)
    ws = torch.empty((rows, N), dtype=torch.float32, device=a.device)
    with _timed("gemm", 2.0 * M * N * K):
        _lib.call("ovk_gemm_bf16_nn_dact_colsum", _p(a), a.stride(0), _p(b), b.stride(0), _p(out), out.stride(0), M, N, K, 1.0,
                  _p(preact), preact.stride(0), _ACT[act], _p(ws), _stream())
    _count()
    cs = torch.zeros(N, dtype=torch.float32, device=a.device)
    with _timed("colsum", 4.0 * rows * N):
        _lib.call("ovk_colsum_f32", _p(ws), rows, N, _p(cs), _stream())
    _count()
    return out, cs


def gemm_tn(a: torch.Tensor, b: torch.Tensor, alpha: float = 1.0, out_dtype=torch.bfloat16,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = alpha * a[K,M]^T @ b[K,N]  (both row-major: the wgrad dW = dY^T @ X, the loss dT = s G^T I)."""
    _require(a, torch.bfloat16, "gemm_tn.a", 2)
    _require(b, torch.bfloat16, "gemm_tn.b", 2)
    K, M = a.shape
    K2, N = b.shape
    if K2 != K:
        raise OvkError(f"gemm_tn: reduction dimensions differ ({K} vs {K2})")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    _require(out, out_dtype, "gemm_tn.out", 2)
    with _timed("gemm", 2.0 * M * N * K):
        _lib.call("ovk_gemm_bf16_tn", _p(a), a.stride(0), _p(b), b.stride(0), _p(out), out.stride(0),
                  int(out_dtype == torch.float32), M, N, K, float(alpha), _stream())
    _count()
    return out


def gemm_scaled(a: torch.Tensor, b: torch.Tensor, a_mn: bool = False, b_mn: bool = False, alpha: float = 1.0,
                alpha_dev: Optional[torch.Tensor] = None, out_dtype=torch.float32,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = alpha * alpha_dev * op(a) @ op(b); a stored [M,K] (or [K,M] when a_mn), b stored [N,K] (or [K,N] when b_mn);
    alpha_dev: optional fp32 device scalar (the temperature, never read back by the host). Kernel: gemm_bf16_kernel."""
    _require(a, torch.bfloat16, "gemm_scaled.a", 2)
    _require(b, torch.bfloat16, "gemm_scaled.b", 2)
    K, M = (a.shape if a_mn else a.shape[::-1])
    K2, N = (b.shape if b_mn else b.shape[::-1])
    if K2 != K:
        raise OvkError(f"gemm_scaled: reduction dimensions differ ({K} vs {K2})")
    if alpha_dev is not None:
        _require(alpha_dev, torch.float32, "gemm_scaled.alpha_dev")
        if alpha_dev.numel() != 1:
            raise OvkError("gemm_scaled: alpha_dev must hold one element")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    _require(out, out_dtype, "gemm_scaled.out", 2)
    if tuple(out.shape) != (M, N):
        raise OvkError("gemm_scaled.out has the wrong shape")
    with _timed("gemm", 2.0 * M * N * K):
        _lib.call("ovk_gemm_bf16_scaled", _p(a), a.stride(0), int(a_mn), _p(b), b.stride(0), int(b_mn), _p(out), out.stride(0),
                  int(out_dtype == torch.float32), M, N, K, float(alpha), _p(alpha_dev), _stream())
    _count()
    return out


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a + b (bf16, same shape, contiguous, numel % 8 == 0). Kernel: add_kernel."""
    _require(a, torch.bfloat16, "add.a")
    _require(b, torch.bfloat16, "add.b")
    if a.shape != b.shape or not a.is_contiguous() or not b.is_contiguous():
        raise OvkError("add: operands must be contiguous and of the same shape")
    y = torch.empty_like(a)
    _lib.call("ovk_add_bf16", _p(a), _p(b), _p(y), a.numel(), _stream())
    _count()
    return y


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, save_stats: bool = False,
              out: Optional[torch.Tensor] = None):
    """LayerNorm over the last dim of x [rows, D] (bf16 in/out, fp32 statistics). Kernel: layernorm_fwd_kernel."""
    _require(x, torch.bfloat16, "layernorm.x", 2)
    _require(gamma, torch.float32, "layernorm.gamma", 1)
    _require(beta, torch.float32, "layernorm.beta", 1)
    rows, D = x.shape
    if out is None:
        out = torch.empty((rows, D), dtype=torch.bfloat16, device=x.device)
    mean = rstd = None
    if save_stats:
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    with _timed("layernorm", 4.0 * rows * D):
        _lib.call("ovk_layernorm_fwd", _p(x), x.stride(0), _p(out), out.stride(0), _p(gamma), _p(beta), _p(mean),
                  _p(rstd), rows, D, float(eps), _stream())
    _count()
    if save_stats:
        return out, mean, rstd
    return out


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, dx=None, dres=None):
    """dx (+ dres: the gradient coming through the residual connection), and (accumulated) dgamma / dbeta of LayerNorm.
    Kernel: layernorm_bwd_kernel."""
    _require(dy, torch.bfloat16, "layernorm_bwd.dy", 2)
    _require(x, torch.bfloat16, "layernorm_bwd.x", 2)
    rows, D = x.shape
    if dx is None:
        dx = torch.empty((rows, D), dtype=torch.bfloat16, device=x.device)
    if dres is not None:
        _require(dres, torch.bfloat16, "layernorm_bwd.dres", 2)
    with _timed("layernorm_bwd", (8.0 if dres is not None else 6.0) * rows * D):
        _lib.call("ovk_layernorm_bwd", _p(dy), dy.stride(0), _p(x), x.stride(0), _p(gamma), _p(mean), _p(rstd), _p(dres),
                  dres.stride(0) if dres is not None else 0, _p(dx), dx.stride(0), _p(dgamma), _p(dbeta), rows, D,
                  _stream())
    _count(2)
    return dx


def im2col_patches(images: torch.Tensor, patch: int, ldc: int, lead_rows: int = 0) -> torch.Tensor:
    """images [B,3,H,W] (fp32 or bf16, NCHW contiguous) -> bf16 [B*(gh*gw + lead_rows), ldc], columns ordered like
    conv1.weight.reshape(D, 3*P*P), zero padded to ldc; lead_rows=1 inserts a zero row per image (the cls slot).
    Kernel: im2col_kernel."""
    if images.dtype not in (torch.float32, torch.bfloat16):
        raise OvkError(f"im2col: images must be fp32 or bf16, got {images.dtype}")
    _require(images, images.dtype, "im2col.images", 4)
    if not images.is_contiguous():
        raise OvkError("im2col: images must be NCHW-contiguous")
    B, C, H, W = images.shape
    if C != 3:
        raise OvkError("im2col: expected 3 input channels")
    rows = B * ((H // patch) * (W // patch) + lead_rows)
    cols = torch.empty((rows, ldc), dtype=torch.bfloat16, device=images.device)
    _lib.call("ovk_im2col_patches", _p(images), int(images.dtype == torch.float32), _p(cols), ldc, B, H, W, patch,
              int(lead_rows), _stream())
    _count()
    return cols


def patch_embed_supported(images: torch.Tensor, patch: int, D: int) -> bool:
    """True when the one-kernel patch embedding (patch_embed_kernel) handles this geometry."""
    if images.dim() != 4 or images.shape[1] != 3 or images.dtype not in (torch.float32, torch.bfloat16):
        return False
    return bool(_lib.load().ovk_patch_embed_supported(int(images.dtype == torch.float32), images.shape[2], images.shape[3],
                                                      int(patch), int(D)))


def pack_patch_weight(w: torch.Tensor, patch: int) -> torch.Tensor:
    """conv1.weight [D,3,P,P] -> bf16 [D, K'] in the k-order of patch_embed_kernel (include/ovk.h, ovk_patch_embed):
    (channel, row group, row in group, pixel padded to PW), zero in the padding columns.  Layout plumbing on torch."""
    D = w.shape[0]
    P = int(patch)
    PW = 16 if P <= 16 else 32
    R = 64 // PW
    PG = (P + R - 1) // R
    out = torch.zeros((D, 3, PG * R, PW), dtype=torch.bfloat16, device=w.device)
    out[:, :, :P, :P] = w.detach().to(torch.bfloat16)
    out = out.reshape(D, 3 * PG * R * PW)
    assert out.shape[1] == _lib.load().ovk_patch_embed_kdim(P)
    return out.contiguous()


def patch_embed(images: torch.Tensor, w_packed: torch.Tensor, patch: int, table: Optional[torch.Tensor] = None) -> torch.Tensor:
    """images [B,3,H,W] (fp32 / bf16 NCHW) -> tokens bf16 [B, N+1, D]: conv1 as a TMA-staged im2col GEMM with the
    class-token / positional-embedding add in its epilogue (table bf16 [N+1, D]; None: plain conv tokens, zero class row).
    Kernel: patch_embed_kernel."""
    if images.dtype not in (torch.float32, torch.bfloat16):
        raise OvkError(f"patch_embed: images must be fp32 or bf16, got {images.dtype}")
    _require(images, images.dtype, "patch_embed.images", 4)
    _require(w_packed, torch.bfloat16, "patch_embed.w_packed", 2)
    if not images.is_contiguous() or not w_packed.is_contiguous():
        raise OvkError("patch_embed: images must be NCHW-contiguous and the packed weight contiguous")
    B, C, H, W = images.shape
    D = w_packed.shape[0]
    if C != 3 or not patch_embed_supported(images, patch, D):
        raise OvkError(f"patch_embed: unsupported geometry {tuple(images.shape)} patch {patch} width {D}")
    if w_packed.shape[1] != _lib.load().ovk_patch_embed_kdim(int(patch)):
        raise OvkError("patch_embed: w_packed must come from pack_patch_weight")
    N = (H // patch) * (W // patch)
    if table is not None:
        _require(table, torch.bfloat16, "patch_embed.table", 2)
        if tuple(table.shape) != (N + 1, D) or not table.is_contiguous():
            raise OvkError("patch_embed: table must be contiguous [N+1, D]")
    tokens = torch.empty((B, N + 1, D), dtype=torch.bfloat16, device=images.device)
    with _timed("gemm", 2.0 * B * N * D * 3 * patch * patch):
        _lib.call("ovk_patch_embed", _p(images), int(images.dtype == torch.float32), _p(w_packed), _p(table), _p(tokens),
                  B, H, W, int(patch), D, _stream())
    _count()
    return tokens


def embed_assemble(patch_tokens: torch.Tensor, cls: torch.Tensor, pos: torch.Tensor, B: int, N: int,
                   inplace: bool = False) -> torch.Tensor:
    """tokens[b,0]=cls+pos[0]; tokens[b,1+n]=patch[b,n]+pos[1+n]  -> bf16 [B, N+1, D]. patch_tokens is [B*N, D] or, with
    one lead row per image (GEMM output of im2col(lead_rows=1)), [B*(N+1), D]; the latter may be updated in place.
    Kernel: embed_assemble_kernel."""
    _require(patch_tokens, torch.bfloat16, "embed_assemble.patch", 2)
    _require(cls, torch.float32, "embed_assemble.cls", 1)
    _require(pos, torch.float32, "embed_assemble.pos", 2)
    D = patch_tokens.shape[1]
    if not patch_tokens.is_contiguous() or not pos.is_contiguous():
        raise OvkError("embed_assemble: inputs must be contiguous")
    lead = 1 if patch_tokens.shape[0] == B * (N + 1) else 0
    if patch_tokens.shape[0] != B * (N + lead) or tuple(pos.shape) != (N + 1, D) or cls.numel() != D:
        raise OvkError("embed_assemble: shape mismatch")
    if inplace and not lead:
        raise OvkError("embed_assemble: in-place needs the [B*(N+1), D] layout")
    tokens = patch_tokens.view(B, N + 1, D) if inplace else torch.empty((B, N + 1, D), dtype=torch.bfloat16,
                                                                        device=patch_tokens.device)
    _lib.call("ovk_embed_assemble", _p(patch_tokens), lead, _p(cls), _p(pos), _p(tokens), B, N, D, _stream())
    _count()
    return tokens


def attention(qkv: torch.Tensor, B: int, L: int, H: int, hd: int, scale: Optional[float] = None,
              save_lse: bool = False, causal: bool = False):
    """qkv bf16 [B*L, 3*H*hd] (in_proj output) -> bf16 [B*L, H*hd]; optional lse fp32 [B,H,L].
    Kernel: attention_fwd_kernel (tcgen05 flash attention)."""
    _require(qkv, torch.bfloat16, "attention.qkv", 2)
    if not qkv.is_contiguous() or tuple(qkv.shape) != (B * L, 3 * H * hd):
        raise OvkError(f"attention: qkv must be contiguous [B*L, 3*H*hd], got {tuple(qkv.shape)}")
    out = torch.empty((B * L, H * hd), dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty((B, H, L), dtype=torch.float32, device=qkv.device) if save_lse else None
    if scale is None:
        scale = 1.0 / math.sqrt(hd)
    with _timed("attention", 4.0 * B * H * L * L * hd):
        _lib.call("ovk_attention_fwd_ex", _p(qkv), _p(out), _p(lse), B, L, H, hd, float(scale), int(bool(causal)), _stream())
    _count()
    return (out, lse) if save_lse else out


def pool_tokens(x: torch.Tensor, mode: str) -> torch.Tensor:
    """x bf16 [B,L,D] -> bf16 [B,D]; mode 'avg' = mean over tokens 1.. (cls excluded), 'tok' = token 0."""
    _require(x, torch.bfloat16, "pool_tokens.x", 3)
    if not x.is_contiguous():
        raise OvkError("pool_tokens: x must be contiguous")
    B, L, D = x.shape
    pooled = torch.empty((B, D), dtype=torch.bfloat16, device=x.device)
    _lib.call("ovk_pool_tokens", _p(x), _p(pooled), B, L, D, {"avg": 0, "tok": 1}[mode], _stream())
    _count()
    return pooled


def pool_head(x: torch.Tensor, mode: str, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], eps: float,
              proj: Optional[torch.Tensor], normalize: bool = False, out_dtype=torch.bfloat16,
              norm_eps: float = 1e-12) -> torch.Tensor:
    """x bf16 [B,L,D] -> [B,E]: token pooling ('avg' over tokens 1.. / 'tok' = 'first' / 'last'), LayerNorm (gamma / beta fp32, or None),
    @ proj (bf16 [D,E], or None), optional F.normalize, in one launch.  Kernel: pool_head_kernel."""
    _require(x, torch.bfloat16, "pool_head.x", 3)
    if not x.is_contiguous():
        raise OvkError("pool_head: x must be contiguous")
    B, L, D = x.shape
    if (gamma is None) != (beta is None):
        raise OvkError("pool_head: gamma and beta come together")
    if gamma is not None:
        _require(gamma, torch.float32, "pool_head.gamma", 1)
        _require(beta, torch.float32, "pool_head.beta", 1)
        if gamma.numel() != D or beta.numel() != D:
            raise OvkError("pool_head: gamma / beta must have D entries")
    E = D
    if proj is not None:
        _require(proj, torch.bfloat16, "pool_head.proj", 2)
        if proj.shape[0] != D or not proj.is_contiguous():
            raise OvkError("pool_head: proj must be contiguous [D, E]")
        E = proj.shape[1]
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise OvkError("pool_head: output dtype must be fp32 or bf16")
    out = torch.empty((B, E), dtype=out_dtype, device=x.device)
    with _timed("pool_head", 2.0 * B * L * D):
        _lib.call("ovk_pool_head", _p(x), B, L, D, {"avg": 0, "tok": 1, "first": 1, "last": 2}[mode], _p(gamma), _p(beta), float(eps), _p(proj), E,
                  int(bool(normalize)), float(norm_eps), _p(out), int(out_dtype == torch.float32), _stream())
    _count()
    return out


def l2_normalize(x: torch.Tensor, out_dtype=torch.float32, eps: float = 1e-12, return_norms: bool = False):
    """F.normalize(x, dim=-1) for x bf16 [rows,E]; output fp32 or bf16. Kernel: l2_normalize_kernel."""
    _require(x, torch.bfloat16, "l2_normalize.x", 2)
    if not x.is_contiguous():
        raise OvkError("l2_normalize: x must be contiguous")
    rows, E = x.shape
    y = torch.empty((rows, E), dtype=out_dtype, device=x.device)
    norms = torch.empty(rows, dtype=torch.float32, device=x.device) if return_norms else None
    _lib.call("ovk_l2_normalize", _p(x), _p(y), int(out_dtype == torch.float32), _p(norms), rows, E, float(eps), _stream())
    _count()
    return (y, norms) if return_norms else y


# ----------------------------------------------------------------------------------------------------------------
# fused CLIP contrastive loss (kernels: clip_loss_fwd_kernel, clip_loss_finalize_kernel, clip_loss_combine_kernel,
# clip_loss_value_kernel, clip_loss_grad_kernel)
# ----------------------------------------------------------------------------------------------------------------
def _scalar_dev(x, device, name: str) -> torch.Tensor:
    """fp32[1] device tensor from a python number or a tensor (tensors stay on the device: no host read-back)."""
    if torch.is_tensor(x):
        t = x.detach()
        if t.numel() != 1:
            raise OvkError(f"{name}: expected a scalar")
        if not t.is_cuda:
            t = t.to(device)
        return t.to(torch.float32).reshape(1)
    return torch.full((1,), float(x), dtype=torch.float32, device=device)


CL_WINDOW = 256   # column windows of clip_loss_fwd start on multiples of this (the kernel's tile width)


def clip_loss_workspace(n_loc: int, n_all: int, device) -> torch.Tensor:
    return torch.empty(_lib.load().ovk_clip_loss_workspace_floats(n_loc, n_all), dtype=torch.float32, device=device)


def clip_loss_fwd_window(a_loc: torch.Tensor, b_rows: torch.Tensor, col_offset: int, n_all: int, row_offset: int, scale,
                         diag: torch.Tensor, ws: torch.Tensor) -> None:
    """Partial statistics of the column window [col_offset, col_offset + b_rows.shape[0]) of z = scale * a_loc @ b_all^T
    (b_rows = the features of those columns) into the workspace; see include/ovk.h.  Kernel: clip_loss_fwd_kernel."""
    _require(a_loc, torch.bfloat16, "clip_loss.a_loc", 2)
    _require(b_rows, torch.bfloat16, "clip_loss.b_rows", 2)
    if not a_loc.is_contiguous() or not b_rows.is_contiguous():
        raise OvkError("clip_loss: features must be contiguous")
    n_loc, E = a_loc.shape
    n_cols, E2 = b_rows.shape
    if E != E2:
        raise OvkError("clip_loss: embedding dims differ")
    scale = _scalar_dev(scale, a_loc.device, "clip_loss.scale")
    with _timed("clip_loss_fwd", 2.0 * n_loc * n_cols * E):
        _lib.call("ovk_clip_loss_fwd", _p(a_loc), _p(b_rows), n_loc, n_cols, int(col_offset), int(n_all), E, int(row_offset),
                  _p(scale), _p(diag), _p(ws), _stream())
    _count()


def clip_loss_finalize(ws: torch.Tensor, n_loc: int, n_all: int):
    """Merge the per-tile partials of all column windows -> (row_lse, col_max, col_sum), fp32. Kernel: clip_loss_finalize_kernel."""
    dev = ws.device
    row_lse = torch.empty(n_loc, dtype=torch.float32, device=dev)
    col_max = torch.empty(n_all, dtype=torch.float32, device=dev)
    col_sum = torch.empty(n_all, dtype=torch.float32, device=dev)
    _lib.call("ovk_clip_loss_finalize", _p(ws), n_loc, n_all, _p(row_lse), _p(col_max), _p(col_sum), _stream())
    _count()
    return row_lse, col_max, col_sum


def clip_loss_fwd(a_loc: torch.Tensor, b_all: torch.Tensor, row_offset: int, scale):
    """Row LSEs, positive-pair logits and this row block's column statistics of z = scale * a_loc @ b_all^T without
    materialising z.  a_loc bf16 [n_loc,E], b_all bf16 [n_all,E]; scale: python number or device scalar.
    Returns (row_lse, diag, col_max, col_sum), fp32."""
    n_loc, n_all = a_loc.shape[0], b_all.shape[0]
    diag = torch.empty(n_loc, dtype=torch.float32, device=a_loc.device)
    ws = clip_loss_workspace(n_loc, n_all, a_loc.device)
    clip_loss_fwd_window(a_loc, b_all, 0, n_all, row_offset, scale, diag, ws)
    row_lse, col_max, col_sum = clip_loss_finalize(ws, n_loc, n_all)
    return row_lse, diag, col_max, col_sum


def clip_loss_combine(col_max_parts: torch.Tensor, col_sum_parts: torch.Tensor) -> torch.Tensor:
    """[parts, n_all] partial column statistics (one row per rank) -> global column LSE [n_all]."""
    _require(col_max_parts, torch.float32, "clip_loss_combine.col_max", 2)
    _require(col_sum_parts, torch.float32, "clip_loss_combine.col_sum", 2)
    if not col_max_parts.is_contiguous() or not col_sum_parts.is_contiguous():
        raise OvkError("clip_loss_combine: inputs must be contiguous")
    parts, n_all = col_max_parts.shape
    col_lse = torch.empty(n_all, dtype=torch.float32, device=col_max_parts.device)
    _lib.call("ovk_clip_loss_combine", _p(col_max_parts), _p(col_sum_parts), parts, n_all, _p(col_lse), _stream())
    _count()
    return col_lse


def clip_loss_value(row_lse, col_lse, diag, row_offset: int) -> torch.Tensor:
    """out[0] = 0.5/n * (sum(row_lse - diag) + sum(col_lse[labels] - diag)); out[1:3] = the two sums."""
    out = torch.empty(3, dtype=torch.float32, device=row_lse.device)
    _lib.call("ovk_clip_loss_value", _p(row_lse), _p(col_lse), _p(diag), row_lse.numel(), int(row_offset), _p(out), _stream())
    _count()
    return out


def clip_loss_grad_logits(a_loc, b_all, row_offset: int, scale, row_lse, col_lse, w_row: float, w_col: float,
                          d_scale: torch.Tensor, grad_out=None) -> torch.Tensor:
    """G = dL/dz for this row block, bf16 [n_loc, n_all]; d_scale (fp32[1]) += sum(G * z) / scale.  `scale` and the optional
    upstream gradient `grad_out` (multiplied into both weights) may be device scalars."""
    n_loc, E = a_loc.shape
    n_all = b_all.shape[0]
    ldg = (n_all + 7) // 8 * 8
    G = torch.empty((n_loc, ldg), dtype=torch.bfloat16, device=a_loc.device)
    scale = _scalar_dev(scale, a_loc.device, "clip_loss_grad.scale")
    go = None if grad_out is None else _scalar_dev(grad_out, a_loc.device, "clip_loss_grad.grad_out")
    with _timed("clip_loss_grad", 2.0 * n_loc * n_all * E):
        _lib.call("ovk_clip_loss_grad_logits", _p(a_loc), _p(b_all), n_loc, n_all, E, int(row_offset), _p(scale),
                  _p(row_lse), _p(col_lse), float(w_row), float(w_col), _p(go), _p(G), ldg, _p(d_scale), _stream())
    _count()
    return G[:, :n_all] if ldg != n_all else G


# ----------------------------------------------------------------------------------------------------------------
# backward kernels of the image tower
# ----------------------------------------------------------------------------------------------------------------
def attention_bwd(qkv, out, dout, lse, B: int, L: int, H: int, hd: int, scale: Optional[float] = None,
                  causal: bool = False) -> torch.Tensor:
    """dqkv bf16 [B*L, 3*H*hd] from qkv, the forward output, its gradient and the saved lse.
    Kernels: attention_bwd_delta_kernel, attention_bwd_kernel<fused>, attention_bwd_dq_convert_kernel (one pass over the
    score tiles); OVK_ATTBWD_FUSED=0: attention_bwd_kernel<dQ>, attention_bwd_kernel<dKdV> (two passes, no scratch)."""
    for t, name, cols in ((qkv, "qkv", 3 * H * hd), (out, "out", H * hd), (dout, "dout", H * hd)):
        _require(t, torch.bfloat16, f"attention_bwd.{name}", 2)
        if not t.is_contiguous() or tuple(t.shape) != (B * L, cols):
            raise OvkError(f"attention_bwd: {name} must be contiguous [B*L, {cols}], got {tuple(t.shape)}")
    _require(lse, torch.float32, "attention_bwd.lse", 3)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty((B, H, L), dtype=torch.float32, device=qkv.device)
    if scale is None:
        scale = 1.0 / math.sqrt(hd)
    import os
    flags = int(bool(causal))
    tail_off = os.environ.get("OVK_ATTBWD_TAIL", "1") == "0"   # A/B: the remainder token as a third tile row / column
    nws = _lib.load().ovk_attention_bwd_workspace_floats(B, L, H, flags)
    nfused = _lib.load().ovk_attention_bwd_fused_workspace_floats(B, L, H, hd, flags)
    # measured (tools/attn_bwd_ab.py): L = 577, hd 64: 4.32 -> 3.40 ms; L = 257: 3.81 -> 3.43 ms; hd 80: no gain (4.89 vs 4.94)
    mode = os.environ.get("OVK_ATTBWD_FUSED", "2")   # 2: attention_bwd_t_kernel, 1: attention_bwd_kernel<fused>, 0: two passes
    if mode != "0" and nfused > 0 and not tail_off:
        if mode == "1":
            flags |= 2   # OVK_ATT_BWD_ONEPASS_V1
        if os.environ.get("OVK_ATTBWD_WARPS", "8") != "16":   # measured: 16 compute warps are 7-10 % slower
            flags |= 4   # OVK_ATT_BWD_8_WARPS
        flags |= (int(os.environ.get("OVK_ATTBWD_DBG", "0")) & 0xff) << 8   # knock-outs (tools/attn_bwd_knockout.py)
        # one pass over the score tiles; dQ partial sums through an fp32 scratch (attention_bwd.cu, MODE_FUSED)
        ws = torch.empty(nfused, dtype=torch.float32, device=qkv.device)
        with _timed("attention_bwd", 10.0 * B * H * L * L * hd):
            _lib.call("ovk_attention_bwd_fused", _p(qkv), _p(out), _p(dout), _p(lse), _p(dqkv), _p(delta), _p(ws), B, L, H, hd,
                      float(scale), flags, _stream())
        _count(3)   # delta kernel OR remainder-token kernel (which leaves delta too), tile kernel, dQ conversion
        return dqkv
    flags &= 1
    if tail_off:
        nws = 0
    ws = torch.empty(nws, dtype=torch.float32, device=qkv.device) if nws > 0 else None   # remainder token of L = 128 k + 1
    with _timed("attention_bwd", 14.0 * B * H * L * L * hd):
        _lib.call("ovk_attention_bwd_ex", _p(qkv), _p(out), _p(dout), _p(lse), _p(dqkv), _p(delta), _p(ws), B, L, H, hd,
                  float(scale), flags, _stream())
    _count(3 if ws is not None else 2)
    return dqkv


def colsum(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[c] (+)= sum_r x[r,c]; x bf16 [rows, cols] -> fp32 [cols]. Kernel: colsum_kernel."""
    _require(x, torch.bfloat16, "colsum.x", 2)
    rows, cols = x.shape
    if out is None:
        out = torch.zeros(cols, dtype=torch.float32, device=x.device)
    with _timed("colsum", 2.0 * rows * cols):
        _lib.call("ovk_colsum_bf16", _p(x), x.stride(0), rows, cols, _p(out), _stream())
    _count()
    return out


def pool_tokens_bwd(dpooled: torch.Tensor, B: int, L: int, mode: str) -> torch.Tensor:
    """gradient of pool_tokens: bf16 [B,D] -> bf16 [B,L,D] (every element written). Kernel: pool_tokens_bwd_kernel."""
    _require(dpooled, torch.bfloat16, "pool_tokens_bwd.dpooled", 2)
    if not dpooled.is_contiguous():
        raise OvkError("pool_tokens_bwd: dpooled must be contiguous")
    D = dpooled.shape[1]
    dx = torch.empty((B, L, D), dtype=torch.bfloat16, device=dpooled.device)
    _lib.call("ovk_pool_tokens_bwd", _p(dpooled), _p(dx), B, L, D, {"avg": 0, "tok": 1}[mode], _stream())
    _count()
    return dx


def l2_normalize_bwd(x: torch.Tensor, dy: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """gradient of F.normalize w.r.t. x (bf16 [rows,E]); dy fp32 or bf16. Kernel: l2_normalize_bwd_kernel."""
    _require(x, torch.bfloat16, "l2_normalize_bwd.x", 2)
    if dy.dtype not in (torch.float32, torch.bfloat16):
        raise OvkError("l2_normalize_bwd: dy must be fp32 or bf16")
    _require(dy, dy.dtype, "l2_normalize_bwd.dy", 2)
    if not x.is_contiguous() or not dy.is_contiguous() or dy.shape != x.shape:
        raise OvkError("l2_normalize_bwd: x and dy must be contiguous and of the same shape")
    rows, E = x.shape
    dx = torch.empty_like(x)
    _lib.call("ovk_l2_normalize_bwd", _p(x), _p(dy), int(dy.dtype == torch.float32), _p(dx), rows, E, float(eps), _stream())
    _count()
    return dx


def col2im_patches(dcols: torch.Tensor, B: int, H: int, W: int, patch: int, lead_rows: int, out_dtype) -> torch.Tensor:
    """gradient w.r.t. the images from the gradient of the im2col matrix. Kernel: col2im_kernel."""
    _require(dcols, torch.bfloat16, "col2im.dcols", 2)
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise OvkError("col2im: output dtype must be fp32 or bf16")
    dimg = torch.empty((B, 3, H, W), dtype=out_dtype, device=dcols.device)
    _lib.call("ovk_col2im_patches", _p(dcols), dcols.stride(0), _p(dimg), int(out_dtype == torch.float32), B, H, W, patch,
              int(lead_rows), _stream())
    _count()
    return dimg


def act_fwd(x: torch.Tensor, act: str) -> torch.Tensor:
    """y = act(x) elementwise (bf16; numel % 8 == 0). Kernel: act_kernel<fwd>."""
    _require(x, torch.bfloat16, "act_fwd.x")
    if not x.is_contiguous():
        raise OvkError("act_fwd: x must be contiguous")
    y = torch.empty_like(x)
    _lib.call("ovk_act_fwd", _p(x), _p(y), x.numel(), _ACT[act], _stream())
    _count()
    return y


def act_bwd(x: torch.Tensor, dy: torch.Tensor, act: str) -> torch.Tensor:
    """dx = dy * act'(x) elementwise (bf16). Kernel: act_kernel<bwd>."""
    _require(x, torch.bfloat16, "act_bwd.x")
    _require(dy, torch.bfloat16, "act_bwd.dy")
    if not x.is_contiguous() or not dy.is_contiguous() or x.shape != dy.shape:
        raise OvkError("act_bwd: x and dy must be contiguous and of the same shape")
    dx = torch.empty_like(x)
    _lib.call("ovk_act_bwd", _p(x), _p(dy), _p(dx), x.numel(), _ACT[act], _stream())
    _count()
    return dx


# ----------------------------------------------------------------------------------------------------------------
# optimizer step on flat buffers (src/optim/build_optax.py:188-278)
# ----------------------------------------------------------------------------------------------------------------
def sumsq(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out (fp32 scalar tensor, ACCUMULATED) += sum x^2 over a flat fp32 / bf16 buffer. Kernel: sumsq_kernel."""
    if x.dtype not in (torch.float32, torch.bfloat16) or not x.is_cuda or not x.is_contiguous():
        raise OvkError("sumsq: x must be a contiguous CUDA fp32 / bf16 tensor")
    _require(out, torch.float32, "sumsq.out")
    with _timed("optimizer", float(x.numel())):
        _lib.call("ovk_sumsq", _p(x), 1 if x.dtype == torch.bfloat16 else 0, x.numel(), _p(out), _stream())
    _count()
    return out


def adamw_step(p: torch.Tensor, g: torch.Tensor, mu: torch.Tensor, nu: torch.Tensor, lr: float, b1: float, b2: float,
               eps: float, wd: float, step: int, gscale: float = 1.0, gnorm_sq: Optional[torch.Tensor] = None,
               max_norm: float = 0.0) -> None:
    """In-place scale_by_adam (bf16 first moment) + decoupled weight decay + lr on flat buffers. Kernel: adamw_kernel."""
    for t, name in ((p, "p"), (g, "g")):
        if t.dtype not in (torch.float32, torch.bfloat16) or not t.is_cuda or not t.is_contiguous():
            raise OvkError(f"adamw_step: {name} must be a contiguous CUDA fp32 / bf16 tensor")
    _require(mu, torch.bfloat16, "adamw_step.mu")
    _require(nu, torch.float32, "adamw_step.nu")
    n = p.numel()
    if g.numel() != n or mu.numel() != n or nu.numel() != n:
        raise OvkError("adamw_step: p, g, mu, nu must have the same number of elements")
    if gnorm_sq is not None:
        _require(gnorm_sq, torch.float32, "adamw_step.gnorm_sq")
    with _timed("optimizer", float(n)):
        _lib.call("ovk_adamw_step", _p(p), 1 if p.dtype == torch.bfloat16 else 0, _p(g), 1 if g.dtype == torch.bfloat16 else 0,
                  _p(mu), _p(nu), n, float(lr), float(b1), float(b2), float(eps), float(wd), int(step), float(gscale),
                  _p(gnorm_sq), float(max_norm), _stream())
    _count()

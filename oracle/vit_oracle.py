"""CPU ORACLE — test infrastructure, not product code.

A plain restatement (explicit tensor math, fp32 or fp64 on the CPU) of the reference's algorithm for the hot path:
the open_clip ViT image tower, the transformer text tower, CLIP.forward and ClipLoss.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; the
product package `openvision_b200` never does.

The arithmetic of the reference lives in a third-party dependency (PyTorch, version unpinned by the reference;
SURVEY.md §8c): F.conv2d, F.layer_norm, nn.MultiheadAttention, F.linear, nn.GELU, F.normalize, F.cross_entropy.  Each
function below restates the published definition of the op at the reference call site it cites (paths relative to
/root/reference/src/convert_upload/open_clip/).  The reference ships no golden vectors for this path; the oracle is
PINNED by `oracle/make_golden.py`, which imports the unmodified reference modules in the build container, checks this
file against them (fp64 agreement <= 1e-9, fp32 <= 1e-4) and commits the reference's outputs under `tests/golden/`.

Parameters are addressed with the reference's state_dict key names.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------------------------
def layer_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float) -> Tensor:
    """transformer.py:15-30 (LayerNorm / LayerNormFp32 -> F.layer_norm): biased variance over the last dim."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * weight + bias


def gelu(x: Tensor, kind: str = "erf") -> Tensor:
    """transformer.py:232-236 act_layer: nn.GELU() exact-erf (image tower), nn.GELU(approximate='tanh') (OpenVision
    text tower act_kwargs), QuickGELU transformer.py:33-36."""
    if kind == "erf":
        return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))
    if kind == "tanh":
        return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))
    if kind == "quick":
        return x * torch.sigmoid(1.702 * x)
    raise ValueError(kind)


def multi_head_attention(x: Tensor, in_w: Tensor, in_b: Tensor, out_w: Tensor, out_b: Tensor, heads: int,
                         attn_mask: Optional[Tensor] = None) -> Tensor:
    """transformer.py:225,239-252: nn.MultiheadAttention(d, heads, batch_first=True)(x, x, x, need_weights=False).
    in_proj_weight rows are q|k|v; head h uses columns [h*hd, (h+1)*hd); scores scaled by hd^-0.5; softmax over keys;
    optional additive mask (text tower causal mask, transformer.py:757-763)."""
    B, L, D = x.shape
    hd = D // heads
    qkv = x @ in_w.t() + in_b
    q, k, v = qkv.split(D, dim=-1)
    q = q.reshape(B, L, heads, hd).transpose(1, 2)
    k = k.reshape(B, L, heads, hd).transpose(1, 2)
    v = v.reshape(B, L, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
    if attn_mask is not None:
        s = s + attn_mask
    s = s - s.max(dim=-1, keepdim=True).values
    p = torch.exp(s)
    p = p / p.sum(dim=-1, keepdim=True)
    o = (p @ v).transpose(1, 2).reshape(B, L, D)
    return o @ out_w.t() + out_b


def residual_attention_block(x: Tensor, sd: Dict[str, Tensor], prefix: str, heads: int, eps: float, act: str,
                             attn_mask: Optional[Tensor] = None, taps: Optional[dict] = None) -> Tensor:
    """transformer.py:254-265: x = x + attn(ln_1(x)); x = x + c_proj(gelu(c_fc(ln_2(x)))); ls_1/ls_2 = Identity."""
    h = layer_norm(x, sd[prefix + "ln_1.weight"], sd[prefix + "ln_1.bias"], eps)
    x = x + multi_head_attention(h, sd[prefix + "attn.in_proj_weight"], sd[prefix + "attn.in_proj_bias"],
                                 sd[prefix + "attn.out_proj.weight"], sd[prefix + "attn.out_proj.bias"], heads, attn_mask)
    h = layer_norm(x, sd[prefix + "ln_2.weight"], sd[prefix + "ln_2.bias"], eps)
    h = gelu(h @ sd[prefix + "mlp.c_fc.weight"].t() + sd[prefix + "mlp.c_fc.bias"], act)
    if taps is not None:
        taps[prefix + "mlp.gelu"] = h
    x = x + (h @ sd[prefix + "mlp.c_proj.weight"].t() + sd[prefix + "mlp.c_proj.bias"])
    return x


def _num_layers(sd: Dict[str, Tensor], prefix: str) -> int:
    n = 0
    while f"{prefix}resblocks.{n}.ln_1.weight" in sd:
        n += 1
    return n


# ------------------------------------------------------------------------------------------------------------------
# image tower
# ------------------------------------------------------------------------------------------------------------------
def patch_embed(images: Tensor, conv_w: Tensor) -> Tensor:
    """transformer.py:469,610-612: Conv2d(3->D, kernel=stride=P, bias=False), then [B,D,g,g] -> [B, g*g, D] with
    patches row-major over the grid; weight flattening order (c, ph, pw)."""
    B, C, H, W = images.shape
    D, _, P, _ = conv_w.shape
    gh, gw = H // P, W // P
    x = images.reshape(B, C, gh, P, gw, P).permute(0, 2, 4, 1, 3, 5).reshape(B, gh * gw, C * P * P)
    return x @ conv_w.reshape(D, C * P * P).t()


def vision_transformer(images: Tensor, sd: Dict[str, Tensor], heads: int, *, prefix: str = "visual.",
                       pool_type: str = "avg", final_ln_after_pool: bool = True, eps: float = 1e-6,
                       act: str = "erf", taps: Optional[dict] = None) -> Tensor:
    """VisionTransformer.forward, transformer.py:609-651 (attn_pool=None branch).
    OpenVision cfg: no ln_pre, pool 'avg' over tokens 1.., LN after pooling, then @ proj.
    Stock cfg: ln_pre, ln_post over all tokens, 'tok' pooling."""
    x = patch_embed(images, sd[prefix + "conv1.weight"])
    B = x.shape[0]
    cls = sd[prefix + "class_embedding"].expand(B, 1, -1)
    x = torch.cat([cls, x], dim=1) + sd[prefix + "positional_embedding"]        # :615-617
    if prefix + "ln_pre.weight" in sd:                                           # :620 (Identity when no_ln_pre)
        x = layer_norm(x, sd[prefix + "ln_pre.weight"], sd[prefix + "ln_pre.bias"], eps)
    if taps is not None:
        taps["embed"] = x
    for i in range(_num_layers(sd, prefix + "transformer.")):                   # Transformer.forward :355-366
        x = residual_attention_block(x, sd, f"{prefix}transformer.resblocks.{i}.", heads, eps, act, taps=taps)
        if taps is not None:
            taps[f"block{i}"] = x
    lnw, lnb = sd[prefix + "ln_post.weight"], sd[prefix + "ln_post.bias"]
    if final_ln_after_pool:                                                      # :638-640
        pooled = _global_pool(x, pool_type)
        pooled = layer_norm(pooled, lnw, lnb, eps)
    else:                                                                        # :641-643
        x = layer_norm(x, lnw, lnb, eps)
        pooled = _global_pool(x, pool_type)
    if prefix + "proj" in sd:                                                    # :645-646
        pooled = pooled @ sd[prefix + "proj"]
    return pooled


def _global_pool(x: Tensor, pool_type: str) -> Tensor:
    """transformer.py:599-607."""
    if pool_type == "avg":
        return x[:, 1:].mean(dim=1)
    if pool_type == "tok":
        return x[:, 0]
    raise ValueError(pool_type)


# ------------------------------------------------------------------------------------------------------------------
# text tower + CLIP wrapper
# ------------------------------------------------------------------------------------------------------------------
def text_transformer(text: Tensor, sd: Dict[str, Tensor], heads: int, *, causal: bool, pool_type: str = "last",
                     eps: float = 1e-6, act: str = "tanh") -> Tensor:
    """CLIP.encode_text, model.py:269-284: token_embedding + positional_embedding -> transformer (additive causal
    mask unless no_causal_mask) -> ln_final over all tokens -> text_global_pool (transformer.py:654-666) -> projection."""
    x = sd["token_embedding.weight"][text] + sd["positional_embedding"][: text.shape[1]]
    L = x.shape[1]
    mask = None
    if causal:
        mask = torch.full((L, L), float("-inf"), dtype=x.dtype).triu_(1)
    for i in range(_num_layers(sd, "transformer.")):
        x = residual_attention_block(x, sd, f"transformer.resblocks.{i}.", heads, eps, act, attn_mask=mask)
    x = layer_norm(x, sd["ln_final.weight"], sd["ln_final.bias"], eps)
    if pool_type == "last":
        pooled = x[:, -1]
    elif pool_type == "first":
        pooled = x[:, 0]
    elif pool_type == "argmax":
        pooled = x[torch.arange(x.shape[0]), text.argmax(dim=-1)]
    else:
        raise ValueError(pool_type)
    return pooled @ sd["text_projection"]


def l2_normalize(x: Tensor, eps: float = 1e-12) -> Tensor:
    """model.py:267,284 F.normalize(x, dim=-1): x / max(||x||_2, eps)."""
    return x / x.norm(dim=-1, keepdim=True).clamp_min(eps)


# ------------------------------------------------------------------------------------------------------------------
# contrastive loss
# ------------------------------------------------------------------------------------------------------------------
def clip_loss(img: Tensor, txt: Tensor, logit_scale: Tensor) -> Tensor:
    """ClipLoss.forward, loss.py:102-131 with world_size == 1:
    z = s * I T^T ; L = (CE(z, arange) + CE(z^T, arange)) / 2."""
    z = logit_scale * img @ txt.t()
    n = z.shape[0]
    idx = torch.arange(n)
    row = torch.logsumexp(z, dim=1) - z[idx, idx]
    col = torch.logsumexp(z, dim=0) - z[idx, idx]
    return 0.5 * (row.mean() + col.mean())


def dual_caption_loss(img: Tensor, txt1: Tensor, txt2: Tensor, logit_scale: Tensor) -> Tensor:
    """src/losses/common.py:139-171 on one device (rank 0 of 1): log-softmax of img . txt_k^T * t along the rows for both
    directions and both caption sets, diagonal picked, local_loss_k = 0.5 (l_img_k + l_txt_k), mean((loss_1 + loss_2) / 2)."""
    def one(txt):
        zi = torch.log_softmax(logit_scale * img @ txt.t(), dim=1)
        zt = torch.log_softmax(logit_scale * txt @ img.t(), dim=1)
        return 0.5 * (-torch.diagonal(zi) - torch.diagonal(zt))
    return ((one(txt1) + one(txt2)) / 2).mean()


def clip_loss_grads(img: Tensor, txt: Tensor, logit_scale: Tensor):
    """Closed-form gradients of clip_loss: G = (P_row + P_col - 2 I) / (2N),
    dI = s G T, dT = s G^T I, d(scale) = sum(G * z) / s."""
    z = logit_scale * img @ txt.t()
    n = z.shape[0]
    p_row = torch.softmax(z, dim=1)
    p_col = torch.softmax(z, dim=0)
    G = (p_row + p_col - 2.0 * torch.eye(n, dtype=z.dtype)) / (2.0 * n)
    return logit_scale * G @ txt, logit_scale * G.t() @ img, (G * z).sum() / logit_scale


def clip_loss_local(img_all: Tensor, txt_all: Tensor, logit_scale: Tensor, rank: int, world_size: int) -> Tensor:
    """ClipLoss with world_size > 1, local_loss=True (loss.py:89-110): rank r scores its own rows against the gathered
    features; labels are arange(n_loc) + n_loc * rank."""
    n = img_all.shape[0] // world_size
    sl = slice(rank * n, (rank + 1) * n)
    zi = logit_scale * img_all[sl] @ txt_all.t()
    zt = logit_scale * txt_all[sl] @ img_all.t()
    lab = torch.arange(n) + n * rank
    li = torch.logsumexp(zi, dim=1) - zi[torch.arange(n), lab]
    lt = torch.logsumexp(zt, dim=1) - zt[torch.arange(n), lab]
    return 0.5 * (li.mean() + lt.mean())

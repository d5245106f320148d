"""Deterministic synthetic parameters / inputs (test infrastructure).

numpy RandomState streams are stable across numpy versions and machines, so the same seed gives the same weights in
the build container (where the unmodified reference is imported to make the golden vectors) and on the GPU box.
Key names and shapes are the reference's state_dict contract (open_clip/transformer.py:437-540,
model.py:223-254).
"""
from __future__ import annotations

import numpy as np
import torch

# named configs. vision cfg keys are CLIPVisionCfg fields (model.py:27-55).
OPENVISION_FLAGS = dict(no_ln_pre=True, pool_type="avg", final_ln_after_pool=True)
STOCK_FLAGS = dict(no_ln_pre=False, pool_type="tok", final_ln_after_pool=False)

CONFIGS = {
    # tiny shapes for fast CPU tests (head_width stays 64, the only head size the image tower uses below H/14)
    "mini-ov": dict(embed_dim=64, vision=dict(image_size=48, patch_size=16, width=128, layers=2, head_width=64, **OPENVISION_FLAGS),
                    text=dict(context_length=8, vocab_size=64, width=64, heads=1, layers=2, no_causal_mask=True,
                              pool_type="last", act_kwargs={"approximate": "tanh"})),
    "mini-stock": dict(embed_dim=64, vision=dict(image_size=56, patch_size=14, width=128, layers=2, head_width=64, **STOCK_FLAGS),
                       text=dict(context_length=8, vocab_size=64, width=64, heads=1, layers=2, pool_type="last")),
    # H/14-style head width (80) at toy size: 2 heads x 80
    "mini-h80": dict(embed_dim=64, vision=dict(image_size=56, patch_size=14, width=160, layers=2, head_width=80, **OPENVISION_FLAGS),
                     text=dict(context_length=8, vocab_size=64, width=64, heads=1, layers=2, no_causal_mask=True,
                               pool_type="last", act_kwargs={"approximate": "tanh"})),
    "Ti16-160": dict(embed_dim=192, vision=dict(image_size=160, patch_size=16, width=192, layers=12, head_width=64, **OPENVISION_FLAGS),
                     text=dict(context_length=80, vocab_size=32000, width=192, heads=3, layers=12, no_causal_mask=True,
                               pool_type="last", act_kwargs={"approximate": "tanh"})),
    "B16-384": dict(embed_dim=512, vision=dict(image_size=384, patch_size=16, width=768, layers=12, head_width=64, **OPENVISION_FLAGS),
                    text=dict(context_length=80, vocab_size=32000, width=512, heads=8, layers=12, no_causal_mask=True,
                              pool_type="last", act_kwargs={"approximate": "tanh"})),
    "L14-224": dict(embed_dim=768, vision=dict(image_size=224, patch_size=14, width=1024, layers=24, head_width=64, **OPENVISION_FLAGS),
                    text=dict(context_length=80, vocab_size=32000, width=768, heads=12, layers=12, no_causal_mask=True,
                              pool_type="last", act_kwargs={"approximate": "tanh"})),
    "H14-224": dict(embed_dim=1024, vision=dict(image_size=224, patch_size=14, width=1280, layers=32, head_width=80, **OPENVISION_FLAGS),
                    text=dict(context_length=80, vocab_size=32000, width=1024, heads=16, layers=24, no_causal_mask=True,
                              pool_type="last", act_kwargs={"approximate": "tanh"})),
}


def _block_shapes(prefix: str, width: int, mlp_width: int):
    return {
        prefix + "ln_1.weight": (width,), prefix + "ln_1.bias": (width,),
        prefix + "attn.in_proj_weight": (3 * width, width), prefix + "attn.in_proj_bias": (3 * width,),
        prefix + "attn.out_proj.weight": (width, width), prefix + "attn.out_proj.bias": (width,),
        prefix + "ln_2.weight": (width,), prefix + "ln_2.bias": (width,),
        prefix + "mlp.c_fc.weight": (mlp_width, width), prefix + "mlp.c_fc.bias": (mlp_width,),
        prefix + "mlp.c_proj.weight": (width, mlp_width), prefix + "mlp.c_proj.bias": (width,),
    }


def vision_shapes(vcfg: dict, embed_dim: int, prefix: str = "visual."):
    w, p = vcfg["width"], vcfg["patch_size"]
    g = vcfg["image_size"] // p
    mlp = int(w * vcfg.get("mlp_ratio", 4.0))
    shapes = {
        prefix + "class_embedding": (w,),
        prefix + "positional_embedding": (g * g + 1, w),
        prefix + "proj": (w, embed_dim),
        prefix + "conv1.weight": (w, 3, p, p),
    }
    if not vcfg.get("no_ln_pre", False):
        shapes[prefix + "ln_pre.weight"] = (w,)
        shapes[prefix + "ln_pre.bias"] = (w,)
    for i in range(vcfg["layers"]):
        shapes.update(_block_shapes(f"{prefix}transformer.resblocks.{i}.", w, mlp))
    shapes[prefix + "ln_post.weight"] = (w,)
    shapes[prefix + "ln_post.bias"] = (w,)
    return shapes


def text_shapes(tcfg: dict, embed_dim: int):
    w = tcfg["width"]
    mlp = int(w * tcfg.get("mlp_ratio", 4.0))
    shapes = {
        "positional_embedding": (tcfg["context_length"], w),
        "text_projection": (w, embed_dim),
        "logit_scale": (),
        "token_embedding.weight": (tcfg["vocab_size"], w),
    }
    for i in range(tcfg["layers"]):
        shapes.update(_block_shapes(f"transformer.resblocks.{i}.", w, mlp))
    shapes["ln_final.weight"] = (w,)
    shapes["ln_final.bias"] = (w,)
    return shapes


def _fill(shapes: dict, seed: int, dtype=torch.float32):
    rs = np.random.RandomState(seed)
    sd = {}
    for name in sorted(shapes):
        shape = shapes[name]
        if name == "logit_scale":
            v = np.full(shape, np.log(1 / 0.07))
        elif name.endswith(("ln_1.weight", "ln_2.weight", "ln_pre.weight", "ln_post.weight", "ln_final.weight")):
            v = 1.0 + 0.1 * rs.standard_normal(shape)
        elif name.endswith(".bias") or name.endswith("in_proj_bias"):
            v = 0.1 * rs.standard_normal(shape)
        elif name.endswith("conv1.weight"):
            v = rs.standard_normal(shape) * (shape[1] * shape[2] * shape[3]) ** -0.5
        elif name.endswith(("class_embedding", "positional_embedding")):
            v = rs.standard_normal(shape) * 0.3
        elif name.endswith("token_embedding.weight"):
            v = rs.standard_normal(shape) * 0.5
        else:  # matrices: fan-in scaled, slightly hot so attention / GELU are exercised away from zero
            fan_in = shape[1] if not name.endswith(("proj", "text_projection")) else shape[0]
            v = rs.standard_normal(shape) * 1.5 * fan_in ** -0.5
        sd[name] = torch.from_numpy(np.asarray(v, dtype=np.float64)).to(dtype)
    return sd


def make_state_dict(cfg_name: str, seed: int = 0, dtype=torch.float32, vision_only: bool = False):
    cfg = CONFIGS[cfg_name]
    shapes = vision_shapes(cfg["vision"], cfg["embed_dim"])
    if not vision_only:
        shapes.update(text_shapes(cfg["text"], cfg["embed_dim"]))
    return _fill(shapes, seed, dtype)


def make_images(cfg_name: str, batch: int, seed: int = 0, dtype=torch.float32):
    s = CONFIGS[cfg_name]["vision"]["image_size"]
    rs = np.random.RandomState(1000 + seed)
    return torch.from_numpy(rs.standard_normal((batch, 3, s, s))).to(dtype)


def make_text(cfg_name: str, batch: int, seed: int = 0):
    t = CONFIGS[cfg_name]["text"]
    rs = np.random.RandomState(2000 + seed)
    return torch.from_numpy(rs.randint(1, t["vocab_size"], size=(batch, t["context_length"]))).long()


def make_features(n: int, e: int, seed: int = 0, dtype=torch.float32):
    """L2-normalised random image / text embeddings with a correlated diagonal (so the loss is not just ln N)."""
    rs = np.random.RandomState(3000 + seed)
    a = rs.standard_normal((n, e))
    b = 0.6 * a + 0.8 * rs.standard_normal((n, e))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    return torch.from_numpy(a).to(dtype), torch.from_numpy(b).to(dtype)


def vision_heads(cfg_name: str) -> int:
    v = CONFIGS[cfg_name]["vision"]
    return v["width"] // v["head_width"]

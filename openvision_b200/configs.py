"""Named model configurations of the OpenVision family (the reference keeps them as
open_clip/model_configs/*.json and src/configs/openvision.py:260, transfer_jax2hf.py:76-83).  `vision` / `text` are the
keyword arguments of CLIPVisionCfg / CLIPTextCfg (model.py:27-84).

OpenVision towers: no pre-LayerNorm, average pool over the patch tokens, LayerNorm after the pool, tanh-GELU non-causal text
tower with last-token pooling; the stock open_clip layout (ln_pre, cls-token pool) is kept for the mini test config.
"""
from __future__ import annotations

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


def vision_heads(cfg_name: str) -> int:
    v = CONFIGS[cfg_name]["vision"]
    return v["width"] // v["head_width"]

"""CPU: host logic of the drop-in module surface — constructor arguments, attribute names and state_dict contract of
the reference, loud failure without a GPU, and that the product never imports the oracle."""
import os
import re

import pytest
import torch
from torch import nn

import openvision_b200 as ovb
from openvision_b200._lib import OvkError
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(cfg_name):
    cfg = synth.CONFIGS[cfg_name]
    return ovb.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]))


@pytest.mark.parametrize("cfg_name", ["mini-ov", "mini-stock", "Ti16-160"])
def test_state_dict_contract(cfg_name):
    cfg = synth.CONFIGS[cfg_name]
    model = build(cfg_name)
    want = dict(synth.vision_shapes(cfg["vision"], cfg["embed_dim"]))
    want.update(synth.text_shapes(cfg["text"], cfg["embed_dim"]))
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}
    # loads the synthetic checkpoint with strict key matching, as a released OpenVision checkpoint would
    model.load_state_dict(synth.make_state_dict(cfg_name, 0), strict=True)


def test_attributes_the_scripts_touch():
    m = build("mini-ov")
    v = m.visual
    for a in ("conv1", "class_embedding", "positional_embedding", "patch_dropout", "ln_pre", "transformer", "attn_pool",
              "ln_post", "proj", "image_size", "patch_size", "grid_size", "pool_type", "final_ln_after_pool",
              "output_dim", "output_tokens"):
        assert hasattr(v, a), a
    assert isinstance(v.ln_pre, nn.Identity) and v.attn_pool is None
    blk = v.transformer.resblocks[-1]
    for a in ("ln_1", "attn", "ls_1", "ln_2", "mlp", "ls_2"):
        assert hasattr(blk, a)
    assert isinstance(blk.attn, nn.MultiheadAttention)        # convert_weights_to_lp contract, model.py:405-409
    assert isinstance(blk.mlp.gelu, nn.GELU)                  # ov-feature-visualization hooks isinstance(m, GELU)
    assert blk.mlp.c_proj.in_features == 512                  # ov-feature-visualization.py:150-155
    assert len(v.transformer.resblocks) == 2
    assert m.logit_scale.shape == () and abs(float(m.logit_scale) - 2.659260) < 1e-5
    assert v.transformer.get_cast_dtype() == torch.float32
    for meth in ("encode_image", "encode_text", "get_logits", "forward", "lock_image_tower", "set_grad_checkpointing"):
        assert callable(getattr(m, meth))


def test_convert_weights_to_lp_keeps_layernorm_fp32():
    m = build("mini-ov")
    ovb.convert_weights_to_lp(m, torch.bfloat16)
    blk = m.visual.transformer.resblocks[0]
    assert blk.attn.in_proj_weight.dtype == torch.bfloat16 and blk.mlp.c_fc.weight.dtype == torch.bfloat16
    assert m.visual.proj.dtype == torch.bfloat16 and m.visual.conv1.weight.dtype == torch.bfloat16
    assert blk.ln_1.weight.dtype == torch.float32


def test_cpu_tensors_are_rejected_not_silently_computed():
    m = build("mini-ov").eval()
    with pytest.raises(OvkError):
        m.encode_image(torch.zeros(1, 3, 48, 48))
    with pytest.raises(OvkError):
        m.visual.ln_post(torch.zeros(2, 128))
    with pytest.raises(OvkError):
        m.visual.transformer(torch.zeros(1, 10, 128))


def test_lock_image_tower_and_gelu_hook_detection():
    m = build("mini-ov")
    m.lock_image_tower()
    assert all(not p.requires_grad for p in m.visual.parameters())
    blk = m.visual.transformer.resblocks[0]
    assert blk._fusable()
    h = blk.mlp.gelu.register_forward_hook(lambda mod, i, o: None)
    assert not blk._fusable()          # hooks on nn.GELU force the module-by-module path so the hook fires
    h.remove()
    assert blk._fusable()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "openvision_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
                assert "vit_oracle" not in src, f"{f} references the oracle"


def test_product_configs_are_the_oracle_configs():
    """bench.py's GPU arm takes its model configurations from the product package (no oracle import on that arm)."""
    from openvision_b200.configs import CONFIGS
    assert CONFIGS == synth.CONFIGS


def test_hooks_anywhere_in_a_block_disable_the_fused_path():
    m = build("mini-ov")
    blk = m.visual.transformer.resblocks[0]
    for mod in (blk.ln_1, blk.attn, blk.attn.out_proj, blk.mlp, blk.mlp.c_fc, blk.mlp.c_proj, blk.ln_2):
        assert blk._fusable()
        for reg in (mod.register_forward_hook, mod.register_forward_pre_hook, mod.register_full_backward_hook):
            h = reg(lambda *a: None)
            assert not blk._fusable(), (type(mod).__name__, reg.__name__)
            h.remove()
    h = blk.register_forward_hook(lambda *a: None)       # a hook on the block itself fires in __call__: fusing stays legal
    assert blk._fusable()
    from openvision_b200.transformer import _own_hooks
    assert _own_hooks(blk)
    h.remove()
    g = torch.nn.modules.module.register_module_forward_hook(lambda *a: None)
    assert not blk._fusable()
    g.remove()
    assert blk._fusable()


def test_decay_mask_is_the_reference_kernel_rule():
    """build_optax.py:257 wd_mults=[('.*/kernel$', 1.0)]: Dense / conv kernels decay; biases, LayerNorm scales, cls / pos
    embeddings, the temperature and the vocabulary embedding (nn.Embed 'embedding', text_transformer.py:633) do not."""
    from openvision_b200.optim import decay_mask_from_modules, default_decay_mask
    m = build("mini-ov")
    named = dict(m.named_parameters())
    by_type = decay_mask_from_modules(m)
    decayed = {n for n, p in named.items() if default_decay_mask(n, p)}
    assert decayed == {n for n, p in named.items() if by_type(n, p)}
    assert "token_embedding.weight" not in decayed
    assert "positional_embedding" not in decayed and "visual.positional_embedding" not in decayed
    assert "visual.class_embedding" not in decayed and "logit_scale" not in decayed
    assert not any(n.endswith(".bias") or "ln_" in n for n in decayed)
    want = {"visual.conv1.weight", "visual.proj", "text_projection"}
    for pre, nl in (("visual.transformer.resblocks.", 2), ("transformer.resblocks.", 2)):
        for i in range(nl):
            want |= {f"{pre}{i}.attn.in_proj_weight", f"{pre}{i}.attn.out_proj.weight", f"{pre}{i}.mlp.c_fc.weight",
                     f"{pre}{i}.mlp.c_proj.weight"}
    assert decayed == want


def test_resize_pos_embed_matches_reference_golden(golden):
    """checkpoint interchange (reference model.py:523-592): positional-embedding resampling at load
    time against tables the reference's own resize_pos_embed / resize_text_pos_embed produced (oracle/make_golden.py)."""
    import numpy as np
    import torch
    import openvision_b200 as ovb
    from oracle.make_golden import POS_CASES, TEXT_POS_CASES, _Grid
    g = golden("pos_embed_resize.npz")
    for (og, ng, mode, aa) in POS_CASES:
        gen = torch.Generator().manual_seed(1000 + og * 37 + ng)
        table = torch.randn(og * og + 1, 48, generator=gen, dtype=torch.float64)
        sd = {"visual.positional_embedding": table.clone()}
        ovb.resize_pos_embed(sd, _Grid((ng, ng), 8, 48), interpolation=mode, antialias=aa)
        ref = g[f"img_{og}_{ng}_{mode}_{int(aa)}"]
        got = sd["visual.positional_embedding"].numpy()
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 1e-12, (og, ng, mode, aa)
        assert np.array_equal(got[0], table[0].numpy()), "class-token row must pass through"
    for (oc, nc, mode, aa) in TEXT_POS_CASES:
        gen = torch.Generator().manual_seed(2000 + oc * 37 + nc)
        table = torch.randn(oc, 40, generator=gen, dtype=torch.float64)
        sd = {"positional_embedding": table.clone()}
        ovb.resize_text_pos_embed(sd, _Grid((4, 4), nc, 40), interpolation=mode, antialias=aa)
        assert np.abs(sd["positional_embedding"].numpy() - g[f"txt_{oc}_{nc}_{mode}_{int(aa)}"]).max() < 1e-12


def test_resize_pos_embed_then_strict_load():
    """a checkpoint trained at another resolution loads strictly after the resize (factory.py:178-179 order)."""
    import torch
    import openvision_b200 as ovb
    from oracle import synth
    cfg = synth.CONFIGS["mini-ov"]                      # 48 px / patch 16: 3 x 3 grid
    model = ovb.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]))
    sd = synth.make_state_dict("mini-ov", seed=0)
    width = sd["visual.positional_embedding"].shape[1]
    sd["visual.positional_embedding"] = torch.randn(5 * 5 + 1, width)      # as if trained at 80 px
    sd["positional_embedding"] = torch.randn(12, sd["positional_embedding"].shape[1])
    ovb.resize_pos_embed(sd, model)
    ovb.resize_text_pos_embed(sd, model)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not missing and not unexpected
    assert tuple(model.visual.positional_embedding.shape) == (3 * 3 + 1, width)

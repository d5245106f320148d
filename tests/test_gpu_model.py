"""GPU parity of the drop-in module surface against (a) the committed outputs of the unmodified reference
(tests/golden) and (b) the CPU oracle on the same seeded inputs.  Bar (north_star): L2-normalised embeddings
cosine >= 0.9995 and max-abs <= 2e-2 in bf16 against fp32; loss within 1e-3 relative."""
import numpy as np
import pytest
import torch

import openvision_b200 as ovb
from oracle import synth, vit_oracle as O

pytestmark = pytest.mark.gpu


def build(cfg_name):
    cfg = synth.CONFIGS[cfg_name]
    m = ovb.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]))
    m.load_state_dict(synth.make_state_dict(cfg_name, 0), strict=True)
    return m.cuda().eval()


def check_embeddings(got, ref, what):
    got = got.float().cpu()
    ref = torch.as_tensor(ref).float()
    err = (got - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=-1).min().item()
    assert err <= 2e-2 and cos >= 0.9995, f"{what}: max-abs {err:.3e}, min cosine {cos:.6f}"


@pytest.mark.parametrize("cfg_name,batch", [("mini-ov", 4), ("mini-stock", 4), ("Ti16-160", 8)])
def test_encode_image_and_text_match_reference_golden(golden, cfg_name, batch):
    g = golden(f"tower_{cfg_name}.npz")
    m = build(cfg_name)
    images = synth.make_images(cfg_name, batch, 0).cuda()
    text = synth.make_text(cfg_name, batch, 0).cuda()
    with torch.no_grad():
        img_n = m.encode_image(images, normalize=True)
        raw = m.encode_image(images)
        check_embeddings(img_n, g["image_features_norm"], f"{cfg_name} image")
        ref_raw = torch.as_tensor(g["image_features"]).float()
        assert (raw.float().cpu() - ref_raw).abs().max().item() <= 3e-2 * ref_raw.abs().max().item()
        # (mini-stock: the stock text tower with its causal mask, transformer.py:757-763 — the causal flag of the attention kernels)
        txt_n = m.encode_text(text, normalize=True)
        check_embeddings(txt_n, g["text_features_norm"], f"{cfg_name} text")
        i2, t2, s = m(images, text)
        assert abs(float(s) - float(g["logit_scale_exp"])) < 1e-4
        check_embeddings(i2, g["image_features_norm"], "forward image")
        check_embeddings(t2, g["text_features_norm"], "forward text")


def test_piecewise_calls_like_ov_zero_shot(golden):
    """ov-zero-shot-test.py:103-126 calls conv1 / ln_pre / transformer / ln_post / proj one by one."""
    cfg_name, batch = "Ti16-160", 8
    g = golden(f"tower_{cfg_name}.npz")
    m = build(cfg_name)
    v = m.visual
    images = synth.make_images(cfg_name, batch, 0).cuda()
    with torch.no_grad():
        x = v.conv1(images)
        assert x.shape == (batch, 192, 10, 10)
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
        x = torch.cat([v.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device), x], dim=1)
        x = x + v.positional_embedding.to(x.dtype)
        x = v.ln_pre(x)
        x = v.transformer(x)
        pooled = x[:, 1:].mean(dim=1)
        pooled = v.ln_post(pooled)
        feats = pooled @ v.proj
    feats = torch.nn.functional.normalize(feats.float(), dim=-1)
    check_embeddings(feats, g["image_features_norm"], "piecewise tower")


def test_gelu_forward_hooks_fire_with_reference_activations(golden):
    """cliptoolsoptimized.py:1149-1164 hangs forward hooks on every nn.GELU; outputs must be the [B, L, mlp] tensors."""
    cfg_name, batch = "mini-ov", 4
    g = golden(f"tower_{cfg_name}.npz")
    m = build(cfg_name)
    taps = {}
    hooks = [blk.mlp.gelu.register_forward_hook(lambda mod, i, o, k=k: taps.__setitem__(k, o.detach()))
             for k, blk in enumerate(m.visual.transformer.resblocks)]
    with torch.no_grad():
        out_hooked = m.encode_image(synth.make_images(cfg_name, batch, 0).cuda(), normalize=True)
    for h in hooks:
        h.remove()
    assert sorted(taps) == [0, 1]
    ref0 = torch.as_tensor(g["gelu_block0"]).float()
    assert tuple(taps[0].shape) == tuple(ref0.shape)
    err = (taps[0].float().cpu() - ref0).abs().max().item()
    assert err <= 2e-2 * ref0.abs().max().item(), err
    check_embeddings(out_hooked, g["image_features_norm"], "hooked path")


@pytest.mark.parametrize("prec", ["fp32", "bf16", "pure_bf16", "amp_fp16"])
def test_precision_modes_of_the_surface(golden, prec):
    """factory.py:275-297 precision modes: fp32 weights, 'bf16' (convert_weights_to_lp), pure bf16, fp16 autocast."""
    cfg_name, batch = "mini-ov", 4
    g = golden(f"tower_{cfg_name}.npz")
    m = build(cfg_name)
    images = synth.make_images(cfg_name, batch, 0).cuda()
    with torch.no_grad():
        if prec == "bf16":
            ovb.convert_weights_to_lp(m, torch.bfloat16)
            out = m.encode_image(images.bfloat16(), normalize=True)
        elif prec == "pure_bf16":
            m = m.to(torch.bfloat16)
            out = m.encode_image(images.bfloat16(), normalize=True)
        elif prec == "amp_fp16":
            with torch.autocast("cuda", dtype=torch.float16):
                out = m.encode_image(images, normalize=True)
        else:
            out = m.encode_image(images, normalize=True)
    check_embeddings(out, g["image_features_norm"], prec)


def test_batch_shard_independence_full_config():
    """Size-independent property at a real config (B/16@384, 577 tokens): encoding a batch equals encoding its shards
    (the data-parallel split of SURVEY.md §8e) bit-exactly."""
    cfg = synth.CONFIGS["B16-384"]
    torch.manual_seed(0)
    v = ovb.model._build_vision_tower(cfg["embed_dim"], cfg["vision"]).cuda().eval()
    images = torch.randn(6, 3, 384, 384, device="cuda")
    with torch.no_grad():
        full = v(images)
        parts = torch.cat([v(images[:2]), v(images[2:])])
    assert torch.isfinite(full).all()
    assert torch.equal(full, parts)


def test_headline_config_shards_and_kernel_variants_agree():
    """ViT-L/14@224 (the benchmark's config; too big for a CPU reference at this batch): (1) encoding a batch equals encoding
    its shards bit-exactly; (2) the alternative kernels of the path — stand-alone LayerNorm instead of the LayerNorm folded
    into the GEMMs, the one-tile attention kernel instead of the pair kernel — give the same normalised embeddings to bf16
    rounding (cosine >= 0.9995, the north-star bar, against each other)."""
    import os
    cfg = synth.CONFIGS["L14-224"]
    torch.manual_seed(0)
    v = ovb.model._build_vision_tower(cfg["embed_dim"], cfg["vision"]).cuda().eval()
    images = torch.randn(48, 3, 224, 224, device="cuda")
    saved = {k: os.environ.get(k) for k in ("OVK_LN_FOLD", "OVK_ATT_V1")}
    try:
        with torch.no_grad():
            full = v(images)
            parts = torch.cat([v(images[:16]), v(images[16:])])
            assert torch.isfinite(full).all() and torch.equal(full, parts)
            ref = torch.nn.functional.normalize(full.float(), dim=-1)
            for var in ("OVK_LN_FOLD=0", "OVK_ATT_V1=1"):
                k, val = var.split("=")
                os.environ[k] = val
                alt = torch.nn.functional.normalize(v(images).float(), dim=-1)
                os.environ.pop(k)
                cos = (ref * alt).sum(-1)
                assert cos.min().item() >= 0.9995, (var, cos.min().item())
                assert (ref - alt).abs().max().item() <= 2e-2, (var, (ref - alt).abs().max().item())
    finally:
        for k, val in saved.items():
            if val is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = val


@pytest.mark.parametrize("cfg_name,batch", [("L14-224", 4), ("B16-384", 2), ("H14-224", 2)])
def test_full_size_towers_match_the_fp32_oracle(cfg_name, batch):
    """The benchmark's own towers (24 x 1024 @ 257 tokens; 12 x 768 @ 577 tokens) against the fp32 CPU oracle on a few
    images: normalised embeddings within the north-star bar.  Exercises the production kernel selection (LayerNorm folded
    into the GEMMs, CTA-pair GEMMs, pair / one-tile attention kernels, remainder-token handling at L = 257)."""
    sd = synth.make_state_dict(cfg_name, 0, vision_only=True)
    images = synth.make_images(cfg_name, batch, 0)
    with torch.no_grad():
        ref = O.l2_normalize(O.vision_transformer(images, sd, synth.vision_heads(cfg_name), pool_type="avg",
                                                  final_ln_after_pool=True))
    cfg = synth.CONFIGS[cfg_name]
    v = ovb.model._build_vision_tower(cfg["embed_dim"], cfg["vision"])
    v.load_state_dict({k[len("visual."):]: t for k, t in sd.items() if k.startswith("visual.")}, strict=True)
    v = v.cuda().eval()
    with torch.no_grad():
        # a batch big enough for the CTA-pair GEMMs (>= 512 rows): the checked images first, fillers after
        filler = torch.randn(8, *images.shape[1:], generator=torch.Generator().manual_seed(5))
        got = torch.nn.functional.normalize(v(torch.cat([images, filler]).cuda()).float(), dim=-1)[:batch]
    check_embeddings(got, ref, f"{cfg_name} vs fp32 oracle")

"""GPU: the parts of the drop-in boundary beyond plain forward parity — forward / backward hooks on ANY
sub-module (cliptoolsoptimized.py:480-489 hooks arbitrary modules), ClipLoss.get_logits / CLIP.get_logits as callables
(loss.py:102-118, 165, 195-201; model.py:286-293), tensors on the wrong device."""
import numpy as np
import pytest
import torch

import openvision_b200 as ovb
from openvision_b200._lib import OvkError
from oracle import synth, vit_oracle as O

pytestmark = pytest.mark.gpu


def build(cfg_name):
    cfg = synth.CONFIGS[cfg_name]
    m = ovb.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]))
    m.load_state_dict(synth.make_state_dict(cfg_name, 0), strict=True)
    return m.cuda().eval()


def _targets(m):
    v = m.visual
    b0, b1 = v.transformer.resblocks[0], v.transformer.resblocks[1]
    return {"resblock0": b0, "resblock1.ln_1": b1.ln_1, "resblock0.attn": b0.attn, "resblock1.mlp.c_fc": b1.mlp.c_fc,
            "resblock0.mlp.c_proj": b0.mlp.c_proj, "resblock1.ln_2": b1.ln_2, "resblock0.mlp": b0.mlp,
            "transformer": v.transformer, "ln_post": v.ln_post, "conv1": v.conv1, "visual": v}


@pytest.mark.parametrize("cfg_name", ["mini-ov", "mini-stock"])
def test_forward_hooks_on_any_submodule_fire_and_do_not_change_the_result(cfg_name):
    m = build(cfg_name)
    images = synth.make_images(cfg_name, 4, 0).cuda()
    with torch.no_grad():
        base = m.encode_image(images, normalize=True).float()
    cfg = synth.CONFIGS[cfg_name]["vision"]
    L = (cfg["image_size"] // cfg["patch_size"]) ** 2 + 1
    D = cfg["width"]
    want_shape = {"resblock0": (4, L, D), "resblock1.ln_1": (4, L, D), "resblock1.mlp.c_fc": (4, L, 4 * D),
                  "resblock0.mlp.c_proj": (4, L, D), "resblock1.ln_2": (4, L, D), "resblock0.mlp": (4, L, D),
                  "transformer": (4, L, D), "conv1": (4, D, cfg["image_size"] // cfg["patch_size"], cfg["image_size"] // cfg["patch_size"]),
                  "visual": (4, synth.CONFIGS[cfg_name]["embed_dim"])}
    for name, mod in _targets(m).items():
        seen = []
        h = mod.register_forward_hook(lambda mod_, inp, out, seen=seen: seen.append(out[0] if isinstance(out, tuple) else out))
        with torch.no_grad():
            got = m.encode_image(images, normalize=True).float()
        h.remove()
        assert len(seen) == 1, f"forward hook on {name} fired {len(seen)} times"
        if name in want_shape:
            assert tuple(seen[0].shape) == want_shape[name], (name, tuple(seen[0].shape))
        err = (got - base).abs().max().item()
        assert err <= 1e-2, f"hook on {name} changed the embeddings by {err:.3e}"
    with torch.no_grad():   # all hooks removed: the fused path is back and agrees with itself
        again = m.encode_image(images, normalize=True).float()
    assert torch.equal(again, base)


def test_hook_can_replace_a_submodule_output_and_pre_hooks_fire():
    """A forward hook that RETURNS a tensor replaces the module's output (nn.Module contract): the residual stream must
    see the replacement, which the fused epilogues cannot do — the module-by-module path has to."""
    m = build("mini-ov")
    images = synth.make_images("mini-ov", 2, 0).cuda()
    blk = m.visual.transformer.resblocks[0]
    with torch.no_grad():
        base = m.visual.transformer.resblocks[0](torch.zeros(2, 10, 128, device="cuda", dtype=torch.bfloat16))
        h = blk.mlp.register_forward_hook(lambda mod, inp, out: torch.zeros_like(out))
        pre_seen = []
        hp = blk.ln_2.register_forward_pre_hook(lambda mod, inp: pre_seen.append(inp[0].shape))
        x = torch.randn(2, 10, 128, device="cuda").bfloat16()
        y = blk(x)
        h.remove()
        hp.remove()
        # with the MLP branch zeroed the block is x + attn(ln_1(x)) only
        ref = ovb.transformer._residual_add(x, blk.attention(q_x=blk.ln_1(x)))
    assert len(pre_seen) == 1 and tuple(pre_seen[0]) == (2, 10, 128)
    assert (y.float() - ref.float()).abs().max().item() <= 2e-2
    assert base.shape == (2, 10, 128)


def test_backward_hooks_fire_and_gradients_match_the_fused_path():
    cfg_name = "mini-ov"
    m = build(cfg_name)
    images = synth.make_images(cfg_name, 4, 0).cuda().requires_grad_(True)
    # (a weighted sum: the squared norm of normalised embeddings is constant and would make every gradient rounding noise)
    wgt = torch.randn(4, 64, generator=torch.Generator().manual_seed(3)).cuda()
    out = m.encode_image(images, normalize=True)
    (out.float() * wgt).sum().backward()
    g_ref = images.grad.detach().clone()
    w_ref = m.visual.transformer.resblocks[1].mlp.c_fc.weight.grad.detach().clone()
    images.grad = None
    m.zero_grad(set_to_none=True)
    seen = []
    blk = m.visual.transformer.resblocks[1]
    h1 = blk.mlp.c_fc.register_full_backward_hook(lambda mod, gin, gout: seen.append(("c_fc", tuple(gout[0].shape))))
    h2 = blk.register_full_backward_hook(lambda mod, gin, gout: seen.append(("block", tuple(gout[0].shape))))
    out = m.encode_image(images, normalize=True)
    (out.float() * wgt).sum().backward()
    h1.remove()
    h2.remove()
    names = [n for n, _ in seen]
    assert "c_fc" in names and "block" in names, seen
    assert dict(seen)["c_fc"] == (4, 10, 512) and dict(seen)["block"] == (4, 10, 128)
    rel = (images.grad - g_ref).norm() / g_ref.norm()
    assert rel.item() <= 3e-2, rel.item()
    relw = (blk.mlp.c_fc.weight.grad - w_ref).norm() / w_ref.norm()
    assert relw.item() <= 3e-2, relw.item()
    # hooks on the tower's direct children take the line-by-line path (VisionTransformer._forward_modules), with autograd
    images.grad = None
    m.zero_grad(set_to_none=True)
    h3 = m.visual.transformer.register_forward_hook(lambda *a: None)
    out = m.encode_image(images, normalize=True)
    (out.float() * wgt).sum().backward()
    h3.remove()
    rel = (images.grad - g_ref).norm() / g_ref.norm()
    assert rel.item() <= 3e-2, rel.item()


def test_text_tower_hooks():
    m = build("mini-ov")
    text = synth.make_text("mini-ov", 4, 0).cuda()
    with torch.no_grad():
        base = m.encode_text(text, normalize=True).float()
    for mod, shape in ((m.transformer, (4, 8, 64)), (m.ln_final, (4, 8, 64)), (m.transformer.resblocks[0].attn, None)):
        seen = []
        h = mod.register_forward_hook(lambda mod_, i, o, seen=seen: seen.append(o[0] if isinstance(o, tuple) else o))
        with torch.no_grad():
            got = m.encode_text(text, normalize=True).float()
        h.remove()
        assert len(seen) == 1
        if shape is not None:
            assert tuple(seen[0].shape) == shape
        assert (got - base).abs().max().item() <= 1e-2


@pytest.mark.parametrize("n,e,scale", [(64, 32, 14.2857), (200, 48, 30.0), (1000, 768, 100.0)])
def test_cliploss_get_logits_is_callable_and_differentiable(n, e, scale):
    """loss.py:102-118: logits_per_image = s I T^T, logits_per_text = s T I^T (fp32), with autograd through both."""
    img, txt = synth.make_features(n, e, seed=n + e)
    ib, tb = img.bfloat16().float(), txt.bfloat16().float()          # the GEMM sees bf16-rounded features
    i_g = img.cuda().requires_grad_(True)
    t_g = txt.cuda().requires_grad_(True)
    ls = torch.tensor(float(np.log(scale)), device="cuda", requires_grad=True)
    crit = ovb.ClipLoss()
    lpi, lpt = crit.get_logits(i_g, t_g, ls.exp())
    assert lpi.shape == (n, n) and lpt.shape == (n, n) and lpi.dtype == torch.float32
    ref = scale * ib.double() @ tb.double().t()
    assert (lpi.double().cpu() - ref).abs().max().item() <= 1e-3 * scale
    assert (lpt.double().cpu() - ref.t()).abs().max().item() <= 1e-3 * scale
    # the reference's loss composed from these logits (loss.py:120-131) and its gradients against the oracle
    labels = torch.arange(n, device="cuda")
    loss = 0.5 * (torch.nn.functional.cross_entropy(lpi, labels) + torch.nn.functional.cross_entropy(lpt, labels))
    loss.backward()
    want = float(O.clip_loss(ib, tb, torch.tensor(scale)))
    assert abs(float(loss) - want) <= 1e-3 * abs(want)
    gi, gt, gs = O.clip_loss_grads(ib, tb, torch.tensor(scale))
    assert (i_g.grad.cpu() - gi).norm() <= 3e-2 * gi.norm()
    assert (t_g.grad.cpu() - gt).norm() <= 3e-2 * gt.norm()
    assert abs(float(ls.grad) - float(gs) * scale) <= 3e-2 * abs(float(gs) * scale) + 1e-5


def test_clip_get_logits_matches_forward():
    m = build("mini-ov")
    images = synth.make_images("mini-ov", 4, 0).cuda()
    text = synth.make_text("mini-ov", 4, 0).cuda()
    with torch.no_grad():
        li, lt = m.get_logits(images, text)
        i, t, s = m(images, text)
    ref = float(s) * i.float() @ t.float().t()
    assert li.shape == (4, 4) and torch.equal(lt, li.T)
    assert (li - ref).abs().max().item() <= 2e-2 * float(s)


def test_tensor_on_another_device_is_rejected():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from openvision_b200 import ops
    x = torch.zeros(8, 64, device="cuda:1", dtype=torch.bfloat16)
    g = torch.ones(64, device="cuda:1")
    with pytest.raises(OvkError):
        ops.layernorm(x, g, g, 1e-6)          # current device is cuda:0
    with torch.cuda.device(1):                # per-device kernel attributes: a second GPU in the same process works
        y = ops.layernorm(torch.randn(8, 64, device="cuda:1").bfloat16(), g, torch.zeros(64, device="cuda:1"), 1e-6)
        a = torch.randn(256, 64, device="cuda:1").bfloat16()
        w = torch.randn(128, 64, device="cuda:1").bfloat16()
        z = ops.gemm(a, w)
        torch.cuda.synchronize()
    assert torch.isfinite(y.float()).all()
    assert (z.float() - a.float() @ w.float().t()).abs().max().item() <= 0.5


def test_cuda_graph_replay_matches_eager_bit_for_bit():
    """Small-batch latency path (the ov-* scripts run batch 1-8): encode_image / encode_text captured once into a CUDA graph
    and replayed — the same libovk kernels, so the results are bitwise those of the eager call, for new inputs too."""
    from oracle import synth
    cfg_name = "Ti16-160"
    cfg = synth.CONFIGS[cfg_name]
    m = ovb.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]))
    m.load_state_dict(synth.make_state_dict(cfg_name, 0), strict=True)
    m = m.cuda().eval()
    a = synth.make_images(cfg_name, 2, 0).cuda()
    b = synth.make_images(cfg_name, 2, 1).cuda()
    g = ovb.graphed_encode_image(m, a, normalize=True)
    with torch.no_grad():
        for x in (a, b, a):
            got = g(x).clone()
            ref = m.encode_image(x, normalize=True)
            assert torch.equal(got, ref)
    t = synth.make_text(cfg_name, 2, 0).cuda()
    gt = ovb.graphed_encode_text(m, t, normalize=True)
    with torch.no_grad():
        assert torch.equal(gt(t).clone(), m.encode_text(t, normalize=True))
    with torch.no_grad():   # a changed weight invalidates the captured packed copies: loud, not stale
        m.visual.proj.mul_(1.0)
    with pytest.raises(OvkError):
        g(a)

"""GPU parity of the backward pass: kernel by kernel against autograd of the CPU oracle, and end to end (tower + loss)
against the gradients of the unmodified reference (tests/golden/tower_*.npz).  The whole chain runs in bf16 with fp32
accumulation, so gradients are compared in relative L2 norm (<= 5e-2) and max-abs relative to the largest entry."""
import math

import numpy as np
import pytest
import torch

import openvision_b200 as ovb
from oracle import synth, vit_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from openvision_b200 import ops as _ops
    return _ops


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def rel_l2(got, ref):
    got = torch.as_tensor(got).double().cpu().reshape(-1)
    ref = torch.as_tensor(ref).double().cpu().reshape(-1)
    return float((got - ref).norm() / (ref.norm() + 1e-30))


def assert_grad(got, ref, what, tol=5e-2, abs_floor=1e-6):
    """||got - ref|| <= tol * ||ref|| + abs_floor * sqrt(n)  (the floor covers gradients that are exactly zero)"""
    assert torch.isfinite(torch.as_tensor(got).float()).all(), f"{what}: non-finite gradient"
    g = torch.as_tensor(got).double().cpu().reshape(-1)
    r = torch.as_tensor(ref).double().cpu().reshape(-1)
    err, bound = float((g - r).norm()), tol * float(r.norm()) + abs_floor * math.sqrt(r.numel())
    assert err <= bound, f"{what}: L2 error {err:.3e} > {bound:.3e} (relative {rel_l2(got, ref):.3e})"


@pytest.mark.parametrize("B,L,H", [(1, 128, 1), (2, 64, 2), (3, 101, 3), (2, 257, 4), (1, 577, 2), (1, 16, 1), (2, 129, 1),
                                   (1, 1, 1), (2, 200, 2),
                                   (24, 257, 8), (9, 577, 6)])   # > 148 work items: persistent CTAs walk several items each
def test_attention_bwd(ops, B, L, H):
    hd = 64
    qkv = rnd(B * L, 3 * H * hd, seed=L).bfloat16()
    dout = rnd(B * L, H * hd, seed=L + 1).bfloat16()
    qc = qkv.cuda()
    out, lse = ops.attention(qc, B, L, H, hd, save_lse=True)
    dqkv = ops.attention_bwd(qc, out, dout.cuda(), lse, B, L, H, hd)
    qf = qkv.float().requires_grad_(True)
    q, k, v = qf.view(B, L, 3, H, hd).permute(2, 0, 3, 1, 4)
    p = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), -1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * L, H * hd)
    o.backward(dout.float())
    g = dqkv.float().cpu().view(B, L, 3, H, hd)
    r = qf.grad.view(B, L, 3, H, hd)
    for i, name in enumerate("qkv"):
        assert_grad(g[:, :, i], r[:, :, i], f"attention d{name} B{B} L{L} H{H}", 3e-2)


@pytest.mark.parametrize("hd", [72, 80])
@pytest.mark.parametrize("B,L,H", [(2, 257, 2), (1, 101, 3), (1, 577, 1), (2, 130, 1), (16, 257, 4)])
def test_attention_bwd_wide_heads(ops, hd, B, L, H):
    qkv = rnd(B * L, 3 * H * hd, seed=L + hd).bfloat16()
    dout = rnd(B * L, H * hd, seed=L + hd + 1).bfloat16()
    qc = qkv.cuda()
    out, lse = ops.attention(qc, B, L, H, hd, save_lse=True)
    dqkv = ops.attention_bwd(qc, out, dout.cuda(), lse, B, L, H, hd)
    qf = qkv.float().requires_grad_(True)
    q, k, v = qf.view(B, L, 3, H, hd).permute(2, 0, 3, 1, 4)
    p = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), -1)
    (p @ v).permute(0, 2, 1, 3).reshape(B * L, H * hd).backward(dout.float())
    g = dqkv.float().cpu().view(B, L, 3, H, hd)
    r = qf.grad.view(B, L, 3, H, hd)
    for i, name in enumerate("qkv"):
        assert_grad(g[:, :, i], r[:, :, i], f"attention hd{hd} d{name} B{B} L{L} H{H}", 3e-2)


@pytest.mark.parametrize("B,L,H,hd,causal", [(2, 577, 3, 64, False), (1, 300, 2, 64, True), (3, 77, 2, 64, True), (2, 577, 2, 80, False),
                                             (1, 200, 3, 72, False), (30, 577, 12, 64, False), (20, 640, 8, 80, False),
                                             (1, 1025, 2, 64, False)])
@pytest.mark.parametrize("mode", ["2", "2w16", "1"])
def test_attention_bwd_one_pass_matches_two_pass(ops, B, L, H, hd, causal, mode, monkeypatch):
    """MODE_FUSED (one walk over the score tiles, dQ through an fp32 scratch + TMA reduce-add) against the two-pass kernels
    on the same inputs: dV comes out of the same MMAs (bitwise equal), dK sees a delta summed in another order, dQ an fp32
    summation order of its own;
    and both against autograd of the fp32 composition (transformer.py:225,250-252; causal mask :757-763)."""
    qkv = rnd(B * L, 3 * H * hd, seed=L + hd + 7).bfloat16()
    dout = rnd(B * L, H * hd, seed=L + hd + 8).bfloat16()
    qc, dc = qkv.cuda(), dout.cuda()
    out, lse = ops.attention(qc, B, L, H, hd, save_lse=True, causal=causal)
    if mode == "2w16":   # attention_bwd_t_kernel with 16 instead of 8 compute warps (A/B variant)
        monkeypatch.setenv("OVK_ATTBWD_WARPS", "16")
        mode = "2"
    monkeypatch.setenv("OVK_ATTBWD_FUSED", mode)   # 2: attention_bwd_t_kernel (transposed tiles), 1: attention_bwd_kernel<fused>
    fused = ops.attention_bwd(qc, out, dc, lse, B, L, H, hd, causal=causal)
    monkeypatch.setenv("OVK_ATTBWD_FUSED", "0")
    two = ops.attention_bwd(qc, out, dc, lse, B, L, H, hd, causal=causal)
    gf = fused.float().cpu().view(B, L, 3, H, hd)
    gt = two.float().cpu().view(B, L, 3, H, hd)
    if mode == "1":
        assert torch.equal(gf[:, :, 2], gt[:, :, 2]), "dV: same MMAs in the same order"
    assert_grad(gf[:, :, 2], gt[:, :, 2], "dV one-pass vs two-pass", 2e-3)
    assert_grad(gf[:, :, 1], gt[:, :, 1], "dK one-pass vs two-pass (delta summed in another order)", 2e-3)
    assert_grad(gf[:, :, 0], gt[:, :, 0], "dQ one-pass vs two-pass", 1e-2)
    if B * H <= 16:
        qf = qkv.float().requires_grad_(True)
        q, k, v = qf.view(B, L, 3, H, hd).permute(2, 0, 3, 1, 4)
        sc = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        if causal:
            sc = sc + torch.full((L, L), float("-inf")).triu_(1)
        (torch.softmax(sc, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, H * hd).backward(dout.float())
        r = qf.grad.view(B, L, 3, H, hd)
        for i, name in enumerate("qkv"):
            assert_grad(gf[:, :, i], r[:, :, i], f"one-pass d{name}", 3e-2)


@pytest.mark.parametrize("B,L,H,hd", [(1024, 257, 16, 64), (512, 577, 12, 64), (256, 257, 16, 80)])
def test_attention_bwd_full_size_invariants(ops, B, L, H, hd):
    """BASELINE-size shapes (L/14 and H/14 at 257 tokens, B/16@384 at 577), where an fp32 reference of the whole problem is
    out of reach on the host: two size-independent identities of softmax attention (transformer.py:225,250-252).
      rows of P sum to one            =>  sum_j dV[j] = sum_i dO[i]                     per (image, head)
      f(a q, k / a) does not depend on a  =>  <dQ, Q> = <dK, K>                          per (image, head)
    plus a sampled (image, head) against autograd of the fp32 composition."""
    g = torch.Generator(device="cuda").manual_seed(L + hd)
    qkv = (torch.randn(B * L, 3 * H * hd, device="cuda", generator=g) * 0.7).bfloat16()
    dout = torch.randn(B * L, H * hd, device="cuda", generator=g).bfloat16()
    out, lse = ops.attention(qkv, B, L, H, hd, save_lse=True)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd)
    q5, d5 = qkv.view(B, L, 3, H, hd).float(), dqkv.view(B, L, 3, H, hd).float()
    do4 = dout.view(B, L, H, hd).float()
    assert torch.isfinite(d5).all()
    # sum_j dV_j = sum_i dO_i.  Two roundings: the bf16 outputs (2^-9 relative each) and the bf16 P the tensor core consumes
    # (rows of the rounded P sum to 1 + eps_i, |eps_i| ~ 2^-9, independent: six standard deviations of sum_i eps_i dO_i)
    lhs, rhs = d5[:, :, 2].sum(1), do4.sum(1)
    tol = 2.0 ** -8 * d5[:, :, 2].abs().sum(1) + 6.0 * 2.0 ** -9 * do4.pow(2).sum(1).sqrt() + 1e-3
    assert bool(((lhs - rhs).abs() <= tol).all()), float((lhs - rhs).abs().max())
    # <dQ, Q> = <dK, K>
    a = (d5[:, :, 0] * q5[:, :, 0]).sum((1, 3))
    b = (d5[:, :, 1] * q5[:, :, 1]).sum((1, 3))
    scale_ab = (d5[:, :, 0] * q5[:, :, 0]).abs().sum((1, 3)) + (d5[:, :, 1] * q5[:, :, 1]).abs().sum((1, 3))
    assert bool(((a - b).abs() <= 2.0 ** -7 * scale_ab + 1e-3).all()), float(((a - b).abs() / (scale_ab + 1e-9)).max())
    # one (image, head) in full against fp32 autograd
    bi, hi = B - 1, H // 2
    qf = q5[bi, :, :, hi].cpu().clone().requires_grad_(True)           # [L, 3, hd]
    sc = (qf[:, 0] @ qf[:, 1].T) / math.sqrt(hd)
    (torch.softmax(sc, -1) @ qf[:, 2]).backward(do4[bi, :, hi].cpu())
    for i, name in enumerate("qkv"):
        assert_grad(d5[bi, :, i, hi].cpu(), qf.grad[:, i], f"full-size d{name} of image {bi} head {hi}", 3e-2)


@pytest.mark.parametrize("B,L,H,hd", [(3, 257, 2, 64), (2, 577, 3, 64), (2, 257, 3, 80), (2, 200, 2, 72), (5, 77, 2, 64)])
def test_attention_bwd_one_pass_stays_inside_its_buffers(ops, B, L, H, hd):
    """ovk_attention_bwd_fused through the C ABI with canaries behind the workspace it declares (remainder-token vectors,
    fp32 dQ accumulator, per-query statistics), behind delta and behind dqkv: nothing past the declared sizes is written."""
    from openvision_b200 import _lib
    lib = _lib.load()
    qkv = rnd(B * L, 3 * H * hd, seed=L).bfloat16().cuda()
    dout = rnd(B * L, H * hd, seed=L + 1).bfloat16().cuda()
    out, lse = ops.attention(qkv, B, L, H, hd, save_lse=True)
    n = lib.ovk_attention_bwd_fused_workspace_floats(B, L, H, hd, 0)
    assert n > 0
    pad = 4096
    ws = torch.full((n + pad,), 777.0, dtype=torch.float32, device="cuda")
    delta = torch.full((B * H * L + pad,), 777.0, dtype=torch.float32, device="cuda")
    dq = torch.full((B * L * 3 * H * hd + pad,), 3.0, dtype=torch.bfloat16, device="cuda")
    for flags in (4, 0, 2):   # 8 compute warps (what the Python mirror passes), 16, the first one-pass kernel
        _lib.call("ovk_attention_bwd_fused", qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dq.data_ptr(),
                  delta.data_ptr(), ws.data_ptr(), B, L, H, hd, 1.0 / math.sqrt(hd), flags, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert bool((ws[n:] == 777.0).all()), f"flags {flags}: write past the workspace"
        assert bool((delta[B * H * L:] == 777.0).all()), f"flags {flags}: write past delta"
        assert bool((dq[B * L * 3 * H * hd:] == 3.0).all()), f"flags {flags}: write past dqkv"
        ref = ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd)
        assert_grad(dq[:B * L * 3 * H * hd].view(B * L, 3 * H * hd), ref, f"flags {flags}: C-ABI call vs the Python mirror", 2e-2)


@pytest.mark.parametrize("B,L,H,hd", [(160, 577, 12, 64), (300, 257, 16, 64), (100, 257, 16, 80)])
def test_attention_bwd_one_pass_is_deterministic(ops, B, L, H, hd):
    """Every reduction of the one-pass backward has a fixed order (dQ partial sums: one warp per accumulator block, key tiles
    in order, each TMA reduce-add issued after the previous one COMPLETED; the remainder-token kernel: fixed-order block
    reductions): repeated calls are bitwise identical.  A race in the tile kernel's pipeline would show up here as well
    (several work items per CTA, items crossing (image, head) boundaries); tools/attn_bwd_stress.py is the longer version."""
    qkv = (rnd(B * L, 3 * H * hd, seed=B) * 0.5).bfloat16().cuda()
    dout = rnd(B * L, H * hd, seed=B + 1).bfloat16().cuda()
    out, lse = ops.attention(qkv, B, L, H, hd, save_lse=True)
    first = ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd)
    for _ in range(6):
        assert torch.equal(ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd), first)


def test_h14_style_tower_forward_backward_vs_oracle():
    """head width 80 (H/14, BASELINE configs[4]) end to end at toy size: embeddings and image / parameter gradients
    against autograd of the CPU oracle (no golden fixture for this config: the oracle is pinned on the others)."""
    cfg_name, batch = "mini-h80", 3
    cfg = synth.CONFIGS[cfg_name]
    sd = synth.make_state_dict(cfg_name, 0)
    m = ovb.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    images = synth.make_images(cfg_name, batch, 0).cuda().requires_grad_(True)
    gy = rnd(batch, cfg["embed_dim"], seed=3)
    out = m.encode_image(images, normalize=True)
    out.backward(gy.cuda())
    sdf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xf = synth.make_images(cfg_name, batch, 0).requires_grad_(True)
    ref = O.l2_normalize(O.vision_transformer(xf, sdf, synth.vision_heads(cfg_name)))
    ref.backward(gy)
    got = out.detach().float().cpu()
    assert (got - ref.detach()).abs().max() <= 2e-2
    assert torch.nn.functional.cosine_similarity(got, ref.detach(), dim=-1).min() >= 0.9995
    assert_grad(images.grad, xf.grad, "d images (hd 80)", 6e-2)
    for name in ("visual.conv1.weight", "visual.transformer.resblocks.0.attn.in_proj_weight",
                 "visual.transformer.resblocks.1.attn.out_proj.weight", "visual.transformer.resblocks.1.mlp.c_fc.weight"):
        p = dict(m.named_parameters())[name]
        assert_grad(p.grad, sdf[name].grad, name, 6e-2)


def test_small_backward_kernels(ops):
    # colsum
    x = rnd(1000, 776, seed=0).bfloat16()
    s = ops.colsum(x.cuda())
    np.testing.assert_allclose(s.cpu().numpy(), x.float().sum(0).numpy(), rtol=1e-4, atol=1e-3)
    # layernorm bwd with the residual gradient folded in
    rows, D = 515, 1024
    xx = (rnd(rows, D, seed=1) * 2 + 0.3).bfloat16()
    dy, dres = rnd(rows, D, seed=2).bfloat16(), rnd(rows, D, seed=3).bfloat16()
    g, b = rnd(D, seed=4), rnd(D, seed=5)
    _, mean, rstd = ops.layernorm(xx.cuda(), g.cuda(), b.cuda(), 1e-6, save_stats=True)
    dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    dx = ops.layernorm_bwd(dy.cuda(), xx.cuda(), g.cuda(), mean, rstd, dg, db, dres=dres.cuda())
    xf = xx.float().requires_grad_(True)
    O.layer_norm(xf, g, b, 1e-6).backward(dy.float())
    assert_grad(dx, xf.grad + dres.float(), "layernorm bwd + residual", 1e-2)
    # l2 normalise bwd
    e = (rnd(300, 768, seed=6) * 3).bfloat16()
    d = rnd(300, 768, seed=7)
    dxn = ops.l2_normalize_bwd(e.cuda(), d.cuda())
    ef = e.float().requires_grad_(True)
    O.l2_normalize(ef).backward(d)
    assert_grad(dxn, ef.grad, "l2 normalise bwd", 1e-2)
    # pooling bwd
    dp = rnd(4, 256, seed=8).bfloat16()
    dxa = ops.pool_tokens_bwd(dp.cuda(), 4, 11, "avg").float().cpu()
    assert torch.equal(dxa[:, 0], torch.zeros(4, 256)) and torch.allclose(dxa[:, 3], (dp.float() / 10).bfloat16().float())
    dxt = ops.pool_tokens_bwd(dp.cuda(), 4, 11, "tok").float().cpu()
    assert torch.equal(dxt[:, 0], dp.float()) and dxt[:, 1:].abs().sum() == 0
    # activation module fwd / bwd
    u = (rnd(64, 512, seed=9) * 2).bfloat16()
    gy = rnd(64, 512, seed=10).bfloat16()
    for kind, okind in (("gelu", "erf"), ("gelu_tanh", "tanh"), ("quick_gelu", "quick")):
        y = ops.act_fwd(u.cuda(), kind)
        uf = u.float().requires_grad_(True)
        ref = O.gelu(uf, okind)
        ref.backward(gy.float())
        assert (y.float().cpu() - ref.detach()).abs().max() <= 1e-2 * ref.abs().max()
        assert_grad(ops.act_bwd(u.cuda(), gy.cuda(), kind), uf.grad, f"act bwd {kind}", 1e-2)


@pytest.mark.parametrize("B,H,P,D", [(2, 48, 16, 128), (2, 56, 14, 128)])
def test_patch_embed_backward(ops, B, H, P, D):
    from openvision_b200.transformer import PatchEmbedConv
    conv = PatchEmbedConv(3, D, P, P, bias=False).cuda()
    img = rnd(B, 3, H, H, seed=0).cuda().requires_grad_(True)
    out = conv(img)                                            # stock contract: [B, D, gh, gw]
    gy = rnd(*out.shape, seed=1).cuda()
    out.backward(gy)
    wf = conv.weight.detach().float().cpu().requires_grad_(True)
    xf = img.detach().float().cpu().requires_grad_(True)
    ref = torch.nn.functional.conv2d(xf, wf, stride=P)         # reference call site transformer.py:469,610
    ref.backward(gy.float().cpu())
    assert_grad(out.detach(), ref.detach(), "conv1 forward", 1e-2)
    assert_grad(img.grad, xf.grad, "d images", 2e-2)
    assert_grad(conv.weight.grad, wf.grad, "d conv1.weight", 2e-2)


@pytest.mark.parametrize("cfg_name,batch", [("mini-ov", 4), ("mini-stock", 4)])
def test_tower_and_loss_gradients_match_reference(golden, cfg_name, batch):
    g = golden(f"tower_{cfg_name}.npz")
    cfg = synth.CONFIGS[cfg_name]
    m = ovb.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]))
    m.load_state_dict(synth.make_state_dict(cfg_name, 0), strict=True)
    m = m.cuda().train()
    images = synth.make_images(cfg_name, batch, 0).cuda().requires_grad_(True)
    text = synth.make_text(cfg_name, batch, 0).cuda()
    img_n, txt_n, scale = m(images, text)      # mini-stock: causal text tower (attention kernels' causal flag, fwd + bwd)
    loss = ovb.ClipLoss()(img_n, txt_n, scale)
    loss.backward()
    ref_loss = float(g["loss"])
    assert abs(float(loss) - ref_loss) <= 2e-2 * abs(ref_loss) + 1e-3, (float(loss), ref_loss)
    assert_grad(images.grad, g["grad_images"], "d images")
    checked = 0
    for name, p in m.named_parameters():
        key = "grad/" + name
        if key not in g:
            continue
        assert p.grad is not None, name
        ref = g[key]
        got = p.grad.detach().float().cpu().numpy()
        if got.size > 4096:
            got = got.reshape(-1)[::5]
        tol = 6e-2 if "ln_" not in name else 8e-2
        assert_grad(got, ref, name, tol)
        checked += 1
    assert checked >= 50, checked


def test_ti16_image_gradient_and_checkpointing(golden):
    """12-layer tower: gradient w.r.t. the images (what ov-gradient-ascent / feature-visualisation optimise) against
    the reference; and activation checkpointing (set_grad_checkpointing) must give the same gradients."""
    cfg_name, batch = "Ti16-160", 8
    g = golden(f"tower_{cfg_name}.npz")
    cfg = synth.CONFIGS[cfg_name]
    m = ovb.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]))
    m.load_state_dict(synth.make_state_dict(cfg_name, 0), strict=True)
    m = m.cuda().train()
    text = synth.make_text(cfg_name, batch, 0).cuda()
    grads = []
    for ckpt in (False, True, "mlp"):
        m.set_grad_checkpointing(ckpt)
        m.zero_grad(set_to_none=True)
        images = synth.make_images(cfg_name, batch, 0).cuda().requires_grad_(True)
        img_n, txt_n, scale = m(images, text)
        ovb.ClipLoss()(img_n, txt_n, scale).backward()
        grads.append((images.grad.clone(), m.visual.conv1.weight.grad.clone()))
    got = grads[0][0].float().cpu().numpy().reshape(-1)[::5]
    assert_grad(got, g["grad_images"], "Ti16 d images", 8e-2)
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1]), "checkpointed != plain"
    # selective mode ('mlp'): backward rebuilds the MLP hidden pair from the stand-alone LayerNorm + fc1 (the forward ran
    # them folded), so the gradients agree to bf16 rounding, not bit for bit
    for a, b, what in ((grads[2][0], grads[0][0], "d images"), (grads[2][1], grads[0][1], "d conv1.weight")):
        err = (a.float() - b.float()).abs().max().item()
        assert err <= 3e-2 * b.float().abs().max().item(), (what, err)


def test_gelu_hook_gradients_reach_the_image():
    """ov-feature-visualization: loss = -mean(post-GELU feature of layer l) optimised w.r.t. the input image."""
    cfg_name = "mini-ov"
    cfg = synth.CONFIGS[cfg_name]
    sd = synth.make_state_dict(cfg_name, 0)
    m = ovb.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    feats = {}
    hook = m.visual.transformer.resblocks[1].mlp.gelu.register_forward_hook(lambda mod, i, o: feats.__setitem__("f", o))
    images = synth.make_images(cfg_name, 2, 0).cuda().requires_grad_(True)
    m.encode_image(images)
    loss = -feats["f"][:, :, 7].float().mean()
    loss.backward()
    hook.remove()
    xf = synth.make_images(cfg_name, 2, 0).requires_grad_(True)
    taps = {}
    O.vision_transformer(xf, sd, synth.vision_heads(cfg_name), taps=taps)
    (-taps["visual.transformer.resblocks.1.mlp.gelu"][:, :, 7].mean()).backward()
    assert_grad(images.grad, xf.grad, "d image through hooked GELU", 6e-2)


def test_l14_full_size_gradients_match_oracle_autograd():
    """The benchmark's tower (ViT-L/14@224, 24 layers, 257 tokens) through forward AND backward against autograd on the
    fp32 CPU oracle, for a handful of parameters spread over the depth and for the input images.  Exercises the production
    kernel selection in training mode: LayerNorm folded forward / recomputed backward, persistent attention backward with
    the remainder tiles of L = 257, CTA-pair and split-K GEMMs."""
    cfg_name, batch = "L14-224", 2
    sd = synth.make_state_dict(cfg_name, 0, vision_only=True)
    watch = ["visual.conv1.weight", "visual.transformer.resblocks.0.attn.in_proj_weight",
             "visual.transformer.resblocks.0.ln_1.weight", "visual.transformer.resblocks.11.mlp.c_fc.weight",
             "visual.transformer.resblocks.23.mlp.c_proj.bias", "visual.transformer.resblocks.23.attn.out_proj.weight",
             "visual.ln_post.bias", "visual.proj"]
    ref_sd = {k: (t.clone().requires_grad_(True) if k in watch else t) for k, t in sd.items()}
    images = synth.make_images(cfg_name, batch, 0)
    ref_img = images.clone().requires_grad_(True)
    w = torch.randn(batch, synth.CONFIGS[cfg_name]["embed_dim"], generator=torch.Generator().manual_seed(3))
    ref_out = O.l2_normalize(O.vision_transformer(ref_img, ref_sd, synth.vision_heads(cfg_name), pool_type="avg",
                                                  final_ln_after_pool=True))
    (ref_out * w).sum().backward()
    cfg = synth.CONFIGS[cfg_name]
    v = ovb.model._build_vision_tower(cfg["embed_dim"], cfg["vision"])
    v.load_state_dict({k[len("visual."):]: t for k, t in sd.items() if k.startswith("visual.")}, strict=True)
    v = v.cuda().train()
    img = images.cuda().requires_grad_(True)
    out = torch.nn.functional.normalize(v(img).float(), dim=-1)
    (out * w.cuda()).sum().backward()
    params = dict(v.named_parameters())
    for k in watch:
        assert_grad(params[k[len("visual."):]].grad, ref_sd[k].grad, k, 8e-2)
    assert_grad(img.grad, ref_img.grad, "d images", 8e-2)


@pytest.mark.parametrize("B,L,H,hd", [(2, 257, 3, 64), (1, 129, 2, 64), (2, 385, 2, 64), (2, 257, 2, 80), (1, 129, 3, 72), (40, 257, 8, 64)])
def test_attention_bwd_remainder_token_outside_the_tiles(ops, B, L, H, hd, monkeypatch):
    """L = 128 k + 1: the remainder token's query row and key / value row come from attention_bwd_tail_kernel and reach the
    other rows as rank-1 terms in the tile kernels' epilogues.  Checked against autograd, row t separately, and against the
    all-tiles path (OVK_ATTBWD_TAIL=0) which computes the same thing with a third tile row / column."""
    qkv = rnd(B * L, 3 * H * hd, seed=L + hd + 3).bfloat16()
    dout = rnd(B * L, H * hd, seed=L + hd + 4).bfloat16()
    qc = qkv.cuda()
    out, lse = ops.attention(qc, B, L, H, hd, save_lse=True)
    monkeypatch.setenv("OVK_ATTBWD_FUSED", "2")
    n0 = ops.launch_count
    dqkv_fused = ops.attention_bwd(qc, out, dout.cuda(), lse, B, L, H, hd)
    assert ops.launch_count - n0 == 3, "tail kernel (leaves delta too) + one-pass tile kernel + dQ conversion"
    monkeypatch.setenv("OVK_ATTBWD_FUSED", "0")
    n0 = ops.launch_count
    dqkv = ops.attention_bwd(qc, out, dout.cuda(), lse, B, L, H, hd)
    assert ops.launch_count - n0 == 3, "tail kernel + dQ + dK/dV"
    monkeypatch.setenv("OVK_ATTBWD_TAIL", "0")
    dqkv_tiles = ops.attention_bwd(qc, out, dout.cuda(), lse, B, L, H, hd)
    qf = qkv.float().requires_grad_(True)
    q, k, v = qf.view(B, L, 3, H, hd).permute(2, 0, 3, 1, 4)
    p = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), -1)
    (p @ v).permute(0, 2, 1, 3).reshape(B * L, H * hd).backward(dout.float())
    g = dqkv.float().cpu().view(B, L, 3, H, hd)
    g2 = dqkv_tiles.float().cpu().view(B, L, 3, H, hd)
    g3 = dqkv_fused.float().cpu().view(B, L, 3, H, hd)
    r = qf.grad.view(B, L, 3, H, hd)
    for i, name in enumerate("qkv"):
        assert_grad(g3[:, :, i], r[:, :, i], f"one-pass d{name} (all rows)", 3e-2)
        assert_grad(g3[:, L - 1, i], r[:, L - 1, i], f"one-pass d{name} (remainder row)", 3e-2)
        assert_grad(g3[:, :, i], g[:, :, i], f"d{name}: one-pass vs two-pass", 2e-2)
        assert_grad(g[:, :, i], r[:, :, i], f"d{name} (all rows)", 3e-2)
        assert_grad(g[:, L - 1, i], r[:, L - 1, i], f"d{name} (remainder row)", 3e-2)
        assert_grad(g[:, :, i], g2[:, :, i], f"d{name}: tail path vs all-tiles path", 2e-2)


@pytest.mark.parametrize("M,N,K,act", [(1000, 512, 128, "gelu"), (4096, 1024, 256, "gelu_tanh"), (777, 3072, 768, "quick_gelu"),
                                       (263, 256, 64, "gelu"), (808, 768, 192, "gelu"), (640, 512, 128, "gelu")])
def test_gelu_backward_gemm_leaves_the_bias_gradient(ops, M, N, K, act):
    """dU = (dY W2) . act'(u) with db1 = dU.sum(0) out of the same GEMM's epilogue (transformer.py:232-236 backward): the
    output must be bitwise the plain fused GELU' GEMM, the column sums those of the bf16 output."""
    dy = rnd(M, K, seed=1).bfloat16().cuda()
    w = rnd(K, N, seed=2, scale=0.1).bfloat16().cuda()
    u = rnd(M, N, seed=3).bfloat16().cuda()
    ref = ops.gemm_nn(dy, w, preact=u, act=act)
    # canary behind the workspace the wrapper is about to allocate from the same pool: the kernel must not write past its rows
    # (CTA pairs: the second CTA of the last pair can sit past the last row of the matrix)
    rows = ops._lib.load().ovk_gemm_colsum_rows(M)
    big = torch.empty((rows + 8, N), dtype=torch.float32, device="cuda")
    big[rows:] = 12345.0
    del big
    out, cs = ops.gemm_nn_dact_colsum(dy, w, u, act)
    assert torch.equal(out, ref)
    want = ref.float().sum(0)
    assert_close_vec = (cs - want).abs().max().item()
    assert assert_close_vec <= 1e-4 * want.abs().max().item() + 1e-3, assert_close_vec

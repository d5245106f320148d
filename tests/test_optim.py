"""Optimizer step: oracle restatement of the reference's optax chain vs torch.optim.AdamW (CPU), and the
libovk kernels / FlatAdamW against the oracle (GPU)."""
import math

import numpy as np
import pytest
import torch

from oracle import optim_oracle as OO


def test_oracle_matches_torch_adamw_known_answer():
    """Same mathematics as torch.optim.AdamW when the first moment is kept in fp32 (decoupled decay scaled by lr)."""
    torch.manual_seed(0)
    p0 = torch.randn(257, dtype=torch.float32)
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p], lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.2)
    pn, mu, nu = p0.numpy().copy(), np.zeros(257, np.float32), np.zeros(257, np.float32)
    for t in range(1, 6):
        g = torch.randn(257, generator=torch.Generator().manual_seed(t))
        p.grad = g.clone()
        opt.step()
        pn, mu, nu = OO.adamw_step(pn, g.numpy(), mu, nu, 3e-3, 0.9, 0.95, 1e-8, 0.2, t, mu_bf16=False)
        np.testing.assert_allclose(pn, p.detach().numpy(), rtol=2e-6, atol=2e-7)


def test_oracle_clip_and_schedule():
    from openvision_b200.optim import cosine_schedule
    g = [np.full(4, 3.0, np.float32), np.full(9, 4.0, np.float32)]          # norm = sqrt(36 + 144)
    assert math.isclose(OO.clip_scale(g, 1.0), 1.0 / math.sqrt(180.0), rel_tol=1e-6)
    assert OO.clip_scale(g, 100.0) == 1.0
    assert cosine_schedule(0, 100, 10) == 0.0 and cosine_schedule(5, 100, 10) == 0.5 and cosine_schedule(10, 100, 10) == 1.0
    assert math.isclose(cosine_schedule(55, 100, 10), 0.5, abs_tol=1e-9) and cosine_schedule(100, 100, 10) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("pdt,gdt", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16)])
@pytest.mark.parametrize("n", [8, 1003, 1 << 20])
def test_adamw_kernel_matches_oracle(pdt, gdt, n):
    from openvision_b200 import ops
    gen = torch.Generator().manual_seed(n)
    p = torch.randn(n, generator=gen).to(pdt)
    mu = torch.zeros(n, dtype=torch.bfloat16)
    nu = torch.zeros(n)
    pc, muc, nuc = p.cuda(), mu.cuda(), nu.cuda()
    pn, mun, nun = p.float().numpy().copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for t in range(1, 4):
        g = (torch.randn(n, generator=gen) * 0.1).to(gdt)
        ops.adamw_step(pc, g.cuda(), muc, nuc, 1e-2, 0.9, 0.95, 1e-8, 0.2, t, gscale=0.5)
        pn, mun, nun = OO.adamw_step(pn, g.float().numpy(), mun, nun, 1e-2, 0.9, 0.95, 1e-8, 0.2, t, gscale=0.5)
        if pdt == torch.bfloat16:
            pn = torch.from_numpy(pn).bfloat16().float().numpy()      # the parameter itself is stored in bf16
        # the stored first moment is bf16: a rounding flip (fma on the GPU vs mul + add here) moves the next update by up to
        # lr * 2^-8 of its size (|update| <= lr / (1 - b1) here), on top of the bf16 storage of a bf16 parameter
        tol = 1e-2 if pdt == torch.bfloat16 else 0.0
        np.testing.assert_allclose(pc.float().cpu().numpy(), pn, rtol=tol, atol=1e-2 * 2.0 ** -7 * t + tol * 1e-1)
        np.testing.assert_allclose(nuc.cpu().numpy(), nun, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(muc.float().cpu().numpy(), mun, rtol=4e-2, atol=2.0 ** -7 * float(np.abs(mun).max()))   # bf16 storage: ulps of the operands


@pytest.mark.gpu
def test_sumsq_and_global_norm_clipping():
    from openvision_b200 import ops
    gen = torch.Generator().manual_seed(3)
    a, b = torch.randn(100003, generator=gen), torch.randn(4096, generator=gen).bfloat16()
    out = torch.zeros(1, device="cuda")
    ops.sumsq(a.cuda(), out)
    ops.sumsq(b.cuda(), out)
    ref = float((a.double() ** 2).sum() + (b.double() ** 2).sum())
    assert abs(float(out) - ref) <= 1e-5 * ref
    # clipping inside the step: same result as scaling the gradient by the oracle's clip factor
    n = 4099
    p, g = torch.randn(n, generator=gen), torch.randn(n, generator=gen)
    gsq = (g.cuda() ** 2).sum().reshape(1).float()
    pc, mu, nu = p.cuda(), torch.zeros(n, dtype=torch.bfloat16, device="cuda"), torch.zeros(n, device="cuda")
    ops.adamw_step(pc, g.cuda(), mu, nu, 1e-2, 0.9, 0.95, 1e-8, 0.0, 1, gscale=1.0, gnorm_sq=gsq, max_norm=1.0)
    pn, _, _ = OO.adamw_step(p.numpy(), g.numpy(), np.zeros(n, np.float32), np.zeros(n, np.float32), 1e-2, 0.9, 0.95, 1e-8, 0.0, 1,
                             gscale=OO.clip_scale([g.numpy()], 1.0))
    np.testing.assert_allclose(pc.cpu().numpy(), pn, rtol=2e-5, atol=2e-6)


@pytest.mark.gpu
def test_flat_adamw_on_the_tower_matches_per_parameter_oracle():
    """FlatAdamW re-points p.data / p.grad into flat buffers; a training step through the drop-in tower then updates every
    parameter as the oracle does (weight matrices decay, biases / LayerNorm / cls / pos do not)."""
    import openvision_b200 as ovb
    from openvision_b200.optim import FlatAdamW, default_decay_mask
    from oracle import synth
    cfg = synth.CONFIGS["mini-ov"]
    torch.manual_seed(0)
    tower = ovb.model._build_vision_tower(cfg["embed_dim"], cfg["vision"]).cuda().train()
    before = {n: p.detach().float().cpu().numpy().copy() for n, p in tower.named_parameters()}
    opt = FlatAdamW(tower.named_parameters(), lr=1e-2, weight_decay=0.2)
    for n, p in tower.named_parameters():
        np.testing.assert_array_equal(p.detach().float().cpu().numpy(), before[n])      # re-pointing kept the values
    side = cfg["vision"]["image_size"]
    img = torch.randn(4, 3, side, side, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    opt.zero_grad()
    y0 = tower(img)
    y0.float().square().mean().backward()
    grads = {n: p.grad.detach().float().cpu().numpy().copy() for n, p in tower.named_parameters()}
    assert any(np.abs(g).max() > 0 for g in grads.values())
    opt.step(lr_mult=0.5)
    for n, p in tower.named_parameters():
        wd = 0.2 if default_decay_mask(n, p) else 0.0
        z = np.zeros_like(before[n]).ravel()
        ref, _, _ = OO.adamw_step(before[n].ravel(), grads[n].ravel(), z, z.copy(), 5e-3, 0.9, 0.95, 1e-8, wd, 1)
        np.testing.assert_allclose(p.detach().float().cpu().numpy().ravel(), ref, rtol=2e-5, atol=1e-5, err_msg=n)
    with torch.no_grad():      # the packed bf16 / LayerNorm-folded weight caches must notice the in-place update
        y1 = tower(img)
    assert (y1.float() - y0.detach().float()).abs().max().item() > 1e-3
    decayed = [n for n, p in tower.named_parameters() if default_decay_mask(n, p)]
    assert "conv1.weight" in decayed and "proj" in decayed and not any(n.endswith("bias") or "ln_" in n for n in decayed)

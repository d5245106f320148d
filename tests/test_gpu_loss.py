"""GPU parity of the fused contrastive loss (kernels + ClipLoss module) against the committed outputs of the unmodified
reference (tests/golden/loss_*.npz) and the CPU oracle.  Bar (north_star): loss within 1e-3 relative; gradients are
compared with max-abs error relative to the largest reference entry (bf16 features and a bf16 G matrix in between)."""
import numpy as np
import pytest
import torch

from oracle import synth, vit_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from openvision_b200 import ops as _ops
    return _ops


def _feat(n, e, seed):
    img, txt = synth.make_features(n, e, seed=seed, dtype=torch.float32)
    return img.bfloat16(), txt.bfloat16()


@pytest.mark.parametrize("n,e,scale", [(64, 32, 14.2857), (256, 64, 100.0), (200, 48, 30.0), (1000, 768, 14.2857),
                                       (129, 8, 5.0), (1, 8, 3.0), (515, 104, 60.0)])
def test_fwd_statistics_match_oracle(ops, n, e, scale):
    img, txt = _feat(n, e, n + e)
    row_lse, diag, col_max, col_sum = ops.clip_loss_fwd(img.cuda(), txt.cuda(), 0, scale)
    z = scale * img.double() @ txt.double().t()
    np.testing.assert_allclose(row_lse.cpu().numpy(), torch.logsumexp(z, 1).numpy(), rtol=0, atol=2e-4 * max(1, scale))
    np.testing.assert_allclose(diag.cpu().numpy(), z.diagonal().numpy(), rtol=0, atol=2e-5 * max(1, scale))
    col_lse = ops.clip_loss_combine(col_max[None].contiguous(), col_sum[None].contiguous())
    np.testing.assert_allclose(col_lse.cpu().numpy(), torch.logsumexp(z, 0).numpy(), rtol=0, atol=2e-4 * max(1, scale))
    out = ops.clip_loss_value(row_lse, col_lse, diag, 0).cpu()
    ref = float(O.clip_loss(img.double(), txt.double(), torch.tensor(scale, dtype=torch.float64)))
    assert abs(float(out[0]) - ref) <= 1e-3 * abs(ref) + 1e-5


@pytest.mark.parametrize("n_loc", [515, 640, 896, 384])
def test_fwd_never_writes_past_its_workspace(ops, n_loc):
    """n_loc >= 512 with an odd number of 128-row tiles: the second CTA of the last pair has no rows (and no workspace slot).
    Canary words directly behind the exact-size workspace must survive (out-of-bounds column-statistics store)."""
    from openvision_b200 import _lib
    n_all, e = 1024, 64
    img, txt = _feat(n_all, e, 3)
    img, txt = img[:n_loc].contiguous().cuda(), txt.cuda()
    nws = _lib.load().ovk_clip_loss_workspace_floats(n_loc, n_all)
    guard = 4 * n_all
    buf = torch.full((nws + guard,), 12345.0, dtype=torch.float32, device="cuda")
    diag = torch.empty(n_loc, dtype=torch.float32, device="cuda")
    ops.clip_loss_fwd_window(img, txt, 0, n_all, 0, 20.0, diag, buf[:nws])
    row_lse, col_max, col_sum = ops.clip_loss_finalize(buf[:nws], n_loc, n_all)
    torch.cuda.synchronize()
    assert bool((buf[nws:] == 12345.0).all()), "clip_loss_fwd wrote behind its workspace"
    z = 20.0 * img.double().cpu() @ txt.double().cpu().t()
    np.testing.assert_allclose(row_lse.cpu().numpy(), torch.logsumexp(z, 1).numpy(), rtol=0, atol=4e-3)


@pytest.mark.parametrize("n_loc,n_all,row_offset,cuts", [(256, 1024, 256, (256, 512)), (640, 1280, 512, (512, 1024)),
                                                        (128, 1000, 256, (256, 768)), (512, 2048, 1024, (1024, 1536))])
def test_fwd_column_windows_equal_one_pass(ops, n_loc, n_all, row_offset, cuts):
    """The all-gather overlap of the multi-rank loss: column windows processed by separate launches (the rank's own block
    first, in any order) must give the statistics of the single launch."""
    e, scale = 96, 40.0
    img, txt = _feat(n_all, e, 11)
    a = img[row_offset:row_offset + n_loc].contiguous().cuda()
    b = txt.cuda()
    want = ops.clip_loss_fwd(a, b, row_offset, scale)
    ws = ops.clip_loss_workspace(n_loc, n_all, a.device)
    ws.fill_(float("nan"))
    diag = torch.full((n_loc,), float("nan"), dtype=torch.float32, device="cuda")
    edges = [0, *cuts, n_all]
    wins = [(edges[i], edges[i + 1]) for i in range(len(edges) - 1)]
    for (c0, c1) in reversed(wins):
        ops.clip_loss_fwd_window(a, b[c0:c1].contiguous(), c0, n_all, row_offset, torch.tensor(scale, device="cuda"), diag, ws)
    row_lse, col_max, col_sum = ops.clip_loss_finalize(ws, n_loc, n_all)
    np.testing.assert_allclose(row_lse.cpu().numpy(), want[0].cpu().numpy(), rtol=0, atol=1e-4)
    np.testing.assert_allclose(diag.cpu().numpy(), want[1].cpu().numpy(), rtol=0, atol=1e-5)
    cl = ops.clip_loss_combine(col_max[None].contiguous(), col_sum[None].contiguous())
    cl_want = ops.clip_loss_combine(want[2][None].contiguous(), want[3][None].contiguous())
    np.testing.assert_allclose(cl.cpu().numpy(), cl_want.cpu().numpy(), rtol=0, atol=1e-4)


def test_fwd_is_exact_for_badly_scaled_rows_and_columns(ops):
    """Rows / columns whose logits sit hundreds of nats below their neighbours (the single-exponential fast path would
    flush them to zero) must take the exact path: un-normalised features with norms spread over 3 decades."""
    n, e, scale = 384, 64, 50.0
    g = torch.Generator().manual_seed(5)
    img = torch.randn(n, e, generator=g)
    txt = torch.randn(n, e, generator=g)
    img[::7] *= 1e-3          # tiny rows
    txt[3::11] *= 1e-3        # tiny columns
    img[5] *= 30.0            # one huge row next to them
    txt[8] *= 30.0
    img, txt = (img * 0.3).bfloat16(), (txt * 0.3).bfloat16()
    row_lse, diag, col_max, col_sum = ops.clip_loss_fwd(img.cuda(), txt.cuda(), 0, scale)
    col_lse = ops.clip_loss_combine(col_max[None].contiguous(), col_sum[None].contiguous())
    z = scale * img.double() @ txt.double().t()
    ref_r, ref_c = torch.logsumexp(z, 1), torch.logsumexp(z, 0)
    assert torch.isfinite(row_lse).all() and torch.isfinite(col_lse).all()
    tol = 1e-5 * z.abs().max().item() + 1e-3
    assert (row_lse.cpu().double() - ref_r).abs().max().item() <= tol
    assert (col_lse.cpu().double() - ref_c).abs().max().item() <= tol


@pytest.mark.parametrize("name", ["loss_N64_E32_s14", "loss_N256_E64_s14", "loss_N256_E64_s100", "loss_N200_E48_s30"])
def test_cliploss_module_matches_reference_golden(golden, name):
    import openvision_b200 as ovb
    g = golden(name + ".npz")
    n, e = g["d_img"].shape
    img, txt = synth.make_features(n, e, seed=int(g["seed"]), dtype=torch.float32)
    img = img.cuda().requires_grad_(True)
    txt = txt.cuda().requires_grad_(True)
    ls = torch.tensor(float(np.log(g["scale"])), device="cuda", requires_grad=True)
    loss = ovb.ClipLoss()(img, txt, ls.exp())
    loss.backward()
    ref = float(g["loss"])
    # the golden is the fp64 reference on UNROUNDED features; ClipLoss rounds them to bf16 (as the reference's bf16 path
    # does), which moves each logit by ~ scale * 2^-9: allow 1e-5 * scale on top of the 1e-3 relative bar ...
    assert abs(float(loss) - ref) <= 1e-3 * abs(ref) + 1e-5 * float(g["scale"]) + 1e-4, (float(loss), ref)
    # ... and hold the kernels themselves to 1e-3 relative against the oracle on the same bf16-rounded features
    ref_b = float(O.clip_loss(img.detach().bfloat16().double().cpu(), txt.detach().bfloat16().double().cpu(),
                              torch.tensor(float(g["scale"]), dtype=torch.float64)))
    assert abs(float(loss) - ref_b) <= 1e-3 * abs(ref_b) + 1e-5, (float(loss), ref_b)
    for got, want, what in ((img.grad, g["d_img"], "d_img"), (txt.grad, g["d_txt"], "d_txt")):
        err = np.abs(got.cpu().numpy() - want).max()
        assert err <= 2e-2 * np.abs(want).max(), (what, err, np.abs(want).max())
    assert abs(float(ls.grad) - float(g["d_logit_scale"])) <= 2e-2 * abs(float(g["d_logit_scale"])) + 1e-4
    d = ovb.ClipLoss()(img.detach(), txt.detach(), ls.exp().detach(), output_dict=True)
    assert set(d) == {"contrastive_loss"}


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("local_loss,gwg", [(True, True), (False, False)])
def test_rank_row_blocks_on_one_gpu_match_multirank_reference(ops, golden, world, local_loss, gwg):
    """The W ranks' work emulated on one GPU (one kernel call per rank's row block, collectives replaced by torch
    indexing): per-rank losses and gradients against the reference run under gloo with W processes."""
    from openvision_b200.loss import loss_weights
    g = golden(f"loss_dist_W{world}_N64_E32_local{int(local_loss)}_gwg{int(gwg)}.npz")
    n, e, scale = 64, 32, float(g["scale"])
    img, txt = _feat(n, e, int(g["seed"]))
    img, txt = img.cuda(), txt.cuda()
    nl = n // world
    parts = [ops.clip_loss_fwd(img[r * nl:(r + 1) * nl].contiguous(), txt, r * nl, scale) for r in range(world)]
    col_lse = ops.clip_loss_combine(torch.stack([p[2] for p in parts]), torch.stack([p[3] for p in parts]))
    losses = [float(ops.clip_loss_value(parts[r][0], col_lse, parts[r][1], r * nl)[0]) for r in range(world)]
    if not local_loss:
        losses = [float(np.mean(losses))] * world
    d_txt_sum = torch.zeros(n, e, device="cuda")
    for r in range(world):
        ref = float(g[f"loss_r{r}"])
        assert abs(losses[r] - ref) <= 1e-3 * abs(ref) + 2e-3
        w = loss_weights(1.0, nl, n, world, local_loss, gwg)
        ds = torch.zeros(1, device="cuda")
        a = img[r * nl:(r + 1) * nl].contiguous()
        G = ops.clip_loss_grad_logits(a, txt, r * nl, scale, parts[r][0], col_lse, w, w, ds)
        d_img = ops.gemm_nn(G.contiguous(), txt, alpha=scale, out_dtype=torch.float32)
        d_txt_sum += ops.gemm_tn(G.contiguous(), a, alpha=scale, out_dtype=torch.float32)      # reduce-scatter = sum, then slice
        want = g[f"d_img_r{r}"]
        assert np.abs(d_img.cpu().numpy() - want).max() <= 3e-2 * np.abs(want).max()
    for r in range(world):
        want = g[f"d_txt_r{r}"]
        assert np.abs(d_txt_sum[r * nl:(r + 1) * nl].cpu().numpy() - want).max() <= 3e-2 * np.abs(want).max()


def test_full_size_properties_32k(ops):
    """BASELINE configs[3] size (N = 32768, E = 768; the 4 GiB logit matrix is never formed).  Size-independent checks:
    (a) row/column LSE >= the positive-pair logit, (b) the loss of (I, T) equals the loss of (T, I) (symmetry of the
    objective), (c) permuting the pairs leaves the loss unchanged, (d) sum of G over a row = w_row * 1 + w_col * (...)
    -> checked through d_scale: sum(G*z)/s computed by the kernel equals the finite-difference slope of the loss."""
    import openvision_b200 as ovb
    n, e, scale = 32768, 768, 14.2857
    gen = torch.Generator(device="cuda").manual_seed(0)
    a = torch.nn.functional.normalize(torch.randn(n, e, device="cuda", generator=gen), dim=-1)
    b = torch.nn.functional.normalize(0.5 * a + 0.9 * torch.nn.functional.normalize(torch.randn(n, e, device="cuda", generator=gen), dim=-1), dim=-1)
    a, b = a.bfloat16(), b.bfloat16()
    row_lse, diag, cmax, csum = ops.clip_loss_fwd(a, b, 0, scale)
    col_lse = ops.clip_loss_combine(cmax[None].contiguous(), csum[None].contiguous())
    assert (row_lse >= diag - 1e-4).all() and (col_lse >= diag - 1e-4).all()
    l_ab = float(ops.clip_loss_value(row_lse, col_lse, diag, 0)[0])
    r2, d2, cm2, cs2 = ops.clip_loss_fwd(b, a, 0, scale)
    c2 = ops.clip_loss_combine(cm2[None].contiguous(), cs2[None].contiguous())
    l_ba = float(ops.clip_loss_value(r2, c2, d2, 0)[0])
    assert abs(l_ab - l_ba) <= 1e-5 * abs(l_ab)
    assert torch.allclose(row_lse, c2, atol=2e-4) and torch.allclose(col_lse, r2, atol=2e-4)
    perm = torch.randperm(n, device="cuda", generator=gen)
    r3, d3, cm3, cs3 = ops.clip_loss_fwd(a[perm].contiguous(), b[perm].contiguous(), 0, scale)
    c3 = ops.clip_loss_combine(cm3[None].contiguous(), cs3[None].contiguous())
    l_perm = float(ops.clip_loss_value(r3, c3, d3, 0)[0])
    assert abs(l_ab - l_perm) <= 1e-5 * abs(l_ab)
    # a blocked fp64 oracle of the same loss on the same bf16 features (CPU-free: blocks of 2048 rows on the GPU in fp64
    # would use ATen math as the checker only; sizes the CPU oracle cannot finish in seconds)
    crit = ovb.ClipLoss()
    af, bf = a.float().requires_grad_(True), b.float().requires_grad_(True)
    ls = torch.tensor(float(np.log(scale)), device="cuda", requires_grad=True)
    loss = crit(af, bf, ls.exp())
    loss.backward()
    assert abs(float(loss) - l_ab) <= 1e-5 * abs(l_ab)
    eps = 1e-2
    lp = float(crit(af.detach(), bf.detach(), torch.tensor(float(np.log(scale)) + eps, device="cuda").exp()))
    lm = float(crit(af.detach(), bf.detach(), torch.tensor(float(np.log(scale)) - eps, device="cuda").exp()))
    fd = (lp - lm) / (2 * eps)
    assert abs(float(ls.grad) - fd) <= 2e-2 * abs(fd) + 1e-3, (float(ls.grad), fd)
    # gradient rows are orthogonal-ish sanity: dI has the shape / dtype of the input and is finite
    assert af.grad.shape == af.shape and torch.isfinite(af.grad).all() and torch.isfinite(bf.grad).all()
    # blocked fp64 check of the gradients of the first 256 pairs (torch on the GPU as the CHECKER only: the CPU oracle
    # cannot finish N = 32768 in seconds; same formulas as oracle/vit_oracle.py clip_loss_grads)
    ad, bd = a.double(), b.double()
    col_lse64 = torch.full((n,), -float("inf"), dtype=torch.float64, device="cuda")
    for r0 in range(0, n, 2048):
        col_lse64 = torch.logaddexp(col_lse64, torch.logsumexp(scale * ad[r0:r0 + 2048] @ bd.t(), dim=0))
    z = scale * ad[:256] @ bd.t()
    G = (torch.softmax(z, dim=1) + torch.exp(z - col_lse64[None, :])) / (2.0 * n)
    G[torch.arange(256), torch.arange(256)] -= 1.0 / n
    d_img_ref = scale * G @ bd
    err = (af.grad[:256].double() - d_img_ref).abs().max().item()
    assert err <= 2e-2 * d_img_ref.abs().max().item(), (err, d_img_ref.abs().max().item())
    assert (col_lse.double() - col_lse64).abs().max().item() <= 1e-3
    # text gradient of the first 256 texts: dT_j = s * sum_i G_ij I_i needs all rows -> blocked as well
    d_txt_ref = torch.zeros(256, e, dtype=torch.float64, device="cuda")
    for r0 in range(0, n, 2048):
        zb = scale * ad[r0:r0 + 2048] @ bd[:256].t()
        Gb = (torch.exp(zb - row_lse[r0:r0 + 2048].double()[:, None]) + torch.exp(zb - col_lse64[None, :256])) / (2.0 * n)
        if r0 == 0:
            Gb[torch.arange(256), torch.arange(256)] -= 1.0 / n
        d_txt_ref += scale * Gb.t() @ ad[r0:r0 + 2048]
    err = (bf.grad[:256].double() - d_txt_ref).abs().max().item()
    assert err <= 2e-2 * d_txt_ref.abs().max().item(), (err, d_txt_ref.abs().max().item())


def test_dual_caption_loss_matches_oracle():
    """image vs two caption sets, (ClipLoss(I, T1) + ClipLoss(I, T2)) / 2 (losses/common.py:139-171);
    loss and the gradients w.r.t. the image features, both caption sets and the temperature against autograd on the oracle."""
    import openvision_b200 as ovb
    n, e, s = 200, 64, 14.2857
    img, t1 = _feat(n, e, 11)
    _, t2 = _feat(n, e, 12)
    ref_in = [x.double().requires_grad_(True) for x in (img, t1, t2)]
    ls = torch.tensor(s, dtype=torch.float64, requires_grad=True)
    ref = O.dual_caption_loss(*ref_in, ls)
    ref.backward()
    xin = [x.cuda().requires_grad_(True) for x in (img, t1, t2)]
    scale = torch.tensor(s, device="cuda", requires_grad=True)
    loss = ovb.DualCaptionClipLoss()(xin[0], xin[1], xin[2], scale)
    loss.backward()
    assert abs(float(loss) - float(ref)) <= 1e-3 * abs(float(ref))
    for got, want, what in zip(xin, ref_in, ("d image", "d captions 1", "d captions 2")):
        err = (got.grad.double().cpu() - want.grad).abs().max().item()
        assert err <= 2e-2 * want.grad.abs().max().item(), (what, err)
    assert abs(float(scale.grad) - float(ls.grad)) <= 2e-2 * abs(float(ls.grad)) + 1e-6

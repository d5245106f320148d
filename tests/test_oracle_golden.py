"""CPU: the oracle restatement (oracle/vit_oracle.py) against the committed outputs of the unmodified reference
(tests/golden/*.npz, written by oracle/make_golden.py in the build container)."""
import numpy as np
import pytest
import torch

from oracle import synth, vit_oracle

SUB, SUB_MIN = 5, 4096


def _sub(a):
    a = np.asarray(a, dtype=np.float64)
    return a.reshape(-1)[::SUB] if a.size > SUB_MIN else a


def _oracle_tower(cfg_name, batch, dtype=torch.float64, taps=None):
    cfg = synth.CONFIGS[cfg_name]
    # weights are defined as fp32 values (what a checkpoint holds), then widened — exactly what make_golden.py did
    sd = {k: v.to(dtype) for k, v in synth.make_state_dict(cfg_name, seed=0, dtype=torch.float32).items()}
    images = synth.make_images(cfg_name, batch, seed=0, dtype=dtype)
    text = synth.make_text(cfg_name, batch, seed=0)
    v, t = cfg["vision"], cfg["text"]
    img = vit_oracle.vision_transformer(images, sd, synth.vision_heads(cfg_name), pool_type=v["pool_type"],
                                        final_ln_after_pool=v["final_ln_after_pool"], taps=taps)
    act = "tanh" if (t.get("act_kwargs") or {}).get("approximate") == "tanh" else "erf"
    txt = vit_oracle.text_transformer(text, sd, t["heads"], causal=not t.get("no_causal_mask", False),
                                      pool_type=t["pool_type"], act=act)
    return img, txt, sd


@pytest.mark.parametrize("cfg_name,batch", [("mini-ov", 4), ("mini-stock", 4), ("Ti16-160", 8)])
def test_tower_matches_reference_fp64(golden, cfg_name, batch):
    g = golden(f"tower_{cfg_name}.npz")
    taps = {}
    img, txt, sd = _oracle_tower(cfg_name, batch, taps=taps)
    np.testing.assert_allclose(img.numpy(), g["image_features"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(txt.numpy(), g["text_features"], rtol=1e-9, atol=1e-11)
    img_n, txt_n = vit_oracle.l2_normalize(img), vit_oracle.l2_normalize(txt)
    np.testing.assert_allclose(img_n.numpy(), g["image_features_norm"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(txt_n.numpy(), g["text_features_norm"], rtol=1e-9, atol=1e-12)
    scale = sd["logit_scale"].exp()
    assert abs(float(scale) - float(g["logit_scale_exp"])) < 1e-12
    loss = vit_oracle.clip_loss(img_n, txt_n, scale)
    assert abs(float(loss) - float(g["loss"])) < 1e-10 * abs(float(g["loss"]))
    # post-GELU activations seen by the reference's forward hooks (cliptoolsoptimized.py:1149-1164)
    gelu0 = taps["visual.transformer.resblocks.0.mlp.gelu"].numpy()
    ref0 = g["gelu_block0"]
    if ref0.ndim == 1:
        gelu0 = gelu0.reshape(-1)[::SUB]
        np.testing.assert_allclose(gelu0, ref0, rtol=1e-5, atol=1e-6)   # fixture stored as fp32 samples
    else:
        np.testing.assert_allclose(gelu0, ref0, rtol=1e-9, atol=1e-11)


def test_tower_fp32_close_to_reference_fp32(golden):
    g = golden("tower_Ti16-160.npz")
    img, _, _ = _oracle_tower("Ti16-160", 8, dtype=torch.float32)
    np.testing.assert_allclose(img.numpy(), g["image_features_f32"], rtol=0, atol=2e-4 * np.abs(g["image_features_f32"]).max())


@pytest.mark.parametrize("name", ["loss_N64_E32_s14", "loss_N256_E64_s14", "loss_N256_E64_s100", "loss_N200_E48_s30"])
def test_loss_and_closed_form_grads(golden, name):
    g = golden(name + ".npz")
    n, e = g["d_img"].shape
    img, txt = synth.make_features(n, e, seed=int(g["seed"]), dtype=torch.float64)
    s = torch.tensor(float(g["scale"]), dtype=torch.float64)
    loss = vit_oracle.clip_loss(img, txt, s)
    assert abs(float(loss) - float(g["loss"])) < 1e-10 * abs(float(g["loss"]))
    gi, gt, gs = vit_oracle.clip_loss_grads(img, txt, s)
    np.testing.assert_allclose(gi.numpy(), g["d_img"], rtol=1e-8, atol=1e-13)
    np.testing.assert_allclose(gt.numpy(), g["d_txt"], rtol=1e-8, atol=1e-13)
    # reference differentiates w.r.t. logit_scale (the log): d/d(log s) = s * d/ds
    assert abs(float(gs * s) - float(g["d_logit_scale"])) < 1e-9 * abs(float(g["d_logit_scale"])) + 1e-14


@pytest.mark.parametrize("world", [2, 4])
def test_local_loss_restatement_matches_multirank_reference(golden, world):
    """clip_loss_local (rank-local rows against gathered features) vs the reference ClipLoss run under gloo."""
    g = golden(f"loss_dist_W{world}_N64_E32_local1_gwg1.npz")
    img, txt = synth.make_features(64, 32, seed=int(g["seed"]), dtype=torch.float64)
    s = torch.tensor(float(g["scale"]), dtype=torch.float64)
    losses = [float(vit_oracle.clip_loss_local(img, txt, s, r, world)) for r in range(world)]
    for r in range(world):
        assert abs(losses[r] - float(g[f"loss_r{r}"])) < 1e-10
    # mean of rank-local losses is the global loss; per-rank dI is W x the global-loss gradient
    glob = float(vit_oracle.clip_loss(img, txt, s))
    assert abs(np.mean(losses) - glob) < 1e-10
    gi, gt, _ = vit_oracle.clip_loss_grads(img, txt, s)
    nl = 64 // world
    for r in range(world):
        np.testing.assert_allclose(world * gi[r * nl:(r + 1) * nl].numpy(), g[f"d_img_r{r}"], rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(world * gt[r * nl:(r + 1) * nl].numpy(), g[f"d_txt_r{r}"], rtol=1e-7, atol=1e-12)


def test_autograd_of_oracle_matches_reference_param_grads(golden):
    """Differentiate the oracle with autograd and compare with the reference's parameter / image gradients."""
    cfg_name, batch = "mini-ov", 4
    g = golden(f"tower_{cfg_name}.npz")
    cfg = synth.CONFIGS[cfg_name]
    sd = {k: v.double().requires_grad_(True) for k, v in synth.make_state_dict(cfg_name, 0, torch.float32).items()}
    images = synth.make_images(cfg_name, batch, 0, torch.float64).requires_grad_(True)
    text = synth.make_text(cfg_name, batch, 0)
    v, t = cfg["vision"], cfg["text"]
    img = vit_oracle.vision_transformer(images, sd, synth.vision_heads(cfg_name), pool_type=v["pool_type"],
                                        final_ln_after_pool=v["final_ln_after_pool"])
    txt = vit_oracle.text_transformer(text, sd, t["heads"], causal=False, pool_type=t["pool_type"], act="tanh")
    loss = vit_oracle.clip_loss(vit_oracle.l2_normalize(img), vit_oracle.l2_normalize(txt), sd["logit_scale"].exp())
    loss.backward()
    np.testing.assert_allclose(images.grad.numpy(), g["grad_images"], rtol=1e-7, atol=1e-12)
    checked = 0
    for k, ref in g.items():
        if not k.startswith("grad/"):
            continue
        name = k[5:]
        np.testing.assert_allclose(_sub(sd[name].grad.numpy()), ref, rtol=1e-7, atol=1e-12, err_msg=name)
        checked += 1
    assert checked > 40

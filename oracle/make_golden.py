"""Golden-vector generator + oracle pinning (test infrastructure; runs in the BUILD container only).

Imports the UNMODIFIED reference modules from /root/reference (oracle/ref_loader.py), feeds them the deterministic
synthetic weights / inputs of oracle/synth.py, and

  1. PINS the CPU restatement oracle/vit_oracle.py against the reference (fp64 agreement <= 1e-9 relative, asserted);
  2. writes the reference's own outputs (fp32 and fp64 runs) as small fixtures under tests/golden/*.npz.

The reference ships no tests or golden vectors for this path, so these fixtures — outputs of the
reference itself — are what parity is anchored on.  Re-run:  python -m oracle.make_golden
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader, synth, vit_oracle  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
SUBSAMPLE = 5          # gradients of tensors with more than SUBSAMPLE_MIN elements are stored as flat[::SUBSAMPLE]
SUBSAMPLE_MIN = 4096

TOWER_CASES = [("mini-ov", 4), ("mini-stock", 4), ("Ti16-160", 8)]
LOSS_CASES = [(64, 32, 1.0 / 0.07), (256, 64, 1.0 / 0.07), (256, 64, 100.0), (200, 48, 30.0)]


def sub(t: torch.Tensor) -> np.ndarray:
    a = t.detach().cpu().double().numpy()
    if a.size > SUBSAMPLE_MIN:
        return a.reshape(-1)[::SUBSAMPLE].copy()
    return a


def build_reference_clip(mdl, cfg_name: str, dtype):
    cfg = synth.CONFIGS[cfg_name]
    model = mdl.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]))
    sd = synth.make_state_dict(cfg_name, seed=0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    # attn_mask is a non-persistent buffer; everything else must match the key contract exactly
    assert not missing and not unexpected, (missing, unexpected)
    return model.to(dtype).eval(), sd


def oracle_forward(cfg_name, sd, images, text):
    cfg = synth.CONFIGS[cfg_name]
    v, t = cfg["vision"], cfg["text"]
    img = vit_oracle.vision_transformer(images, sd, synth.vision_heads(cfg_name), pool_type=v["pool_type"],
                                        final_ln_after_pool=v["final_ln_after_pool"])
    act = "tanh" if (t.get("act_kwargs") or {}).get("approximate") == "tanh" else "erf"
    txt = vit_oracle.text_transformer(text, sd, t["heads"], causal=not t.get("no_causal_mask", False),
                                      pool_type=t["pool_type"], act=act)
    return img, txt


def relerr(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def tower_case(tr, mdl, lossm, cfg_name, batch):
    out = {}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        model, sd = build_reference_clip(mdl, cfg_name, dtype)
        images = synth.make_images(cfg_name, batch, seed=0, dtype=dtype)
        text = synth.make_text(cfg_name, batch, seed=0)
        images.requires_grad_(True)
        taps = {}
        hooks = []
        for i, blk in enumerate(model.visual.transformer.resblocks):
            hooks.append(blk.mlp.gelu.register_forward_hook(
                lambda m, inp, o, i=i: taps.__setitem__(i, o.detach())))
        img_raw = model.encode_image(images)
        txt_raw = model.encode_text(text)
        img_n, txt_n, scale = model(images, text)
        for h in hooks:
            h.remove()
        loss = lossm.ClipLoss()(img_n, txt_n, scale)
        loss.backward()
        # ---- pin the restatement against the reference
        sdd = {k: v.to(dtype) for k, v in sd.items()}
        o_img, o_txt = oracle_forward(cfg_name, sdd, images.detach(), text)
        e_img, e_txt = relerr(o_img, img_raw.detach()), relerr(o_txt, txt_raw.detach())
        o_loss = vit_oracle.clip_loss(vit_oracle.l2_normalize(o_img), vit_oracle.l2_normalize(o_txt), scale.detach())
        e_loss = abs(float(o_loss) - float(loss)) / abs(float(loss))
        tol = 1e-9 if dtype == torch.float64 else 2e-4
        print(f"[pin {cfg_name} {tag}] oracle vs reference: image {e_img:.2e} text {e_txt:.2e} loss {e_loss:.2e}")
        assert e_img < tol and e_txt < tol and e_loss < tol, "oracle restatement disagrees with the reference"
        if dtype == torch.float64:
            out.update(
                image_features=img_raw.detach().numpy(), text_features=txt_raw.detach().numpy(),
                image_features_norm=img_n.detach().numpy(), text_features_norm=txt_n.detach().numpy(),
                logit_scale_exp=np.float64(scale.detach()), loss=np.float64(loss.detach()),
                grad_images=images.grad.numpy(), gelu_block0=taps[0].numpy(),
                gelu_last=taps[len(model.visual.transformer.resblocks) - 1].numpy())
            if not cfg_name.startswith("mini"):   # keep the big-config fixture small: strided fp32 samples
                for k in ("grad_images", "gelu_block0", "gelu_last"):
                    out[k] = out[k].reshape(-1)[::SUBSAMPLE].astype(np.float32)
            if cfg_name.startswith("mini"):
                for name, p in model.named_parameters():
                    if p.grad is not None:
                        out["grad/" + name] = sub(p.grad)
        else:
            out["image_features_f32"] = img_raw.detach().numpy()
            out["loss_f32"] = np.float32(loss.detach())
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"tower_{cfg_name}.npz"), **out)
    print(f"  wrote tower_{cfg_name}.npz ({len(out)} arrays)")


def loss_case(lossm, n, e, scale):
    img, txt = synth.make_features(n, e, seed=n + e, dtype=torch.float64)
    img.requires_grad_(True)
    txt.requires_grad_(True)
    ls = torch.tensor(np.log(scale), dtype=torch.float64, requires_grad=True)
    loss = lossm.ClipLoss()(img, txt, ls.exp())
    loss.backward()
    # pin the restatement (closed-form gradients included)
    o = vit_oracle.clip_loss(img.detach(), txt.detach(), ls.exp().detach())
    gi, gt, gs = vit_oracle.clip_loss_grads(img.detach(), txt.detach(), ls.exp().detach())
    assert abs(float(o) - float(loss)) < 1e-10 * abs(float(loss))
    assert relerr(gi, img.grad) < 1e-9 and relerr(gt, txt.grad) < 1e-9
    assert abs(float(gs * ls.exp().detach()) - float(ls.grad)) < 1e-9 * abs(float(ls.grad)) + 1e-14
    print(f"[pin loss N={n} E={e} s={scale:.3f}] loss {float(loss):.6f}: oracle closed-form grads agree")
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"loss_N{n}_E{e}_s{int(round(scale))}.npz"),
                        loss=np.float64(loss.detach()), d_img=img.grad.numpy(), d_txt=txt.grad.numpy(),
                        d_logit_scale=np.float64(ls.grad), scale=np.float64(scale), seed=np.int64(n + e))


def _dist_worker(rank, world, n, e, scale, local_loss, gather_with_grad, outdir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, _, lossm = ref_loader.load()
    img, txt = synth.make_features(n, e, seed=n + e, dtype=torch.float64)
    nl = n // world
    img_l = img[rank * nl:(rank + 1) * nl].clone().requires_grad_(True)
    txt_l = txt[rank * nl:(rank + 1) * nl].clone().requires_grad_(True)
    ls = torch.tensor(np.log(scale), dtype=torch.float64, requires_grad=True)
    crit = lossm.ClipLoss(local_loss=local_loss, gather_with_grad=gather_with_grad, rank=rank, world_size=world)
    loss = crit(img_l, txt_l, ls.exp())
    loss.backward()
    torch.save(dict(loss=loss.detach(), d_img=img_l.grad, d_txt=txt_l.grad, d_ls=ls.grad),
               os.path.join(outdir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def dist_case(n, e, scale, world, local_loss, gather_with_grad, port):
    import torch.multiprocessing as mp
    os.environ["MASTER_PORT"] = str(port)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_dist_worker, args=(world, n, e, scale, local_loss, gather_with_grad, d), nprocs=world, join=True)
        res = [torch.load(os.path.join(d, f"r{r}.pt")) for r in range(world)]
    out = dict(scale=np.float64(scale), world=np.int64(world), seed=np.int64(n + e))
    for r, x in enumerate(res):
        out[f"loss_r{r}"] = np.float64(x["loss"])
        out[f"d_img_r{r}"] = x["d_img"].numpy()
        out[f"d_txt_r{r}"] = x["d_txt"].numpy()
        out[f"d_logit_scale_r{r}"] = np.float64(x["d_ls"])
    tag = f"W{world}_N{n}_E{e}_local{int(local_loss)}_gwg{int(gather_with_grad)}"
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"loss_dist_{tag}.npz"), **out)
    print(f"[dist {tag}] per-rank losses {[float(x['loss']) for x in res]}")


class _Grid:
    """the two attributes model.py:523-592 read from the model"""

    def __init__(self, grid, ctx, width):
        self.visual = type("V", (), {})()
        self.visual.grid_size = grid
        self.positional_embedding = torch.zeros(ctx, width)


POS_CASES = [(16, 24, "bicubic", True), (24, 16, "bicubic", True), (16, 27, "bicubic", False), (14, 10, "bilinear", False)]
TEXT_POS_CASES = [(77, 80, "linear", False), (80, 32, "linear", False)]


def pos_embed_case(mdl):
    """the reference's checkpoint-load resampling of positional embeddings (model.py:523-592) on seeded tables"""
    out = {}
    for (og, ng, mode, aa) in POS_CASES:
        g = torch.Generator().manual_seed(1000 + og * 37 + ng)
        table = torch.randn(og * og + 1, 48, generator=g, dtype=torch.float64)
        sd = {"visual.positional_embedding": table.clone()}
        mdl.resize_pos_embed(sd, _Grid((ng, ng), 8, 48), interpolation=mode, antialias=aa)
        out[f"img_{og}_{ng}_{mode}_{int(aa)}"] = sd["visual.positional_embedding"].numpy()
    for (oc, nc, mode, aa) in TEXT_POS_CASES:
        g = torch.Generator().manual_seed(2000 + oc * 37 + nc)
        table = torch.randn(oc, 40, generator=g, dtype=torch.float64)
        sd = {"positional_embedding": table.clone()}
        mdl.resize_text_pos_embed(sd, _Grid((4, 4), nc, 40), interpolation=mode, antialias=aa)
        out[f"txt_{oc}_{nc}_{mode}_{int(aa)}"] = sd["positional_embedding"].numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "pos_embed_resize.npz"), **out)
    print(f"  wrote pos_embed_resize.npz ({len(out)} arrays)")


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    tr, mdl, lossm = ref_loader.load()
    pos_embed_case(mdl)
    if "--only-pos" in sys.argv:
        return
    for cfg_name, batch in TOWER_CASES:
        tower_case(tr, mdl, lossm, cfg_name, batch)
    for n, e, s in LOSS_CASES:
        loss_case(lossm, n, e, s)
    port = 29540
    for world in (2, 4):
        for ll, gwg in ((True, True), (False, False), (False, True)):
            dist_case(64, 32, 1.0 / 0.07, world, ll, gwg, port)
            port += 1
    print("golden vectors written to", GOLDEN_DIR)


if __name__ == "__main__":
    main()

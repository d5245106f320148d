"""Debug: per-parameter gradient agreement between the fused block path and the module-by-module (hooked) path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import openvision_b200 as ovb
from oracle import synth

cfg_name = "mini-ov"
cfg = synth.CONFIGS[cfg_name]
m = ovb.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]))
m.load_state_dict(synth.make_state_dict(cfg_name, 0), strict=True)
m = m.cuda().eval()
images = synth.make_images(cfg_name, 4, 0).cuda()
wgt = torch.randn(4, 64, generator=torch.Generator().manual_seed(3)).cuda()


def run(hook_on=None, kind="fwd"):
    m.zero_grad(set_to_none=True)
    x = images.clone().requires_grad_(True)
    hs = []
    if hook_on is not None:
        for mod in hook_on:
            if kind == "fwd":
                hs.append(mod.register_forward_hook(lambda *a: None))
            else:
                hs.append(mod.register_full_backward_hook(lambda *a: None))
    out = m.encode_image(x, normalize=True)
    (out.float() * wgt).sum().backward()
    for h in hs:
        h.remove()
    g = {n: p.grad.detach().float().clone() for n, p in m.visual.named_parameters() if p.grad is not None}
    g["__image__"] = x.grad.detach().float().clone()
    g["__out__"] = out.detach().float().clone()
    return g


ref = run()
again = run()
print("fused vs fused image-grad rel:", float((again["__image__"] - ref["__image__"]).norm() / ref["__image__"].norm()))
blk0, blk1 = m.visual.transformer.resblocks
cases = {"fwd:gelu1": ([blk1.mlp.gelu], "fwd"), "fwd:c_fc1": ([blk1.mlp.c_fc], "fwd"), "fwd:ln_1_1": ([blk1.ln_1], "fwd"),
         "fwd:attn1": ([blk1.attn], "fwd"), "fwd:blk1": ([blk1], "fwd"), "bwd:c_fc1": ([blk1.mlp.c_fc], "bwd"),
         "bwd:blk1": ([blk1], "bwd"), "fwd:c_fc0": ([blk0.mlp.c_fc], "fwd"), "fwd:transformer": ([m.visual.transformer], "fwd"),
         "fwd:ln_post": ([m.visual.ln_post], "fwd")}
for name, (mods, kind) in cases.items():
    g = run(mods, kind)
    worst = max(((float((g[k] - ref[k]).norm() / (ref[k].norm() + 1e-12)), k) for k in ref if k in g), key=lambda t: t[0])
    missing = [k for k in ref if k not in g]
    print(f"{name:18s} image rel {float((g['__image__'] - ref['__image__']).norm() / ref['__image__'].norm()):.4f}  "
          f"out rel {float((g['__out__'] - ref['__out__']).norm() / ref['__out__'].norm()):.4f}  worst {worst[0]:.4f} {worst[1]}  missing {missing[:3]}")

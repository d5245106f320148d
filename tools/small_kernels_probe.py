"""Stand-alone timings (burst clocks) of the patch-embedding and pooling-head kernels at the headline shapes, and the
attention4 truncating-pack A/B.  usage (GPU box): python tools/small_kernels_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvision_b200 import ops


def t(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


for (B, H, P, D, dt) in ((1024, 224, 14, 1024, torch.float32), (1024, 224, 14, 1280, torch.float32), (512, 384, 16, 768, torch.float32),
                         (1024, 224, 14, 1024, torch.bfloat16)):
    img = torch.randn(B, 3, H, H, device="cuda").to(dt)
    w = torch.randn(D, 3, P, P, device="cuda") * 0.05
    N = (H // P) ** 2
    table = torch.randn(N + 1, D, device="cuda").bfloat16()
    wk = ops.pack_patch_weight(w, P)
    ms = t(lambda: ops.patch_embed(img, wk, P, table))
    K = 3 * P * P
    kpad = (K + 7) // 8 * 8
    wp = torch.zeros(D, kpad, dtype=torch.bfloat16, device="cuda")
    wp[:, :K] = w.reshape(D, K).bfloat16()
    ms_old = t(lambda: ops.gemm_rowadd(ops.im2col_patches(img, P, kpad, lead_rows=1), wp, table))
    byts = img.numel() * img.element_size() + B * (N + 1) * D * 2
    print(f"patch_embed B{B} {H}px P{P} D{D} {str(dt)[6:]}: {ms:.3f} ms  {2.0 * B * N * D * K / ms / 1e9:.0f} TF/s  {byts / ms / 1e6:.0f} GB/s "
          f"(im2col + GEMM: {ms_old:.3f} ms)", flush=True)
    del img

for (B, L, D, E) in ((1024, 257, 1024, 768), (1024, 257, 1280, 1024), (512, 577, 768, 512), (8, 257, 1024, 768)):
    x = torch.randn(B, L, D, device="cuda").bfloat16()
    g, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    proj = (torch.randn(D, E, device="cuda") * D ** -0.5).bfloat16()
    ms = t(lambda: ops.pool_head(x, "avg", g, b, 1e-6, proj, normalize=True, out_dtype=torch.float32))

    def old():
        p = ops.pool_tokens(x, "avg")
        p = ops.layernorm(p, g, b, 1e-6)
        p = ops.gemm(p, proj.t().contiguous() if False else projT)
        return ops.l2_normalize(p)
    projT = proj.t().contiguous()
    ms_old = t(old)
    print(f"pool_head B{B} L{L} D{D} E{E}: {ms * 1e3:.1f} us  {x.numel() * 2 / ms / 1e6:.0f} GB/s  (4 kernels: {ms_old * 1e3:.1f} us)", flush=True)

for (B, H, L) in ((1024, 16, 257), (1024, 16, 256), (256, 16, 1025)):
    qkv = torch.randn(B * L, 3 * H * 64, device="cuda").bfloat16()
    fl = 4.0 * B * H * L * L * 64
    res = {}
    for rep in range(2):
        for name, tr in (("rn", "0"), ("trunc", "1")):
            os.environ["OVK_ATT4_TRUNC"] = tr
            res.setdefault(name, []).append(t(lambda: ops.attention(qkv, B, L, H, 64)))
    os.environ["OVK_ATT4_TRUNC"] = "0"
    print(f"attention4 B{B} H{H} L{L}: " + "  ".join(f"{k} {min(v):.3f} ms {fl / min(v) / 1e9:.0f} TF/s" for k, v in res.items()), flush=True)

for (rows, D) in ((263168, 1024), (263168, 1280), (295424, 768)):
    x = torch.randn(rows, D, device="cuda").bfloat16()
    g, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    ms = t(lambda: ops.layernorm(x, g, b, 1e-6))
    y, mean, rstd = ops.layernorm(x, g, b, 1e-6, save_stats=True)
    dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    dy = torch.randn_like(x)
    msb = t(lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, dg, db, dres=x))
    print(f"layernorm rows {rows} D {D}: fwd {ms * 1e3:.0f} us {4.0 * rows * D / ms / 1e6:.0f} GB/s   bwd(+res) {msb * 1e3:.0f} us {8.0 * rows * D / msb / 1e6:.0f} GB/s", flush=True)

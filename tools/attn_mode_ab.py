"""A/B: shape-specialised instances of attention4 vs the generic one (OVK_ATT4_GENERIC=1).  usage (GPU box): python tools/attn_mode_ab.py"""
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


for (B, H, L) in ((1024, 16, 257), (1024, 16, 256), (256, 16, 513), (256, 16, 1025), (256, 16, 512)):
    qkv = (torch.randn(B * L, 3 * H * 64, device="cuda") * 0.5).bfloat16()
    fl = 4.0 * B * H * L * L * 64
    res = {}
    for rep in range(3):
        for g in ("0", "1"):
            os.environ["OVK_ATT4_GENERIC"] = g
            res.setdefault("specialised" if g == "0" else "generic", []).append(t(lambda: ops.attention(qkv, B, L, H, 64)))
    print(f"B{B} H{H} L{L}: " + "  ".join(f"{k} {min(v):.3f} ms {fl / min(v) / 1e9:.0f} TF/s" for k, v in res.items()), flush=True)
os.environ.pop("OVK_ATT4_GENERIC", None)

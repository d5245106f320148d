"""A/B: pair kernel (attention4) forced onto an odd number of query tiles vs the one-tile-per-CTA kernel (attention.cu).
usage (GPU box): python tools/attn_odd_probe.py"""
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


for (B, H, L) in ((512, 12, 577), (256, 16, 385), (1024, 16, 129), (256, 16, 641), (512, 12, 576), (64, 12, 1601)):
    qkv = torch.randn(B * L, 3 * H * 64, device="cuda").bfloat16()
    fl = 4.0 * B * H * L * L * 64
    res = {}
    for rep in range(2):
        for name, odd in (("v1", "0"), ("v4odd", "1")):
            os.environ["OVK_ATT_PAIR_ODD"] = odd
            res.setdefault(name, []).append(t(lambda: ops.attention(qkv, B, L, H, 64)))
    os.environ["OVK_ATT_PAIR_ODD"] = "0"
    print(f"B{B} H{H} L{L}: " + "  ".join(f"{k} {min(v):.3f} ms {fl / min(v) / 1e9:.0f} TF/s" for k, v in res.items()), flush=True)

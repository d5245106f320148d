"""Attention forward A/B: attention3 (default) vs the round-1 pair kernel (OVK_ATT_V2=1) vs SDPA, burst timing."""
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


for (B, H, L, hd) in ((1024, 16, 257, 64), (1024, 16, 256, 64), (256, 16, 513, 64), (256, 16, 1025, 64), (2048, 12, 80, 64), (64, 16, 257, 64)):
    qkv = torch.randn(B * L, 3 * H * hd, device="cuda").bfloat16()
    fl = 4.0 * B * H * L * L * hd
    res = {}
    for rep in range(2):
        for name, ver, mode in (("v4", "4", "0"), ("v3", "3", "0"), ("v3dual", "3", "1"), ("v2", "2", "0")):
            os.environ["OVK_ATT_VER"] = ver
            os.environ["OVK_ATT3_MODE"] = mode
            ms = t(lambda: ops.attention(qkv, B, L, H, hd))
            res.setdefault(name, []).append(ms)
    os.environ["OVK_ATT_VER"] = "4"
    os.environ["OVK_ATT3_MODE"] = "0"
    print(f"B{B} H{H} L{L}: " + "  ".join(f"{k} {min(v):.3f} ms {fl / min(v) / 1e9:.0f} TF/s" for k, v in res.items()), flush=True)

"""A/B: start-up stagger of tile slot B in attention4 (OVK_ATT4_STAGGER_NS).  usage (GPU box): python tools/attn_stagger_ab.py"""
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


for (B, H, L) in ((1024, 16, 257), (1024, 16, 256), (256, 16, 1025)):
    qkv = (torch.randn(B * L, 3 * H * 64, device="cuda") * 0.5).bfloat16()
    res = {}
    for rep in range(2):
        for ns in ("0", "150", "300", "450", "600", "900", "1500"):
            os.environ["OVK_ATT4_STAGGER_NS"] = ns
            res.setdefault(ns, []).append(t(lambda: ops.attention(qkv, B, L, H, 64)))
    print(f"B{B} H{H} L{L}: " + "  ".join(f"{k}ns {min(v):.3f}" for k, v in res.items()), flush=True)
os.environ.pop("OVK_ATT4_STAGGER_NS", None)

"""Where a softmax step of attention4 spends its time: run the kernel with parts knocked out (results are wrong on purpose).
Needs libovk built with -DOVK_ATT4_DEBUG_VARIANTS (make NVCCFLAGS_EXTRA=-DOVK_ATT4_DEBUG_VARIANTS).  usage: python tools/attn_knockout.py"""
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


names = {64: "no HBM (L2-resident operands)", 79: "skeleton, no HBM", 127: "barriers only, no HBM", 16: "no epilogue", 31: "skeleton, no epilogue", 47: "skeleton, no main MMAs", 63: "barriers + TMA only", 0: "full", 1: "no exp/pack", 2: "no max", 4: "no P store", 8: "no S load", 5: "no exp, no store", 7: "no exp/max/store", 15: "nothing but barriers + MMA"}
for (B, H, L) in ((1024, 16, 257), (1024, 16, 256), (256, 16, 1025)):
    qkv = (torch.randn(B * L, 3 * H * 64, device="cuda") * 0.5).bfloat16()
    out = []
    for d in (0, 64, 1, 16, 15, 79, 31, 47, 63, 127, 0):
        os.environ["OVK_ATT4_DBG"] = str(d)
        out.append(f"{names[d]} {t(lambda: ops.attention(qkv, B, L, H, 64)):.3f}")
    print(f"B{B} H{H} L{L} (ms): " + " | ".join(out), flush=True)
os.environ["OVK_ATT4_DBG"] = "0"

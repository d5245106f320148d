"""Knock-out timings of attention_bwd_t_kernel (results wrong on purpose): which part of an iteration bounds the kernel.
bits: 1 no TMA reduce-add, 2 no dQ drain at all, 4 no softmax work (TMEM loads, exponentials, P / dS stores), 8 no dQ MMA, 16 spinning instead of suspending barrier waits,
32 per-wait / per-issue-site cycle counters printed by block 0, 64 no dV / dK MMAs, 128 no score MMAs.
Needs libovk built with the debug hooks: make -C openvision_b200/csrc clean all NVCCFLAGS_EXTRA=-DOVK_ATTBWD_DEBUG
(the shipped library compiles them out; the bits are then ignored).
usage (GPU box): python tools/attn_bwd_knockout.py [B H L hd]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvision_b200 import ops

B, H, L, hd = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (512, 12, 577, 64)
qkv = (torch.randn(B * L, 3 * H * hd, device="cuda") * 0.5).bfloat16()
dout = torch.randn(B * L, H * hd, device="cuda").bfloat16()
out, lse = ops.attention(qkv, B, L, H, hd, save_lse=True)


def t(iters=10):
    if int(os.environ.get('OVK_ATTBWD_DBG', '0')) & 32:
        iters = 1
    for _ in range(2):
        ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


os.environ["OVK_ATTBWD_FUSED"] = "2"
for bits in (32, 32 | 6, 32 | 14):
    os.environ["OVK_ATTBWD_DBG"] = str(bits)
    print(f"B{B} H{H} L{L} hd{hd} knock-out {bits:2d}: {t():.3f} ms (whole call: delta + tile kernel + conversion)", flush=True)

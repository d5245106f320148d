"""Timings of the attention kernels that the training workloads use: forward of the one-tile kernel (L = 577, head width 80)
and the two backward kernels.  usage (GPU box): python tools/attn_bwd_ab.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvision_b200 import ops


def t(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


for (B, H, L, hd) in ((256, 16, 257, 64), (1024, 16, 257, 64), (512, 12, 577, 64), (256, 16, 257, 80), (1024, 16, 257, 80), (256, 16, 256, 64), (64, 16, 1025, 64)):
    qkv = (torch.randn(B * L, 3 * H * hd, device="cuda") * 0.5).bfloat16()
    dout = torch.randn(B * L, H * hd, device="cuda").bfloat16()
    out, lse = ops.attention(qkv, B, L, H, hd, save_lse=True)
    f = t(lambda: ops.attention(qkv, B, L, H, hd, save_lse=True))
    os.environ["OVK_ATTBWD_FUSED"] = "0"
    os.environ["OVK_ATTBWD_TAIL"] = "0"
    b0 = t(lambda: ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd))
    os.environ["OVK_ATTBWD_TAIL"] = "1"
    b = t(lambda: ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd))
    os.environ["OVK_ATTBWD_FUSED"] = "1"
    bf1 = t(lambda: ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd))
    os.environ["OVK_ATTBWD_FUSED"] = "2"
    os.environ["OVK_ATTBWD_WARPS"] = "16"
    bf16w = t(lambda: ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd))
    os.environ["OVK_ATTBWD_WARPS"] = "8"
    bf = t(lambda: ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd))
    fl = 4.0 * B * H * L * L * hd
    print(f"B{B} H{H} L{L} hd{hd}: fwd {f:.3f} ms {fl / f / 1e9:.0f} TF/s   bwd one-pass {bf:.3f} ms {2.5 * fl / bf / 1e9:.0f} TF/s (10 B H L^2 hd)   [16 compute warps {bf16w:.3f} ms, v1 {bf1:.3f} ms]   "
          f"two-pass {b:.3f} ms {3.5 * fl / b / 1e9:.0f} TF/s (14 B H L^2 hd)   [two-pass, remainder token as tiles: {b0:.3f} ms]", flush=True)

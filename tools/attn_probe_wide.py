"""Attention forward / backward timing at head width 80 (ViT-H/14) and 72 (So400m), L = 257."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openvision_b200 import ops
B, H, L = int(os.environ.get("B", 256)), 16, 257
def t(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for hd in (80, 72, 64):
    qkv = torch.randn(B * L, 3 * H * hd, device="cuda").bfloat16()
    ms = t(lambda: ops.attention(qkv, B, L, H, hd))
    out, lse = ops.attention(qkv, B, L, H, hd, save_lse=True)
    do = torch.randn_like(out)
    msb = t(lambda: ops.attention_bwd(qkv, out, do, lse, B, L, H, hd))
    print(f"hd={hd} L={L}: fwd {ms:.3f} ms ({4*B*H*L*L*hd/ms/1e9:.0f} TF/s)  bwd {msb:.3f} ms ({10*B*H*L*L*hd/msb/1e9:.0f} TF/s alg)", flush=True)

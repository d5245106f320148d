import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openvision_b200 import ops
B, H = int(os.environ.get("B", 256)), 16
def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for L in (128, 256, 257, 264, 384, 577):
    qkv = torch.randn(B * L, 3 * H * 64, device="cuda").bfloat16()
    ms = t(lambda: ops.attention(qkv, B, L, H, 64))
    out, lse = ops.attention(qkv, B, L, H, 64, save_lse=True)
    do = torch.randn_like(out)
    msb = t(lambda: ops.attention_bwd(qkv, out, do, lse, B, L, H, 64), 10)
    print(f"L={L}: fwd {ms:.3f} ms ({4*B*H*L*L*64/ms/1e9:.0f} TF/s)  bwd {msb:.3f} ms ({10*B*H*L*L*64/msb/1e9:.0f} TF/s alg)", flush=True)

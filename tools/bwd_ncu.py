"""A few attention-backward launches for ncu (B/16@384 shape: L = 577, 12 heads)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openvision_b200 import ops
B, H, L = int(os.environ.get("B", 128)), 12, int(os.environ.get("L", 577))
qkv = torch.randn(B * L, 3 * H * 64, device="cuda").bfloat16()
out, lse = ops.attention(qkv, B, L, H, 64, save_lse=True)
do = torch.randn_like(out)
for _ in range(3):
    ops.attention_bwd(qkv, out, do, lse, B, L, H, 64)
torch.cuda.synchronize()

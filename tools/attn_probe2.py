import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openvision_b200 import ops
B, H = 256, 16
for L in (256, 257, 272):
    qkv = torch.randn(B * L, 3 * H * 64, device="cuda").bfloat16()
    for _ in range(2):
        out, lse = ops.attention(qkv, B, L, H, 64, save_lse=True)
    torch.cuda.synchronize()
print("done")

import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openvision_b200 import ops
B, H, L = int(os.environ.get("B", 256)), 16, int(os.environ.get("L", 257))
qkv = torch.randn(B * L, 3 * H * 64, device="cuda").bfloat16()
for _ in range(4):
    ops.attention(qkv, B, L, H, 64)
torch.cuda.synchronize()

"""One call of the attention backward per mode for an ncu launch list (per-kernel durations of the one-pass path).
usage (GPU box): ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/attn_bwd_ncu.py [B H L hd]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvision_b200 import ops

B, H, L, hd = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (512, 12, 577, 64)
qkv = (torch.randn(B * L, 3 * H * hd, device="cuda") * 0.5).bfloat16()
dout = torch.randn(B * L, H * hd, device="cuda").bfloat16()
out, lse = ops.attention(qkv, B, L, H, hd, save_lse=True)
for mode in ("2", "1", "0"):
    os.environ["OVK_ATTBWD_FUSED"] = mode
    for _ in range(2):
        ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd)
    torch.cuda.synchronize()

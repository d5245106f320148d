"""Race hunt for the one-pass attention backward: the same call repeated, results must be BITWISE identical from run to run
(every reduction in the kernels has a fixed order) and close to the two-pass kernels.  usage (GPU box): python tools/attn_bwd_stress.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvision_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
bad = 0
for (B, H, L, hd) in ((512, 12, 577, 64), (1024, 16, 257, 64), (256, 16, 257, 80), (64, 16, 1025, 64), (37, 5, 300, 72), (200, 16, 77, 64)):
    qkv = (torch.randn(B * L, 3 * H * hd, device="cuda") * 0.5).bfloat16()
    dout = torch.randn(B * L, H * hd, device="cuda").bfloat16()
    out, lse = ops.attention(qkv, B, L, H, hd, save_lse=True)
    os.environ["OVK_ATTBWD_FUSED"] = "0"
    ref = ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd).float()
    os.environ["OVK_ATTBWD_FUSED"] = "2"
    first = ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd)
    rel = float((first.float() - ref).norm() / ref.norm())
    diffs = 0
    for _ in range(reps):
        again = ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd)
        diffs += int(not torch.equal(again, first))
    torch.cuda.synchronize()
    print(f"B{B} H{H} L{L} hd{hd}: one-pass vs two-pass rel L2 {rel:.2e}; {diffs} of {reps} repeats differ bitwise", flush=True)
    bad += diffs + int(rel > 2e-2)
sys.exit(1 if bad else 0)

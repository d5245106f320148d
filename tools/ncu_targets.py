"""One launch of each kernel that the bench's ncu pass does not capture at full size: patch embedding, pooling head, the loss
kernels at N = 32 768, attention backward (tail + dQ + dK/dV) and forward at the headline shape.  Run under
  ncu --set full --clock-control none --import-source on -k regex:"patch_embed|pool_head|clip_loss|attention" -o ...
usage (GPU box): python tools/ncu_targets.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvision_b200 import ops

torch.manual_seed(0)
B, H_, P, D = 1024, 224, 14, 1024
img = torch.randn(B, 3, H_, H_, device="cuda")
w = torch.randn(D, 3, P, P, device="cuda") * 0.05
N = (H_ // P) ** 2
table = torch.randn(N + 1, D, device="cuda").bfloat16()
tok = ops.patch_embed(img, ops.pack_patch_weight(w, P), P, table)
g, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
proj = (torch.randn(D, 768, device="cuda") * D ** -0.5).bfloat16()
feat = ops.pool_head(tok, "avg", g, b, 1e-6, proj, normalize=True, out_dtype=torch.float32)
del img, tok
n, e = 32768, 768
fi = torch.nn.functional.normalize(torch.randn(n, e, device="cuda"), dim=-1).bfloat16()
ft = torch.nn.functional.normalize(torch.randn(n, e, device="cuda"), dim=-1).bfloat16()
scale = torch.tensor([14.2857], device="cuda")
row_lse, diag, col_max, col_sum = ops.clip_loss_fwd(fi, ft, 0, scale)
col_lse = ops.clip_loss_combine(col_max.unsqueeze(0), col_sum.unsqueeze(0))
ds = torch.zeros(1, device="cuda")
G = ops.clip_loss_grad_logits(fi, ft, 0, scale, row_lse, col_lse, 0.5 / n, 0.5 / n, ds)
del G
Bq, H, L, hd = 256, 16, 257, 64
qkv = (torch.randn(Bq * L, 3 * H * hd, device="cuda") * 0.5).bfloat16()
out, lse = ops.attention(qkv, Bq, L, H, hd, save_lse=True)
dqkv = ops.attention_bwd(qkv, out, torch.randn_like(out), lse, Bq, L, H, hd)
torch.cuda.synchronize()
print("ok")

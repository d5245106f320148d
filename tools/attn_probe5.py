import sys, os, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openvision_b200 import ops
L, B, H = 130, 1, 1
torch.manual_seed(L)
qkv = torch.randn(B * L, 3 * H * 64, device="cuda").bfloat16()
out, lse = ops.attention(qkv, B, L, H, 64, save_lse=True)
out = out.float()
q, k, v = qkv.float().view(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
s = (q @ k.transpose(-1, -2)) / 8.0
def att(sl):
    p = torch.softmax(s[..., sl], -1)
    return (p @ v[..., sl, :]).permute(0, 2, 1, 3).reshape(B * L, 64)
full = att(slice(0, 130)); first = att(slice(0, 128))
s2 = s.clone(); 
for r in (3, 5, 6, 10, 0, 1, 2):
    e_full = (out[r] - full[r]).abs().max().item(); e_first = (out[r] - first[r]).abs().max().item()
    srow = s[0, 0, r]
    print(r, "err vs full %.4f  vs first128-only %.4f" % (e_full, e_first), "max0 %.3f  tail scores %s  lse got %.4f ref %.4f" % (srow[:128].max().item(), srow[128:].tolist(), lse[0,0,r].item(), torch.logsumexp(srow, -1).item()))

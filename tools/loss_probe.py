import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openvision_b200 as ovb
from openvision_b200 import ops
n, e = int(os.environ.get("N", 32768)), 768
g = torch.Generator(device="cuda").manual_seed(0)
img = torch.nn.functional.normalize(torch.randn(n, e, device="cuda", generator=g), dim=-1).bfloat16().requires_grad_(True)
txt = torch.nn.functional.normalize(torch.randn(n, e, device="cuda", generator=g), dim=-1).bfloat16().requires_grad_(True)
ls = torch.tensor(2.6592, device="cuda", requires_grad=True)
crit = ovb.ClipLoss()
def step():
    img.grad = txt.grad = ls.grad = None
    loss = crit(img, txt, ls.exp()); loss.backward(); return loss
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
print(f"N={n}: fwd+bwd {e0.elapsed_time(e1)/10:.3f} ms")
rec = {}
class R:
    def __init__(s, k, w): s.k, s.w = k, w
    def __enter__(s):
        s.a, s.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); s.a.record()
    def __exit__(s, *a):
        s.b.record(); rec.setdefault(s.k, []).append((s.a, s.b, s.w)); return False
ops.recorder = lambda k, w: R(k, w)
step(); torch.cuda.synchronize()
for k, l in rec.items():
    for a, b, w in l:
        print(f"   {k:16s} {a.elapsed_time(b):7.3f} ms  {w/a.elapsed_time(b)/1e9:8.1f} TFLOP/s")

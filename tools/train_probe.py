"""fwd / fwd+bwd timing of the image tower at BASELINE configs (python tools/train_probe.py CFG BATCH [ckpt])."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openvision_b200 as ovb
from openvision_b200 import ops
from oracle import synth
cfg_name, batch = sys.argv[1], int(sys.argv[2])
ckpt = len(sys.argv) > 3 and sys.argv[3] == "ckpt"
cfg = synth.CONFIGS[cfg_name]
torch.manual_seed(0)
v = ovb.model._build_vision_tower(cfg["embed_dim"], cfg["vision"]).cuda().train()
v.set_grad_checkpointing(ckpt)
s = cfg["vision"]["image_size"]
images = torch.randn(batch, 3, s, s, device="cuda")
gy = torch.randn(batch, cfg["embed_dim"], device="cuda")
flops = {"L14-224": 162.03e9, "B16-384": 110.97e9, "Ti16-160": 1.2e9, "H14-224": 334.59e9}.get(cfg_name, 0) * batch
def ev(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
def fwd():
    with torch.no_grad(): return v(images)
def fwdbwd():
    for p in v.parameters(): p.grad = None
    out = v(images)
    out.backward(gy)
ms_f = ev(fwd)
torch.cuda.reset_peak_memory_stats()
ms_fb = ev(fwdbwd)
print(f"{cfg_name} batch {batch} ckpt={ckpt}: fwd {ms_f:.1f} ms ({flops/ms_f/1e9:.0f} TFLOP/s, {batch/ms_f*1e3:.0f} img/s)  "
      f"fwd+bwd {ms_fb:.1f} ms ({3*flops/ms_fb/1e9:.0f} TFLOP/s alg, {batch/ms_fb*1e3:.0f} img/s)  "
      f"peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
rec = {}
class R:
    def __init__(s, k, w): s.k, s.w = k, w
    def __enter__(s):
        s.a, s.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); s.a.record()
    def __exit__(s, *a):
        s.b.record(); rec.setdefault(s.k, []).append((s.a, s.b, s.w)); return False
ops.recorder = lambda k, w: R(k, w)
fwdbwd(); torch.cuda.synchronize()
for k, l in sorted(rec.items()):
    ms = sum(a.elapsed_time(b) for a, b, _ in l); w = sum(x for _, _, x in l)
    print(f"   {k:16s} {len(l):4d} launches {ms:8.2f} ms  {w/ms/1e9:8.1f} {'TFLOP/s' if k in ('gemm','attention','attention_bwd') else 'GB/s*1e3'}")

# per-kind device time of one fwd+bwd (CUDA events around every libovk launch)
class _Rec:
    def __init__(self): self.r = {}
    def __call__(self, kind, work):
        rec = self
        class B:
            def __enter__(s):
                s.a, s.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); s.a.record()
            def __exit__(s, *x):
                s.b.record(); rec.r.setdefault(kind, []).append((s.a, s.b)); return False
        return B()
rec = _Rec(); ops.recorder = rec
fwdbwd(); torch.cuda.synchronize(); ops.recorder = None
parts = {k: sum(a.elapsed_time(b) for a, b in e) for k, e in rec.r.items()}
print("per-kind ms in one fwd+bwd: " + "  ".join(f"{k}={t:.1f}({len(rec.r[k])})" for k, t in sorted(parts.items(), key=lambda kv: -kv[1])), "  sum=%.1f" % sum(parts.values()), flush=True)

"""A/B of environment switches on the ViT-L/14@224 tower forward, alternating the arms inside one process so that
clock / thermal drift hits both equally.  usage: ab_probe.py VAR=a,b [batch] ; prints the median ms per arm."""
import os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openvision_b200 as ov
from oracle import synth

var, vals = sys.argv[1].split("=")
vals = vals.split(",")
Bt = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cfg = synth.CONFIGS[os.environ.get("OVK_AB_CONFIG", "L14-224")]
vt = ov.model._build_vision_tower(cfg["embed_dim"], cfg["vision"]).cuda().eval()
size = cfg["vision"]["image_size"]
imgs = torch.randn(Bt, 3, size, size, device="cuda")
times = {v: [] for v in vals}
with torch.no_grad():
    for v in vals:
        os.environ[var] = v
        for _ in range(3):
            vt(imgs)
    for rnd in range(6):
        for v in vals:
            os.environ[var] = v
            vt(imgs)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                vt(imgs)
            e1.record()
            torch.cuda.synchronize()
            times[v].append(e0.elapsed_time(e1) / 5)
for v in vals:
    ms = statistics.median(times[v])
    print(f"[ab B={Bt}] {var}={v}: median {ms:.3f} ms  {Bt / ms * 1e3:.0f} img/s   all: {' '.join(f'{t:.2f}' for t in times[v])}", flush=True)

# per-kind device time inside the tower (CUDA events around every libovk launch), one pass per arm
from openvision_b200 import ops as _ops


class _Rec:
    def __init__(self):
        self.r = {}

    def __call__(self, kind, work):
        rec = self

        class B:
            def __enter__(s):
                s.a, s.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.a.record()

            def __exit__(s, *x):
                s.b.record()
                rec.r.setdefault(kind, []).append((s.a, s.b))
                return False
        return B()


with torch.no_grad():
    for v in vals:
        os.environ[var] = v
        for _ in range(2):
            vt(imgs)
        rec = _Rec()
        _ops.recorder = rec
        for _ in range(3):
            vt(imgs)
        torch.cuda.synchronize()
        _ops.recorder = None
        parts = {k: sum(a.elapsed_time(b) for a, b in e) / 3 for k, e in rec.r.items()}
        print(f"[ab B={Bt}] {var}={v} per-kind ms/forward: " + "  ".join(f"{k}={t:.2f}({len(rec.r[k]) // 3})" for k, t in sorted(parts.items())),
              flush=True)

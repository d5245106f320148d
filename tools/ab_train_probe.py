"""Alternating A/B of an environment switch on a tower fwd+bwd step (usage: ab_train_probe.py VAR=a,b CFG BATCH [ckpt])."""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openvision_b200 as ovb
from oracle import synth
var, vals = sys.argv[1].split("=")
vals = vals.split(",")
cfg_name, batch = sys.argv[2], int(sys.argv[3])
ckpt = sys.argv[4] if len(sys.argv) > 4 else "none"
cfg = synth.CONFIGS[cfg_name]
torch.manual_seed(0)
v = ovb.model._build_vision_tower(cfg["embed_dim"], cfg["vision"]).cuda().train()
v.set_grad_checkpointing("mlp" if ckpt == "mlp" else ckpt == "block")
s = cfg["vision"]["image_size"]
images = torch.randn(batch, 3, s, s, device="cuda")
gy = torch.randn(batch, cfg["embed_dim"], device="cuda")
def step():
    for p in v.parameters(): p.grad = None
    v(images).backward(gy)
times = {x: [] for x in vals}
for x in vals:
    os.environ[var] = x
    step(); step()
for rnd in range(4):
    for x in vals:
        os.environ[var] = x
        step(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): step()
        e1.record(); torch.cuda.synchronize()
        times[x].append(e0.elapsed_time(e1) / 3)
for x in vals:
    ms = statistics.median(times[x])
    print(f"[ab-train {cfg_name} B={batch} ckpt={ckpt}] {var}={x}: median {ms:.2f} ms   all: {' '.join(f'{t:.1f}' for t in times[x])}", flush=True)

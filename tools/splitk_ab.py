"""Alternating A/B of the split-K policy on the weight-gradient GEMM shapes (OVK_SPLITK_SIMPLE=1 vs default)."""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openvision_b200 import ops
cfgs = {"B16": (512 * 577, 768, 3072), "L14": (1024 * 257, 1024, 4096), "H14": (512 * 257, 1280, 5120)}
for name, (M, D, F) in cfgs.items():
    x = torch.randn(M, D, device="cuda").bfloat16()
    f = torch.randn(M, F, device="cuda").bfloat16()
    tot = {"0": 0.0, "1": 0.0}
    for nm, N, b_in in (("qkv", 3 * D, x), ("out", D, x), ("fc1", F, x), ("fc2", D, f)):
        dy = torch.randn(M, N, device="cuda").bfloat16()
        times = {"0": [], "1": []}
        for rnd in range(5):
            for pol in ("0", "1"):
                os.environ["OVK_SPLITK_SIMPLE"] = pol
                for _ in range(2): ops.gemm_tn(dy, b_in, out_dtype=torch.float32)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5): ops.gemm_tn(dy, b_in, out_dtype=torch.float32)
                e1.record(); torch.cuda.synchronize()
                times[pol].append(e0.elapsed_time(e1) / 5)
        m0, m1 = statistics.median(times["0"]), statistics.median(times["1"])
        tot["0"] += m0; tot["1"] += m1
        fl = 2.0 * M * N * b_in.shape[1]
        print(f"{name} {nm:4s} wave-aware {m0:7.3f} ms {fl/m0/1e9:7.1f} TF/s | simple {m1:7.3f} ms {fl/m1/1e9:7.1f} TF/s", flush=True)
        del dy
    print(f"{name} sum: wave-aware {tot['0']:.3f} ms, simple {tot['1']:.3f} ms", flush=True)
    del x, f

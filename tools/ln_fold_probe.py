"""Times the LayerNorm-folded GEMMs against the plain ones at ViT-L/14 shapes (device events, L2 flushed between)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openvision_b200 import ops

B = int(os.environ.get("OVK_PERF_BATCH", "256"))
M, D = B * 257, 1024
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(M, D, device=dev, generator=g).bfloat16()
f = torch.randn(M, 4 * D, device=dev, generator=g).bfloat16()
wq = (torch.randn(3 * D, D, device=dev, generator=g) * 0.02).bfloat16()
wo = (torch.randn(D, D, device=dev, generator=g) * 0.02).bfloat16()
w1 = (torch.randn(4 * D, D, device=dev, generator=g) * 0.02).bfloat16()
w2 = (torch.randn(D, 4 * D, device=dev, generator=g) * 0.02).bfloat16()
bq, bo, b1, b2 = (torch.zeros(n, device=dev) for n in (3 * D, D, 4 * D, D))
st = ops.row_stats(x)
st8 = torch.empty(8, M, 2, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gam, bet = torch.ones(D, device=dev), torch.zeros(D, device=dev)


def timeit(name, fn, flops, iters=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / iters
    print(f"{name:44s} {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TF/s", flush=True)


qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
y = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
u = torch.empty(M, 4 * D, device=dev, dtype=torch.bfloat16)
timeit("qkv plain", lambda: ops.gemm(x, wq, bias=bq, out=qkv), 2.0 * M * D * 3 * D)
timeit("qkv ln-fold (1 slot)", lambda: ops.gemm_ln(x, wq, bias=bq, row_stats=st, out=qkv), 2.0 * M * D * 3 * D)
timeit("qkv ln-fold (8 slots)", lambda: ops.gemm_ln(x, wq, bias=bq, row_stats=st8, out=qkv), 2.0 * M * D * 3 * D)
timeit("qkv FUSE kernel, no fold", lambda: ops.gemm_ln(x, wq, bias=bq, out=qkv), 2.0 * M * D * 3 * D)
timeit("out+res plain", lambda: ops.gemm(x, wo, bias=bo, residual=y, out=y), 2.0 * M * D * D)
timeit("out+res stats", lambda: ops.gemm_ln(x, wo, bias=bo, residual=y, out=y, stats_out=st8), 2.0 * M * D * D)
timeit("fc1 gelu plain", lambda: ops.gemm(x, w1, bias=b1, act="gelu", out=u), 2.0 * M * D * 4 * D)
timeit("fc1 gelu ln-fold", lambda: ops.gemm_ln(x, w1, bias=b1, row_stats=st8, act="gelu", out=u), 2.0 * M * D * 4 * D)
timeit("fc2+res plain", lambda: ops.gemm(f, w2, bias=b2, residual=y, out=y), 2.0 * M * D * 4 * D)
timeit("fc2+res stats", lambda: ops.gemm_ln(f, w2, bias=b2, residual=y, out=y, stats_out=st8), 2.0 * M * D * 4 * D)
timeit("layernorm", lambda: ops.layernorm(x, gam, bet, 1e-6, out=y), 0.0)
timeit("row_stats", lambda: ops.row_stats(x), 0.0)

"""Backward GEMM rates (dgrad = gemm_nn, wgrad = gemm_tn with fp32 / bf16 outputs) at training shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openvision_b200 import ops
cfgs = {"B16": (512 * 577, 768, 3072), "L14": (256 * 257, 1024, 4096), "H14": (256 * 257, 1280, 5120)}
name = sys.argv[1] if len(sys.argv) > 1 else "B16"
M, D, F = cfgs[name]
dev = "cuda"
def t(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
x = torch.randn(M, D, device=dev).bfloat16()
f = torch.randn(M, F, device=dev).bfloat16()
for (nm, N, K, a_in, b_in) in (("qkv", 3 * D, D, None, x), ("out", D, D, None, x), ("fc1", F, D, None, x), ("fc2", D, F, None, f)):
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    dy = torch.randn(M, N, device=dev).bfloat16()
    fl = 2.0 * M * N * K
    ms = t(lambda: ops.gemm(b_in, w)); print(f"{name} {nm:4s} fwd   NT  {ms:7.3f} ms {fl/ms/1e9:7.1f} TF/s")
    ms = t(lambda: ops.gemm_nn(dy, w)); print(f"{name} {nm:4s} dgrad NN  {ms:7.3f} ms {fl/ms/1e9:7.1f} TF/s")
    ms = t(lambda: ops.gemm_tn(dy, b_in, out_dtype=torch.float32)); print(f"{name} {nm:4s} wgrad TN f32 {ms:7.3f} ms {fl/ms/1e9:7.1f} TF/s")
    ms = t(lambda: ops.gemm_tn(dy, b_in, out_dtype=torch.bfloat16)); print(f"{name} {nm:4s} wgrad TN bf16 {ms:7.3f} ms {fl/ms/1e9:7.1f} TF/s")
    ms = t(lambda: torch.matmul(dy.t(), b_in)); print(f"{name} {nm:4s} wgrad cuBLAS bf16 {ms:7.3f} ms {fl/ms/1e9:7.1f} TF/s")

"""Power / clock per kernel type: loops one kernel for ~2.5 s, samples NVML power and SM clock meanwhile, and reports
average power, clock, time per launch and hence ENERGY per launch.  The B200s of this pool run power-capped, so the
step time of a long forward is set by its energy, not by the sum of burst kernel times."""
import os, sys, threading, time
import torch
import pynvml
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openvision_b200 import ops

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
print("power limit W:", pynvml.nvmlDeviceGetPowerManagementLimit(h) / 1000, flush=True)

B = int(os.environ.get("OVK_PERF_BATCH", "1024"))
L, D, H = 257, 1024, 16
M = B * L
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(M, D, device=dev, generator=g).bfloat16()
f = torch.randn(M, 4 * D, device=dev, generator=g).bfloat16()
qkv = torch.randn(M, 3 * D, device=dev, generator=g).bfloat16()
wq = (torch.randn(3 * D, D, device=dev, generator=g) * 0.02).bfloat16()
wo = (torch.randn(D, D, device=dev, generator=g) * 0.02).bfloat16()
w1 = (torch.randn(4 * D, D, device=dev, generator=g) * 0.02).bfloat16()
w2 = (torch.randn(D, 4 * D, device=dev, generator=g) * 0.02).bfloat16()
bq, bo, b1 = (torch.zeros(n, device=dev) for n in (3 * D, D, 4 * D))
gam, bet = torch.ones(D, device=dev), torch.zeros(D, device=dev)
st8 = torch.zeros(8, M, 2, device=dev)
st8[:, :, 1] = 128.0
oq = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
ou = torch.empty(M, 4 * D, device=dev, dtype=torch.bfloat16)
oy = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
oa = torch.empty(M, D, device=dev, dtype=torch.bfloat16)


def run(name, fn, flops=0.0, secs=2.5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    samples, stop = [], [False]

    def sampler():
        while not stop[0]:
            samples.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            time.sleep(0.05)
    th = threading.Thread(target=sampler)
    th.start()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop[0] = True
    th.join()
    ms = e0.elapsed_time(e1) / n
    half = samples[len(samples) // 2:]          # second half: the governor has settled
    pw = sum(p for p, _ in half) / len(half)
    ck = sum(c for _, c in half) / len(half)
    tf = f" {flops / ms / 1e9:7.1f} TF/s" if flops else ""
    print(f"{name:34s} {ms:7.3f} ms {pw:6.0f} W {ck:5.0f} MHz  {pw * ms / 1e3:7.3f} J/launch{tf}", flush=True)
    time.sleep(1.0)


run("idle-ish (row_stats)", lambda: ops.row_stats(x))
run("layernorm", lambda: ops.layernorm(x, gam, bet, 1e-6, out=oy))
run("attention fwd", lambda: ops.attention(qkv, B, L, H, 64, out=oa) if False else ops.attention(qkv, B, L, H, 64), 4.0 * B * H * L * L * 64)
run("gemm qkv plain", lambda: ops.gemm(x, wq, bias=bq, out=oq), 2.0 * M * D * 3 * D)
run("gemm qkv ln-fold", lambda: ops.gemm_ln(x, wq, bias=bq, row_stats=st8, out=oq), 2.0 * M * D * 3 * D)
run("gemm out+res plain", lambda: ops.gemm(x, wo, bias=bo, residual=oy, out=oy), 2.0 * M * D * D)
run("gemm out+res stats", lambda: ops.gemm_ln(x, wo, bias=bo, residual=oy, out=oy, stats_out=st8), 2.0 * M * D * D)
run("gemm fc1 gelu plain", lambda: ops.gemm(x, w1, bias=b1, act="gelu", out=ou), 2.0 * M * D * 4 * D)
run("gemm fc1 gelu ln-fold", lambda: ops.gemm_ln(x, w1, bias=b1, row_stats=st8, act="gelu", out=ou), 2.0 * M * D * 4 * D)
run("gemm fc1 linear plain", lambda: ops.gemm(x, w1, bias=b1, out=ou), 2.0 * M * D * 4 * D)
run("gemm fc2+res plain", lambda: ops.gemm(f, w2, bias=bo, residual=oy, out=oy), 2.0 * M * D * 4 * D)
run("cuBLAS fc1 (torch.matmul)", lambda: torch.matmul(x, w1.t(), out=ou), 2.0 * M * D * 4 * D)
run("cuBLAS fc2 (torch.matmul)", lambda: torch.matmul(f, w2.t(), out=oy), 2.0 * M * D * 4 * D)

#!/usr/bin/env python
"""Headline benchmark of the hot path (BASELINE.json): ViT-L/14@224 image-encoder forward, bf16, batch 1024 synthetic
images per GPU, random-init weights -> images/s; plus the CLIP contrastive-loss fwd+bwd time at global batch 32 768.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one pass of the image tower over one batch (the batch shards over ranks with no data-path collective:
weak scaling).  `value` is timed with the batch already in HBM; `e2e` goes through the public
module call (CLIP.encode_image) from PINNED HOST memory with the H2D copy of the images and the D2H copy of the
embeddings inside the timed region.  Prints ONE JSON line on rank 0.

--impl reference times the reference's CPU implementation of the same path: the oracle port (oracle/vit_oracle.py,
a plain-PyTorch fp32 restatement of the reference ops; the reference itself is Python and cannot travel to the GPU
box) with all host threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "vit_l14_224_image_encoder_throughput"
UNIT = "images/s"
WORKLOAD = "OpenVision ViT-L/14@224 image-encoder forward (BASELINE.json configs[1])"
CFG_NAME = "L14-224"
FWD_FLOPS_PER_IMAGE = 162.03e9   # 2*M*N*K of every GEMM-shaped op, L/14@224


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))["mean_dram_bytes_per_launch"]
        except Exception:
            return None
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ----------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7:
                self.samples.append((time.time(), parts))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [p for (t, p) in self.samples if t0 <= t <= t1] or [p for (_, p) in self.samples]
        sm, mx, reasons = [], None, set()
        for p in rows:
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on host cores (reported baseline, and the whole of --impl reference)
# ----------------------------------------------------------------------------------------------------------------
def cpu_port_throughput(sample_batch: int, iters: int, warmup: int):
    import torch
    from oracle import synth, vit_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.make_state_dict(CFG_NAME, 0, vision_only=True)
    images = synth.make_images(CFG_NAME, sample_batch, 0)
    heads = synth.vision_heads(CFG_NAME)

    def step():
        with torch.no_grad():
            return O.l2_normalize(O.vision_transformer(images, sd, heads, pool_type="avg", final_ln_after_pool=True))

    for _ in range(warmup):
        step()
    times = []
    for _ in range(iters):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return sample_batch / statistics.median(times), cores, times


def run_reference(args):
    """Reference arm: CPU implementation of the path (oracle port), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 8
    t0 = time.perf_counter()
    ips, cores, times = cpu_port_throughput(sample, max(1, args.steps), max(1, args.warmup))
    ms = statistics.median(times) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample} images per step (bounded sample of the 1024-image batch)",
                   "weights": "random-init (oracle/synth.py seed 0)"},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample}-image L/14@224 forward, fp32, torch CPU ops, {cores} threads"},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
class KernelRecorder:
    """Brackets individual libovk launches with CUDA events on the launching (current) stream."""

    def __init__(self):
        self.enabled = False
        self.records = {}   # kind -> list of (start_event, end_event, work)

    def __call__(self, kind, work):
        return _Bracket(self, kind, work) if self.enabled else _NULLCTX

    def summary(self):
        out = {}
        for kind, recs in self.records.items():
            ms = [s.elapsed_time(e) for (s, e, _) in recs]
            work = [w for (_, _, w) in recs]
            out[kind] = dict(launches=len(recs), total_ms=sum(ms), total_work=sum(work))
        return out


class _Bracket:
    def __init__(self, rec, kind, work):
        self.rec, self.kind, self.work = rec, kind, work

    def __enter__(self):
        import torch
        self.s = torch.cuda.Event(enable_timing=True)
        self.e = torch.cuda.Event(enable_timing=True)
        self.s.record()

    def __exit__(self, *a):
        self.e.record()
        self.rec.records.setdefault(self.kind, []).append((self.s, self.e, self.work))
        return False


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULLCTX = _NullCtx()


def bench_loss(torch, ovb, n_global, embed, world, rank, steps, warmup, peaks):
    """CLIP contrastive loss fwd+bwd at global batch `n_global` (BASELINE.json configs[3]); None if not in this build."""
    if not hasattr(ovb, "ClipLoss"):
        return None
    import torch.distributed as dist
    n_loc = n_global // world
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    img = torch.nn.functional.normalize(torch.randn(n_loc, embed, device="cuda", generator=g), dim=-1)
    txt = torch.nn.functional.normalize(0.6 * img + 0.8 * torch.nn.functional.normalize(
        torch.randn(n_loc, embed, device="cuda", generator=g), dim=-1), dim=-1)
    img = img.bfloat16().requires_grad_(True)
    txt = txt.bfloat16().requires_grad_(True)
    log_scale = torch.tensor(2.6592, device="cuda", requires_grad=True)
    crit = ovb.ClipLoss(local_loss=world > 1, gather_with_grad=world > 1, rank=rank, world_size=world)

    def step():
        img.grad = txt.grad = log_scale.grad = None
        loss = crit(img, txt, log_scale.exp())
        loss.backward()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    flops = 6.0 * n_global * n_global * embed / world
    tf = flops / ms / 1e9
    return {"metric": "clip_loss_fwd_bwd_ms", "value": ms, "unit": "ms", "global_batch": n_global, "embed_dim": embed,
            "mode": "local_loss+gather_with_grad" if world > 1 else "single", "loss": float(loss.detach()),
            "achieved_tflops_per_gpu": tf, "frac_of_measured_sustained": tf / peaks["tf_sustained"]}


def torch_eager_throughput(torch, cfg, sample_batch: int, iters: int = 3):
    """Comparison point of SURVEY.md §8(d): what the reference's modules ARE on a GPU — stock torch.nn building blocks
    (nn.Conv2d patch embed, nn.LayerNorm, nn.MultiheadAttention, nn.Linear, nn.GELU; transformer.py:210-265,609-651) run
    eagerly in bf16 (cuBLAS / cuDNN / ATen kernels) on the same box.  None of this repo's kernels or modules is involved;
    reported next to `value` as a baseline, never used by the product path."""
    nn, F = torch.nn, torch.nn.functional
    v = cfg["vision"]
    D, P, size, layers = v["width"], v["patch_size"], v["image_size"], v["layers"]
    H, E = D // v["head_width"], cfg["embed_dim"]
    n_tok = (size // P) ** 2

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.ln_1, self.ln_2 = nn.LayerNorm(D, eps=1e-6), nn.LayerNorm(D, eps=1e-6)
            self.attn = nn.MultiheadAttention(D, H, batch_first=True)
            self.c_fc, self.c_proj, self.gelu = nn.Linear(D, 4 * D), nn.Linear(4 * D, D), nn.GELU()

        def forward(self, x):
            h = self.ln_1(x)
            x = x + self.attn(h, h, h, need_weights=False)[0]
            return x + self.c_proj(self.gelu(self.c_fc(self.ln_2(x))))

    class Tower(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(3, D, P, P, bias=False)
            self.cls = nn.Parameter(torch.randn(D) * D ** -0.5)
            self.pos = nn.Parameter(torch.randn(n_tok + 1, D) * D ** -0.5)
            self.blocks = nn.ModuleList([Block() for _ in range(layers)])
            self.ln_post = nn.LayerNorm(D, eps=1e-6)
            self.proj = nn.Parameter(torch.randn(D, E) * D ** -0.5)

        def forward(self, img):
            x = self.conv1(img).flatten(2).transpose(1, 2)
            x = torch.cat([self.cls.expand(x.shape[0], 1, -1).to(x.dtype), x], 1) + self.pos.to(x.dtype)
            for b in self.blocks:
                x = b(x)
            return F.normalize(self.ln_post(x[:, 1:].mean(1)) @ self.proj, dim=-1)

    torch.manual_seed(0)
    tower = Tower().cuda().to(torch.bfloat16).eval()
    img = torch.randn(sample_batch, 3, size, size, device="cuda").to(torch.bfloat16)
    with torch.no_grad():
        for _ in range(2):
            tower(img)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            tower(img)
        e.record()
        torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    del tower, img
    torch.cuda.empty_cache()
    return sample_batch / ms * 1e3


def run_ours(args):
    import torch
    import torch.distributed as dist

    import openvision_b200 as ovb
    from openvision_b200 import _lib, ops
    from oracle import synth   # synthetic weights only (deterministic generator); no oracle compute on this arm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    lib = _lib.load()
    if lib.ovk_device_supported() != 0:
        raise SystemExit("libovk: " + lib.ovk_last_error().decode())
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peaks = load_peaks()

    cfg = synth.CONFIGS[CFG_NAME]
    batch = args.batch
    torch.manual_seed(0)
    model = ovb.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]))
    del model.transformer, model.token_embedding     # text tower is not part of this benchmark
    model = model.cuda().eval()
    ovb.convert_weights_to_lp(model, torch.bfloat16)  # precision='bf16' semantics (factory.py:275-292)
    g = torch.Generator(device="cuda").manual_seed(rank)
    images_dev = torch.randn(batch, 3, 224, 224, device="cuda", generator=g)          # fp32 NCHW, as the callers pass
    images_host = torch.empty((batch, 3, 224, 224), dtype=torch.float32).pin_memory()
    images_host.copy_(images_dev)
    out_host = torch.empty((batch, cfg["embed_dim"]), dtype=torch.float32).pin_memory()

    rec = KernelRecorder()
    ops.recorder = rec

    def step_resident():
        with torch.no_grad():
            return model.encode_image(images_dev, normalize=True)

    # End-to-end: every step copies ITS batch from pinned host memory (H2D, 616 MB) and reads its embeddings back (D2H),
    # all inside the timed region.  The copy of batch i+1 runs on a side stream into the other of two device buffers
    # while batch i is being encoded (ordinary input double-buffering around the public encode_image call).
    copy_stream = torch.cuda.Stream()
    bufs = [torch.empty_like(images_dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def run_e2e(steps):
        main = torch.cuda.current_stream()
        with torch.no_grad():
            with torch.cuda.stream(copy_stream):
                bufs[0].copy_(images_host, non_blocking=True)
                ready[0].record(copy_stream)
            for i in range(steps):
                cur, nxt = i % 2, (i + 1) % 2
                if i + 1 < steps:
                    with torch.cuda.stream(copy_stream):
                        if i >= 1:
                            copy_stream.wait_event(consumed[nxt])
                        bufs[nxt].copy_(images_host, non_blocking=True)
                        ready[nxt].record(copy_stream)
                main.wait_event(ready[cur])
                y = model.encode_image(bufs[cur], normalize=True)
                consumed[cur].record(main)
                out_host.copy_(y.float(), non_blocking=True)
        main.synchronize()
        return out_host

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, record=False, runner=False):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = ops.launch_count
        rec.enabled = record
        w0 = time.time()
        s.record()
        if runner:
            fn(steps)
        else:
            for _ in range(steps):
                fn()
        e.record()
        barrier()
        rec.enabled = False
        w1 = time.time()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, ops.launch_count - n0, w0, w1

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    total_ms, launches, w0, w1 = timed(step_resident, args.steps, record=True)
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = world * batch * args.steps / (total_ms / 1e3)

    run_e2e(max(2, min(args.warmup, 3)))
    e2e_ms, _, _, _ = timed(run_e2e, args.steps, runner=True)
    e2e_value = world * batch * args.steps / (e2e_ms / 1e3)

    ksum = rec.summary()
    if args.profile_range:
        # one forward step + a short loss leg between cudaProfilerStart / Stop: `ncu --profile-from-start off` sees exactly
        # these launches, whatever the warm-up and step counts were (numbers printed by a run under ncu are not bench values)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        loss_line = bench_loss(torch, ovb, args.loss_batch, 768, world, rank, 1, 1, peaks)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    else:
        loss_line = bench_loss(torch, ovb, args.loss_batch, 768, world, rank, max(3, args.steps), max(3, args.warmup), peaks)

    if rank == 0:
        gm = ksum.get("gemm", dict(launches=0, total_ms=0.0, total_work=0.0))
        gemm_tf = gm["total_work"] / gm["total_ms"] / 1e9 if gm["total_ms"] else 0.0
        kernels = {}
        for kind, d in ksum.items():
            per = d["total_work"] / d["total_ms"] / 1e9 if d["total_ms"] else 0.0
            kernels[kind] = {"launches_per_step": d["launches"] / args.steps, "ms_per_step": d["total_ms"] / args.steps,
                             ("achieved_tflops" if kind in ("gemm", "attention") else "achieved_gbs"):
                                 per if kind in ("gemm", "attention") else per * 1e3}
        cpu_val, cpu_cores, cpu_times = (None, None, None)
        eager_val = None
        if not args.no_cpu_baseline:
            cpu_val, cpu_cores, cpu_times = cpu_port_throughput(8, 3, 1)
            del images_dev, bufs[:]
            torch.cuda.empty_cache()
            eager_val = torch_eager_throughput(torch, cfg, 256)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": batch, "global_batch": batch * world, "tokens": 257,
                       "weights": "random-init, precision='bf16' (LayerNorm fp32)", "parallelism": f"dp{world}",
                       "l2": "inputs larger than L2 (616 MB of fp32 images per step; every layer's activations > 126 MB)"},
            "model_tflops": value * FWD_FLOPS_PER_IMAGE / 1e12 / world,
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel (QKV / out-proj / fc1+GELU / fc2 projections)",
                         "achieved": gemm_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": gemm_tf / peaks["tf_sustained"], "traffic": load_traffic(),
                         "traffic_note": "mean dram__bytes_read+write per launch over the four projection shapes, profiles/r01_gemm_ncu_full.csv",
                         "peak_source": f"{peaks['src']} sustained cuBLAS bf16 (kernel timed inside a long step)",
                         "launches_timed": gm["launches"]},
            "kernels": kernels,
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cpu_cores, "kind": "port",
                             "sample": "8-image L/14@224 forward, fp32 oracle port (torch CPU ops), median of 3"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": images_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
                    "api": "CLIP.encode_image(images, normalize=True); every step's fp32 NCHW batch comes from pinned host memory "
                           "(H2D of batch i+1 double-buffered under the encode of batch i), embeddings read back to the host"},
            "gpu_launches": launches, "clocks": clocks,
            "torch_eager_gpu": {"value": eager_val, "unit": UNIT, "sample": "256-image batches, bf16, stock torch.nn modules "
                                "(Conv2d / LayerNorm / MultiheadAttention / Linear / GELU) run eagerly on the same GPU: what the "
                                "reference's PyTorch path is on a B200 (SURVEY.md 8(d) comparison point); baseline only"},
        }
        if loss_line is not None:
            line["clip_loss"] = loss_line
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train(args):
    """Extra workloads (BASELINE.json configs[2] / configs[4]; not the headline line): image tower fwd+bwd (+ contrastive
    loss fwd+bwd against synthetic text embeddings) per step; with N > 1 ranks the loss runs its NCCL exchange and the
    tower gradients are all-reduced (plain NCCL on the optimizer's flat gradient buffers) and the AdamW step of the training
    recipe (openvision_b200.optim.FlatAdamW) runs, all inside the timed region."""
    import torch
    import torch.distributed as dist

    import openvision_b200 as ovb
    from openvision_b200 import ops
    from oracle import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    name = {"b16_384_train": "B16-384", "h14_train": "H14-224", "l14_train": "L14-224"}[args.workload]
    cfg = synth.CONFIGS[name]
    batch = args.batch if args.batch != 1024 or name != "B16-384" else 512
    torch.manual_seed(0)
    tower = ovb.model._build_vision_tower(cfg["embed_dim"], cfg["vision"]).cuda().train()
    tower.set_grad_checkpointing("mlp" if args.checkpoint == "mlp" else args.checkpoint == "block")
    side = cfg["vision"]["image_size"]
    g = torch.Generator(device="cuda").manual_seed(rank)
    images = torch.randn(batch, 3, side, side, device="cuda", generator=g)
    txt = torch.nn.functional.normalize(torch.randn(batch, cfg["embed_dim"], device="cuda", generator=g), dim=-1)
    log_scale = torch.tensor(2.6592, device="cuda", requires_grad=True)
    crit = ovb.ClipLoss(local_loss=world > 1, gather_with_grad=world > 1, rank=rank, world_size=world)
    named = list(tower.named_parameters()) + [("logit_scale", log_scale)]
    # optimizer of the training recipe on flat buffers (build_optax.py:188-278): p.data / p.grad become views of them
    opt = ovb.FlatAdamW(named, lr=1e-4, b1=0.9, b2=0.95, weight_decay=0.2, grad_clip_norm=1.0)
    step_no = [0]

    def step():
        opt.zero_grad()
        feats = tower(images)
        feats = ovb.model._normalize(feats)
        loss = crit(feats, txt, log_scale.exp())
        loss.backward()
        if world > 1:
            for buf in opt.grad_buffers():     # DP gradient sum over NVLink (one NCCL call per flat buffer)
                dist.all_reduce(buf)
        step_no[0] += 1
        opt.step(lr_mult=ovb.cosine_schedule(step_no[0], 10000, 100), grad_scale=1.0 / world)
        return loss

    for _ in range(max(1, args.warmup)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n0 = ops.launch_count
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.reset_peak_memory_stats()
    s.record()
    for _ in range(args.steps):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    if rank == 0:
        fl = {"B16-384": 110.97e9, "H14-224": 334.59e9, "L14-224": 162.03e9}[name]
        per = ms / args.steps
        line = {"metric": "image_tower_plus_clip_loss_train_step", "value": world * batch / per * 1e3, "unit": "images/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"OpenVision ViT-{name} image tower fwd+bwd + CLIP loss fwd+bwd"
                                       f"{' + NCCL gradient all-reduce' if world > 1 else ''} + AdamW step "
                                       "(bf16 first moment, global-norm clipping, cosine schedule)",
                           "batch_per_gpu": batch, "global_batch": batch * world, "activation_checkpointing": args.checkpoint,
                           "text_features": "synthetic unit vectors (text tower not on this path)"},
                "model_tflops_algorithmic": 3 * fl * batch / per / 1e9, "loss": float(loss.detach()),
                "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30,
                "gpu_launches": ops.launch_count - n0}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="images per GPU per step (BASELINE config: 1024)")
    ap.add_argument("--loss-batch", type=int, default=32768, help="global batch of the contrastive-loss leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-range", action="store_true",
                    help="bracket one forward step + a short loss leg with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    ap.add_argument("--workload", default="l14_fwd", choices=["l14_fwd", "b16_384_train", "h14_train", "l14_train"],
                    help="l14_fwd = the headline benchmark (default); *_train = extra fwd+bwd(+loss) workloads")
    ap.add_argument("--checkpoint", nargs="?", const="block", default="none", choices=["none", "block", "mlp"],
                    help="activation checkpointing (train workloads): per block (the reference's switch), or 'mlp' = keep the "
                         "attention-side activations and recompute only the MLP hidden pair in backward")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "l14_fwd":
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

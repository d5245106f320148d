#!/usr/bin/env python
"""Headline benchmark of the hot path (BASELINE.json): ViT-L/14@224 image-encoder forward, bf16, batch 1024 synthetic
images per GPU, random-init weights -> images/s; plus the CLIP contrastive-loss fwd+bwd time at global batch 32 768.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one pass of the image tower over one batch (the batch shards over ranks with no data-path collective:
weak scaling).  `value` is timed with the batch already in HBM; `e2e` goes through the public
module call (CLIP.encode_image) from PINNED HOST memory with the H2D copy of the images and the D2H copy of the
embeddings inside the timed region.  Prints ONE JSON line on rank 0.  Besides the headline keys the line carries
  clip_loss        : BASELINE's second metric (configs[3]) with its own roofline / cpu_baseline, and at N > 1 a parity
                     check of the NCCL path against the single-GPU fused loss on the SAME global feature set
  reference_gpu    : the UNMODIFIED reference modules (baseline/_ref, see oracle/ref_loader.py) on the same GPU in bf16
  small_batch_latency (inside extra_workloads): ms per encode_image call at batch 1 / 8, eager vs one CUDA-graph launch vs the
                     reference modules on the same GPU (the ov-* scripts' regime)
  extra_workloads  : BASELINE configs[2] (B/16@384 fwd+bwd, batch 512) and configs[4] (H/14 training step, 1024 / GPU)

--impl reference times the reference's own CPU implementation of the path: the unmodified reference modules when
baseline/_ref travelled to the box (cpu_baseline.kind = "reference"), else the oracle port, with all host threads, on a
bounded sample of the same workload.
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
import traceback

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "vit_l14_224_image_encoder_throughput"
UNIT = "images/s"
WORKLOAD = "OpenVision ViT-L/14@224 image-encoder forward (BASELINE.json configs[1])"
CFG_NAME = "L14-224"
FWD_FLOPS_PER_IMAGE = {"Ti16-160": 1.20e9, "L14-224": 162.03e9, "B16-384": 110.97e9, "H14-224": 334.59e9}


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this command
    (profiles/gemm_traffic.json names the capture), or None."""
    p = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d["mean_dram_bytes_per_launch"], d.get("source")
        except Exception:
            return None, None
    return None, None


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
# CPU arm: the reference's own modules (baseline/_ref) or, without them, the oracle port — host cores only.
# This is the one place besides tests/ and smoke() that executes anything under oracle/ (checker / baseline, never product).
# ----------------------------------------------------------------------------------------------------------------
def _reference_available() -> bool:
    from oracle import ref_loader
    return ref_loader.available()


def _build_reference_clip(torch, cfg, vision_only=True):
    """The UNMODIFIED reference CLIP (open_clip/model.py:220) with the named config, random init seed 0."""
    from oracle import ref_loader
    _, mdl, _ = ref_loader.load()
    torch.manual_seed(0)
    model = mdl.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]))
    if vision_only:
        del model.transformer, model.token_embedding
    return model, mdl


def cpu_tower_throughput(sample_batch: int, iters: int, warmup: int):
    """-> (images/s, cores, times, kind, sample description): L/14@224 forward of `sample_batch` images in fp32 on the host."""
    import torch
    from openvision_b200.configs import CONFIGS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = CONFIGS[CFG_NAME]
    g = torch.Generator().manual_seed(0)
    images = torch.randn(sample_batch, 3, 224, 224, generator=g)
    if _reference_available():
        model, _ = _build_reference_clip(torch, cfg)
        model = model.eval()
        kind = "reference"
        desc = (f"{sample_batch}-image L/14@224 forward through the unmodified reference CLIP.encode_image(normalize=True) "
                f"(baseline/_ref), fp32, eval, no_grad, {cores} threads")

        def step():
            with torch.no_grad():
                return model.encode_image(images, normalize=True)
    else:
        from oracle import synth, vit_oracle as O
        sd = synth.make_state_dict(CFG_NAME, 0, vision_only=True)
        heads = synth.vision_heads(CFG_NAME)
        kind = "port"
        desc = f"{sample_batch}-image L/14@224 forward, fp32 oracle port (torch CPU ops), {cores} threads"

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
    return sample_batch / statistics.median(times), cores, times, kind, desc


def cpu_loss_ms(n: int, embed: int, iters: int = 3):
    """Reference ClipLoss fwd+bwd on the host at a bounded N (fp32): -> (ms, cores, kind, sample)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    img = torch.nn.functional.normalize(torch.randn(n, embed, generator=g), dim=-1).requires_grad_(True)
    txt = torch.nn.functional.normalize(torch.randn(n, embed, generator=g), dim=-1).requires_grad_(True)
    ls = torch.tensor(2.6592, requires_grad=True)
    if _reference_available():
        from oracle import ref_loader
        crit = ref_loader.load()[2].ClipLoss()
        kind = "reference"
        fn = lambda: crit(img, txt, ls.exp())   # noqa: E731
    else:
        from oracle import vit_oracle as O
        kind = "port"
        fn = lambda: O.clip_loss(img, txt, ls.exp())   # noqa: E731
    times = []
    for i in range(iters + 1):
        img.grad = txt.grad = ls.grad = None
        t0 = time.perf_counter()
        fn().backward()
        if i:
            times.append(time.perf_counter() - t0)
    return statistics.median(times) * 1e3, cores, kind, f"ClipLoss fwd+bwd at N={n}, E={embed}, fp32, {cores} threads (bounded sample of N=32768)"


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 8
    t0 = time.perf_counter()
    ips, cores, times, kind, desc = cpu_tower_throughput(sample, max(1, args.steps), max(1, args.warmup))
    ms = statistics.median(times) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample} images per step (bounded sample of the 1024-image batch)",
                   "weights": "random-init seed 0"},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
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

    def summary(self, reset=False):
        out = {}
        for kind, recs in self.records.items():
            ms = [s.elapsed_time(e) for (s, e, _) in recs]
            work = [w for (_, _, w) in recs]
            out[kind] = dict(launches=len(recs), total_ms=sum(ms), total_work=sum(work))
        if reset:
            self.records = {}
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
TENSOR_KINDS = ("gemm", "attention", "attention_bwd", "clip_loss_fwd", "clip_loss_grad", "clip_loss_bwd")


def kernel_table(ksum, steps, peaks):
    """per-kind launches / ms per step and achieved rate (tensor kernels: TFLOP/s and fraction of the measured sustained
    cuBLAS rate; bandwidth kernels: GB/s and fraction of the measured copy rate)"""
    out = {}
    for kind, d in ksum.items():
        if not d["total_ms"]:
            continue
        per = d["total_work"] / d["total_ms"] / 1e9
        row = {"launches_per_step": d["launches"] / steps, "ms_per_step": d["total_ms"] / steps}
        if kind in TENSOR_KINDS:
            row.update(achieved_tflops=per, frac_of_measured_sustained=per / peaks["tf_sustained"])
        else:
            row.update(achieved_gbs=per * 1e3, frac_of_measured_hbm=per * 1e3 / peaks["hbm"])
        out[kind] = row
    return out


def global_features(torch, n_global, embed, device="cuda"):
    """ONE global feature set (identical on every rank: same seed), unit rows with a correlated diagonal so the loss is not
    just ln N; ranks take row slices of it, which is what lets rank 0 check the multi-rank result against W = 1."""
    g = torch.Generator(device=device).manual_seed(100)
    F = torch.nn.functional
    img = F.normalize(torch.randn(n_global, embed, device=device, generator=g), dim=-1)
    txt = F.normalize(0.6 * img + 0.8 * F.normalize(torch.randn(n_global, embed, device=device, generator=g), dim=-1), dim=-1)
    return img.bfloat16(), txt.bfloat16()


def bench_loss(torch, ovb, ops, rec, n_global, embed, world, rank, steps, warmup, peaks, check=True):
    """CLIP contrastive loss fwd+bwd at global batch `n_global` (BASELINE.json configs[3]): every rank owns n_global / W rows
    (local_loss + gather_with_grad, the mode whose DDP-averaged gradient is the exact global gradient); the
    time is all-gather + fwd + bwd to (dI_local, dT_local, d logit_scale) per step, max over ranks."""
    import torch.distributed as dist
    n_loc = n_global // world
    img_all, txt_all = global_features(torch, n_global, embed)
    img = img_all[rank * n_loc:(rank + 1) * n_loc].clone().requires_grad_(True)
    txt = txt_all[rank * n_loc:(rank + 1) * n_loc].clone().requires_grad_(True)
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
    torch.cuda.synchronize()
    n0 = ops.launch_count
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    launches = ops.launch_count - n0
    ms = s.elapsed_time(e) / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    # per-kernel times of one more, recorded, step (event brackets serialise nothing but add two records per launch)
    rec.summary(reset=True)
    rec.enabled = True
    step()
    torch.cuda.synchronize()
    rec.enabled = False
    kernels = kernel_table(rec.summary(reset=True), 1, peaks)
    flops = 6.0 * n_global * n_global * embed / world
    tf = flops / ms / 1e9
    out = {"metric": "clip_loss_fwd_bwd_ms", "value": ms, "unit": "ms", "higher_is_better": False,
           "scaling": "strong", "global_batch": n_global, "rows_per_gpu": n_loc, "embed_dim": embed, "dtype": "bf16",
           "mode": "local_loss+gather_with_grad" if world > 1 else "single", "loss": float(loss.detach()),
           "gpu_launches_per_step": launches / steps,
           "roofline": {"bound": "tensor", "kernel": "clip_loss_fwd_kernel + clip_loss_grad_kernel + the two gradient GEMMs",
                        "achieved": tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["tf_sustained"],
                        "algorithmic_flops_per_gpu": flops, "executed_flops_per_gpu": flops * 8.0 / 6.0,
                        "traffic": None,
                        "note": "algorithmic = 6 N^2 E / W (SURVEY.md 8d); the backward recomputes the logits tile by tile, so "
                                "8 N^2 E / W are executed: frac <= 0.75 x the tensor-pipe utilisation",
                        "peak_source": f"{peaks['src']} sustained cuBLAS bf16"},
           "kernels": kernels}
    if check:
        out.update(loss_parity(torch, ovb, dist, img_all, txt_all, img, txt, log_scale, loss, world, rank, n_loc))
    return out


def loss_parity(torch, ovb, dist, img_all, txt_all, img, txt, log_scale, loss, world, rank, n_loc):
    """Rank 0 recomputes the loss with the single-GPU fused path on the WHOLE global feature set and compares:
    mean of the per-rank losses == W=1 loss (1e-3 relative); concatenated per-rank dI / dT == W x the W=1 gradients (each
    rank differentiates its own local objective through the gathers; 2e-2 relative L2); mean of the per-rank
    d logit_scale == the W=1 value.  At W = 1 the check is against a blocked fp32 torch evaluation of a 2048-row slice."""
    if world == 1:
        # W = 1: the fused loss against a blocked fp32 torch evaluation of the same objective on the same bf16 features
        # (1024-row blocks; row LSEs directly, column LSEs by a running logaddexp) — a checker, not the product path
        with torch.no_grad():
            s = log_scale.detach().exp()
            n = img_all.shape[0]
            tf32 = txt_all.float()
            col = torch.full((n,), float("-inf"), device="cuda")
            row_term = torch.zeros((), device="cuda", dtype=torch.float64)
            diag = torch.empty(n, device="cuda")
            for r0 in range(0, n, 1024):
                z = s * img_all[r0:r0 + 1024].float() @ tf32.t()
                idx = torch.arange(r0, min(n, r0 + 1024), device="cuda")
                diag[idx] = z[idx - r0, idx]
                row_term += (torch.logsumexp(z, dim=1) - diag[idx]).double().sum()
                col = torch.logaddexp(col, torch.logsumexp(z, dim=0))
            ref = float(0.5 * (row_term + (col - diag).double().sum()) / n)
        rel = abs(float(loss.detach()) - ref) / abs(ref)
        return {"parity_ok": bool(rel <= 1e-3), "parity": {"loss_torch_fp32_blocked": ref, "rel_loss": rel, "bars": "loss 1e-3 relative"}}
    losses = [torch.zeros_like(loss.detach()) for _ in range(world)]
    dist.all_gather(losses, loss.detach())
    gi = torch.empty((world * n_loc, img.shape[1]), dtype=torch.float32, device="cuda")
    gt = torch.empty_like(gi)
    dist.all_gather_into_tensor(gi, img.grad.float().contiguous())
    dist.all_gather_into_tensor(gt, txt.grad.float().contiguous())
    gs = [torch.zeros((), device="cuda") for _ in range(world)]
    dist.all_gather(gs, log_scale.grad.detach().float().reshape(()))
    res = {}
    if rank == 0:
        i1 = img_all.clone().requires_grad_(True)
        t1 = txt_all.clone().requires_grad_(True)
        ls1 = log_scale.detach().clone().requires_grad_(True)
        l1 = ovb.ClipLoss()(i1, t1, ls1.exp())
        l1.backward()
        mean_loss = float(torch.stack(losses).mean())
        rel_loss = abs(mean_loss - float(l1)) / abs(float(l1))
        rel_i = float((gi / world - i1.grad.float()).norm() / i1.grad.float().norm())
        rel_t = float((gt / world - t1.grad.float()).norm() / t1.grad.float().norm())
        mean_gs = float(torch.stack(gs).mean())
        rel_s = abs(mean_gs - float(ls1.grad)) / (abs(float(ls1.grad)) + 1e-12)
        ok = rel_loss <= 1e-3 and rel_i <= 2e-2 and rel_t <= 2e-2 and rel_s <= 2e-2
        res = {"parity_ok": bool(ok),
               "parity": {"loss_mean_over_ranks": mean_loss, "loss_w1": float(l1), "rel_loss": rel_loss, "rel_l2_dI": rel_i,
                          "rel_l2_dT": rel_t, "rel_d_logit_scale": rel_s,
                          "bars": "loss 1e-3 relative; gradients 2e-2 relative L2 (per-rank gradients / W vs W=1)"}}
        del i1, t1
    return res


def reference_on_gpu(torch, cfg, batch: int, iters: int, loss_n: int, embed: int):
    """SURVEY.md §8(d) 'the thing to beat': the UNMODIFIED reference modules (baseline/_ref) moved to the GPU with
    precision='bf16' semantics (convert_weights_to_lp(bf16), LayerNormFp32 via cast_dtype; factory.py:275-297), PyTorch eager,
    same box.  Tower: CLIP.encode_image(normalize=True) at `batch`; loss: the reference ClipLoss fwd+bwd at N = loss_n."""
    out = {}
    from oracle import ref_loader
    _, mdl, lossm = ref_loader.load()
    torch.manual_seed(0)
    model = mdl.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]),
                     cast_dtype=torch.bfloat16)
    del model.transformer, model.token_embedding
    model = model.cuda().eval()
    mdl.convert_weights_to_lp(model, dtype=torch.bfloat16)
    g = torch.Generator(device="cuda").manual_seed(0)
    b = batch
    while b >= 64:
        try:
            img = torch.randn(b, 3, 224, 224, device="cuda", generator=g).to(torch.bfloat16)
            with torch.no_grad():
                for _ in range(2):
                    model.encode_image(img, normalize=True)
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(iters):
                    model.encode_image(img, normalize=True)
                e.record()
                torch.cuda.synchronize()
            ms = s.elapsed_time(e) / iters
            out["tower"] = {"value": b / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "batch": b, "iters": iters,
                            "what": "unmodified reference CLIP.encode_image(normalize=True), precision='bf16' semantics "
                                    "(convert_weights_to_lp, LayerNormFp32), PyTorch eager on the same GPU"}
            break
        except torch.cuda.OutOfMemoryError:
            b //= 2
            torch.cuda.empty_cache()
    del model
    torch.cuda.empty_cache()
    n = loss_n
    while n >= 2048:
        try:
            ia, ta = global_features(torch, n, embed)
            ia = ia.float().requires_grad_(True)
            ta = ta.float().requires_grad_(True)
            ls = torch.tensor(2.6592, device="cuda", requires_grad=True)
            crit = lossm.ClipLoss()

            def step():
                ia.grad = ta.grad = ls.grad = None
                loss = crit(ia, ta, ls.exp())
                loss.backward()
                return loss

            for _ in range(2):
                step()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(5):
                loss = step()
            e.record()
            torch.cuda.synchronize()
            out["clip_loss"] = {"value": s.elapsed_time(e) / 5, "unit": "ms", "global_batch": n, "loss": float(loss.detach()),
                                "what": "unmodified reference ClipLoss fwd+bwd, fp32 features (bf16-rounded values), 1 GPU"}
            break
        except torch.cuda.OutOfMemoryError:
            n //= 2
            torch.cuda.empty_cache()
    torch.cuda.empty_cache()
    return out


def small_batch_latency(torch, ovb, cfg, batches=(1, 8), iters=30):
    """The ov-* scripts' regime (batch 1-8): ms per CLIP.encode_image(normalize=True) call, launch-bound.  Eager libovk path,
    the same kernels replayed from one CUDA graph (openvision_b200.graphed_encode_image), and the unmodified reference
    modules (bf16, PyTorch eager) on the same GPU when baseline/_ref is present."""
    torch.manual_seed(0)
    model = ovb.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]))
    del model.transformer, model.token_embedding
    model = model.cuda().eval()
    ovb.convert_weights_to_lp(model, torch.bfloat16)
    ref = None
    if _reference_available():
        from oracle import ref_loader
        _, mdl, _ = ref_loader.load()
        ref = mdl.CLIP(embed_dim=cfg["embed_dim"], vision_cfg=dict(cfg["vision"]), text_cfg=dict(cfg["text"]),
                       cast_dtype=torch.bfloat16)
        del ref.transformer, ref.token_embedding
        ref = ref.cuda().eval()
        mdl.convert_weights_to_lp(ref, dtype=torch.bfloat16)

    def t(fn):
        with torch.no_grad():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(iters):
                fn()
            e.record()
            torch.cuda.synchronize()
        return s.elapsed_time(e) / iters

    side = cfg["vision"]["image_size"]
    rows = []
    for b in batches:
        img = torch.randn(b, 3, side, side, device="cuda")
        graphed = ovb.graphed_encode_image(model, img, normalize=True)
        row = {"batch": b, "eager_ms": t(lambda: model.encode_image(img, normalize=True)), "cuda_graph_ms": t(lambda: graphed(img))}
        if ref is not None:
            img16 = img.to(torch.bfloat16)
            row["reference_gpu_eager_ms"] = t(lambda: ref.encode_image(img16, normalize=True))
            row["reference_over_graph"] = row["reference_gpu_eager_ms"] / row["cuda_graph_ms"]
        rows.append(row)
        del graphed
    del model, ref
    torch.cuda.empty_cache()
    return rows


def attention_vs_sdpa(torch, ops, peaks, iters=10):
    """Head-to-head of ovk_attention_fwd against F.scaled_dot_product_attention (cuDNN and flash backends) on the tower's
    attention shapes (transformer.py:225,250-252), same GPU, bf16, CUDA events, burst (kernel timed alone)."""
    from torch.nn.attention import SDPBackend, sdpa_kernel
    F = torch.nn.functional
    rows = []
    for (B, H, L, hd) in ((1024, 16, 257, 64), (512, 12, 577, 64), (1024, 16, 257, 80), (1024, 16, 256, 64)):
        qkv = torch.randn(B * L, 3 * H * hd, device="cuda").bfloat16()
        flops = 4.0 * B * H * L * L * hd

        def t(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(iters):
                fn()
            e.record()
            torch.cuda.synchronize()
            return s.elapsed_time(e) / iters

        row = {"shape": [B, H, L, hd]}
        ms = t(lambda: ops.attention(qkv, B, L, H, hd))
        row["ours_ms"] = ms
        row["ours_tflops"] = flops / ms / 1e9
        row["ours_frac_of_measured_burst"] = flops / ms / 1e9 / peaks["tf_burst"]
        q, k, v = (x.reshape(B, L, H, hd).transpose(1, 2) for x in qkv.view(B * L, 3, H * hd).unbind(1))   # strided views, no copy
        for name, backend in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION)):
            try:
                with sdpa_kernel(backend):
                    ms_b = t(lambda: F.scaled_dot_product_attention(q, k, v))
                row[f"sdpa_{name}_ms"] = ms_b
                row[f"sdpa_{name}_tflops"] = flops / ms_b / 1e9
            except Exception as ex:   # backend not available for this shape / build
                row[f"sdpa_{name}_ms"] = None
                row[f"sdpa_{name}_error"] = str(ex).splitlines()[0][:120]
        best = min([x for x in (row.get("sdpa_cudnn_ms"), row.get("sdpa_flash_ms")) if x], default=None)
        row["speedup_vs_best_sdpa"] = (best / ms) if best else None
        # backward (ovk_attention_bwd_fused: delta / remainder-token kernel + tile kernel + dQ conversion) against autograd of SDPA
        out, lse = ops.attention(qkv, B, L, H, hd, save_lse=True)
        dout = torch.randn(B * L, H * hd, device="cuda").bfloat16()
        ms_bwd = t(lambda: ops.attention_bwd(qkv, out, dout, lse, B, L, H, hd))
        row["ours_bwd_ms"] = ms_bwd
        row["ours_bwd_tflops"] = 2.5 * flops / ms_bwd / 1e9
        for name, backend in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION)):
            try:
                qg, kg, vg = (x.detach().contiguous().requires_grad_(True) for x in (q, k, v))
                with sdpa_kernel(backend):
                    o = F.scaled_dot_product_attention(qg, kg, vg)
                    go = torch.randn_like(o)
                    ms_b = t(lambda: torch.autograd.grad(o, (qg, kg, vg), go, retain_graph=True))
                row[f"sdpa_{name}_bwd_ms"] = ms_b
                del o, go, qg, kg, vg
            except Exception as ex:
                row[f"sdpa_{name}_bwd_ms"] = None
                row[f"sdpa_{name}_bwd_error"] = str(ex).splitlines()[0][:120]
        best_b = min([x for x in (row.get("sdpa_cudnn_bwd_ms"), row.get("sdpa_flash_bwd_ms")) if x], default=None)
        row["bwd_speedup_vs_best_sdpa"] = (best_b / ms_bwd) if best_b else None
        rows.append(row)
        del qkv, q, k, v, out, lse, dout
    torch.cuda.empty_cache()
    return rows


def train_workload(torch, ovb, ops, rec, name, batch, world, rank, steps, warmup, checkpoint, peaks, optimizer=True):
    """BASELINE configs[2] / configs[4]: image tower fwd+bwd + contrastive loss fwd+bwd against synthetic text embeddings
    (+ with N > 1 the loss's NCCL exchange and the gradient all-reduce) + the AdamW step of the training recipe
    (openvision_b200.optim.FlatAdamW), all inside the timed region."""
    import torch.distributed as dist
    from openvision_b200.configs import CONFIGS
    cfg = CONFIGS[name]
    torch.manual_seed(0)
    tower = ovb.model._build_vision_tower(cfg["embed_dim"], cfg["vision"]).cuda().train()
    tower.set_grad_checkpointing("mlp" if checkpoint == "mlp" else checkpoint == "block")
    side = cfg["vision"]["image_size"]
    g = torch.Generator(device="cuda").manual_seed(rank)
    images = torch.randn(batch, 3, side, side, device="cuda", generator=g)
    txt = torch.nn.functional.normalize(torch.randn(batch, cfg["embed_dim"], device="cuda", generator=g), dim=-1)
    log_scale = torch.tensor(2.6592, device="cuda", requires_grad=True)
    crit = ovb.ClipLoss(local_loss=world > 1, gather_with_grad=world > 1, rank=rank, world_size=world)
    named = list(tower.named_parameters()) + [("logit_scale", log_scale)]
    opt = ovb.FlatAdamW(named, lr=1e-4, b1=0.9, b2=0.95, weight_decay=0.2, grad_clip_norm=1.0)
    step_no = [0]
    # DP gradient sum over NVLink: buckets of the flat gradient buffers, each all-reduced (NCCL, async) from a
    # post-accumulate-grad hook as soon as its last gradient is in, i.e. under the rest of the backward pass
    reducer = opt.overlap_all_reduce(bucket_bytes=64 << 20) if world > 1 else None

    def step():
        opt.zero_grad()
        feats = tower(images)
        feats = ovb.model._normalize(feats)
        loss = crit(feats, txt, log_scale.exp())
        loss.backward()
        if reducer is not None:
            reducer.finish()
        step_no[0] += 1
        opt.step(lr_mult=ovb.cosine_schedule(step_no[0], 10000, 100), grad_scale=1.0 / world)
        return loss

    for _ in range(max(1, warmup)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = ops.launch_count
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.reset_peak_memory_stats()
    rec.summary(reset=True)
    rec.enabled = True
    s.record()
    for _ in range(steps):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    rec.enabled = False
    ms = s.elapsed_time(e)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    per = ms / steps
    fl = FWD_FLOPS_PER_IMAGE[name]
    tf = 3 * fl * batch / per / 1e9
    kernels = kernel_table(rec.summary(reset=True), steps, peaks)
    out = {"metric": "image_tower_plus_clip_loss_train_step", "value": world * batch / per * 1e3, "unit": "images/s",
           "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": per, "higher_is_better": True,
           "scaling": "weak", "dtype": "bf16", "data": "synthetic",
           "config": {"workload": f"OpenVision ViT-{name} image tower fwd+bwd + CLIP loss fwd+bwd"
                                  f"{' + NCCL gradient all-reduce' if world > 1 else ''} + AdamW step "
                                  "(bf16 first moment, global-norm clipping, cosine schedule)",
                      "batch_per_gpu": batch, "global_batch": batch * world, "activation_checkpointing": checkpoint,
                      "text_features": "synthetic unit vectors (text tower not on this path)"},
           "roofline": {"bound": "tensor", "kernel": "whole step (GEMMs + attention fwd/bwd)", "achieved": tf,
                        "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["tf_sustained"],
                        "algorithmic_flops_per_gpu": 3 * fl * batch, "traffic": None,
                        "note": "algorithmic = 3 x forward GEMM-shaped FLOPs (SURVEY.md 8d), recompute not counted",
                        "peak_source": f"{peaks['src']} sustained cuBLAS bf16"},
           "kernels": kernels, "loss": float(loss.detach()), "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30,
           "gpu_launches": ops.launch_count - n0}
    if reducer is not None:
        reducer.remove()
        out["config"]["gradient_all_reduce"] = f"{len(reducer.buckets)} buckets of <= 64 MiB launched from backward hooks (overlapped)"
    del tower, opt, images
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import openvision_b200 as ovb
    from openvision_b200 import _lib, ops
    from openvision_b200.configs import CONFIGS

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
    rec = KernelRecorder()
    ops.recorder = rec

    if args.workload != "l14_fwd":     # a single extra workload as its own line
        name = {"b16_384_train": "B16-384", "h14_train": "H14-224", "l14_train": "L14-224"}[args.workload]
        batch = args.batch if args.batch != 1024 or name != "B16-384" else 512
        line = train_workload(torch, ovb, ops, rec, name, batch, world, rank, args.steps, args.warmup, args.checkpoint, peaks)
        if rank == 0:
            line["vs_baseline"] = None
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    cfg = CONFIGS[CFG_NAME]
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
    ksum = rec.summary(reset=True)

    run_e2e(max(2, min(args.warmup, 3)))
    e2e_ms, _, _, _ = timed(run_e2e, args.steps, runner=True)
    e2e_value = world * batch * args.steps / (e2e_ms / 1e3)

    if args.profile_range:
        # one forward step + a short loss leg between cudaProfilerStart / Stop: `ncu --profile-from-start off` sees exactly
        # these launches, whatever the warm-up and step counts were (numbers printed by a run under ncu are not bench values)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        loss_line = bench_loss(torch, ovb, ops, rec, args.loss_batch, 768, world, rank, 1, 1, peaks, check=False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    else:
        loss_line = bench_loss(torch, ovb, ops, rec, args.loss_batch, 768, world, rank, max(5, args.steps), max(3, args.warmup), peaks)

    # everything below is context beside the headline: comparison points and the other BASELINE configs.  Each leg is
    # fenced so that a failure there is reported in the line instead of losing it.
    del images_dev, bufs[:]
    torch.cuda.empty_cache()
    extras, errors = {}, {}

    def leg(key, fn):
        try:
            extras[key] = fn()
        except Exception as ex:   # noqa: BLE001
            errors[key] = f"{type(ex).__name__}: {str(ex).splitlines()[0][:200] if str(ex) else ''}"
            traceback.print_exc(file=sys.stderr)
            torch.cuda.empty_cache()

    if not args.no_extras:
        if rank == 0:
            leg("attention_vs_sdpa", lambda: attention_vs_sdpa(torch, ops, peaks))
            leg("small_batch_latency", lambda: small_batch_latency(torch, ovb, cfg))
        del model
        torch.cuda.empty_cache()
        w_steps = max(2, min(args.steps, 4))
        leg("b16_384_fwd_bwd", lambda: train_workload(torch, ovb, ops, rec, "B16-384", 512, world, rank, w_steps, 2, "none", peaks))
        leg("h14_train_step", lambda: train_workload(torch, ovb, ops, rec, "H14-224", 1024, world, rank, max(2, w_steps - 1), 1, "mlp", peaks))
    cpu = dict(value=None, unit=UNIT, cores=None, kind=None, sample=None)
    if rank == 0 and not args.no_cpu_baseline:
        if not args.no_extras and _reference_available():
            leg("reference_gpu", lambda: reference_on_gpu(torch, cfg, batch, 10, args.loss_batch, 768))
        try:
            v, cores, _, kind, desc = cpu_tower_throughput(8, 3, 1)
            cpu = dict(value=v, unit=UNIT, cores=cores, kind=kind, sample=desc + ", median of 3")
            lv, lcores, lkind, ldesc = cpu_loss_ms(4096, 768)
            loss_line["cpu_baseline"] = dict(value=lv, unit="ms", cores=lcores, kind=lkind, sample=ldesc)
        except Exception as ex:   # noqa: BLE001
            errors["cpu_baseline"] = f"{type(ex).__name__}: {ex}"

    if rank == 0:
        gm = ksum.get("gemm", dict(launches=0, total_ms=0.0, total_work=0.0))
        gemm_tf = gm["total_work"] / gm["total_ms"] / 1e9 if gm["total_ms"] else 0.0
        traffic, traffic_src = load_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": batch, "global_batch": batch * world, "tokens": 257,
                       "weights": "random-init, precision='bf16' (LayerNorm fp32)", "parallelism": f"dp{world}",
                       "l2": "inputs larger than L2 (616 MB of fp32 images per step; every layer's activations > 126 MB)"},
            "model_tflops": value * FWD_FLOPS_PER_IMAGE[CFG_NAME] / 1e12 / world,
            "model_frac_of_measured_sustained": value * FWD_FLOPS_PER_IMAGE[CFG_NAME] / 1e12 / world / peaks["tf_sustained"],
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel (QKV / out-proj / fc1+GELU / fc2 projections)",
                         "achieved": gemm_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": gemm_tf / peaks["tf_sustained"], "traffic": traffic,
                         "traffic_note": f"mean dram__bytes_read+write per launch over the projection shapes, from the committed "
                                         f"ncu --set full capture of this command ({traffic_src})",
                         "peak_source": f"{peaks['src']} sustained cuBLAS bf16 (kernel timed inside a long step)",
                         "launches_timed": gm["launches"]},
            "kernels": kernel_table(ksum, args.steps, peaks),
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": images_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
                    "api": "CLIP.encode_image(images, normalize=True); every step's fp32 NCHW batch comes from pinned host memory "
                           "(H2D of batch i+1 double-buffered under the encode of batch i), embeddings read back to the host"},
            "gpu_launches": launches, "clocks": clocks,
            "clip_loss": loss_line,
        }
        if "reference_gpu" in extras:
            rg = extras.pop("reference_gpu")
            if "tower" in rg:
                rg["tower"]["ours_over_reference"] = value / world / rg["tower"]["value"]
            if "clip_loss" in rg and world == 1 and rg["clip_loss"]["global_batch"] == args.loss_batch:
                rg["clip_loss"]["reference_over_ours"] = rg["clip_loss"]["value"] / loss_line["value"]
            line["reference_gpu"] = rg
        att = line["kernels"].get("attention")
        if att and att.get("launches_per_step"):
            # the attention core at this shape sits BELOW the machine's ridge point (4 L hd FLOP per (3 + 1) * hd * 2 bytes of
            # q, k, v in and o out = L / 2 FLOP per byte): its binding roofline is HBM, reported next to the tensor one
            vcfg = cfg["vision"]
            heads, hd = vcfg["width"] // vcfg["head_width"], vcfg["head_width"]
            ltok = (vcfg["image_size"] // vcfg["patch_size"]) ** 2 + 1
            byts = batch * ltok * 4 * heads * hd * 2
            gbs = byts * att["launches_per_step"] / att["ms_per_step"] / 1e6
            att.update(hbm_bytes_per_launch=byts, achieved_gbs=gbs, frac_of_measured_hbm=gbs / peaks["hbm"],
                       flop_per_byte=ltok / 2.0, ridge_flop_per_byte=peaks["tf_sustained"] * 1e3 / peaks["hbm"],
                       bound="hbm" if ltok / 2.0 < peaks["tf_sustained"] * 1e3 / peaks["hbm"] else "tensor")
        if "attention_vs_sdpa" in extras:
            line["kernels"].setdefault("attention", {})["vs_sdpa"] = extras.pop("attention_vs_sdpa")
        if extras:
            line["extra_workloads"] = extras
        if errors:
            line["leg_errors"] = errors
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
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the host-core legs (and the reference-on-GPU comparator)")
    ap.add_argument("--no-extras", action="store_true", help="skip the SDPA head-to-head and the configs[2] / configs[4] legs")
    ap.add_argument("--profile-range", action="store_true",
                    help="bracket one forward step + a short loss leg with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    ap.add_argument("--workload", default="l14_fwd", choices=["l14_fwd", "b16_384_train", "h14_train", "l14_train"],
                    help="l14_fwd = the headline benchmark (default); *_train = one extra workload as its own line")
    ap.add_argument("--checkpoint", nargs="?", const="block", default="none", choices=["none", "block", "mlp"],
                    help="activation checkpointing (train workloads): per block (the reference's switch), or 'mlp' = keep the "
                         "attention-side activations and recompute only the MLP hidden pair in backward")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

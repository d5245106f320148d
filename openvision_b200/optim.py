"""Optimizer of the reference's training recipe on flat buffers (src/optim/build_optax.py:188-278, configs/openvision.py:265-289):

    clip_by_global_norm -> scale_by_adam(b1=0.9, b2=0.95, mu_dtype=bfloat16) -> add_decayed_weights(wd=0.2 on '.*/kernel$')
    -> scale(lr) -> scale_by_schedule(cosine with linear warm-up) -> scale(-1)

The reference trains in JAX; on the PyTorch surface the same transformation is applied to the parameters of the drop-in
modules.  All parameters of one (dtype, decayed?) group live in ONE flat buffer (the module's `p.data` / `p.grad` become
views of it), so a step is one `sumsq` launch + one `adamw` launch per group and the data-parallel gradient all-reduce is
one NCCL call per group on the flat gradient buffer.  No CPU path: the buffers must be CUDA tensors.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import torch

from . import ops
from ._lib import OvkError


def cosine_schedule(step: int, total_steps: int, warmup_steps: int, min_mult: float = 0.0) -> float:
    """Learning-rate multiplier of `create_learning_rate_schedule(decay_type='cosine', warmup_steps=...)`
    (src/helpers/utils.py, used by build_optax.py:199-201): linear warm-up to 1, then cosine decay to `min_mult`."""
    if warmup_steps > 0 and step < warmup_steps:
        return step / warmup_steps
    progress = (step - warmup_steps) / max(1, total_steps - warmup_steps)
    progress = min(max(progress, 0.0), 1.0)
    return min_mult + (1.0 - min_mult) * 0.5 * (1.0 + math.cos(math.pi * progress))


def default_decay_mask(name: str, p: torch.Tensor) -> bool:
    """wd_mults = [('.*/kernel$', 1.0)] (build_optax.py:257): the weight matrices / conv kernels decay; biases, LayerNorm
    parameters, cls / positional embeddings, the temperature AND the vocabulary embedding do not (the reference's text
    embedding is nn.Embed, whose parameter is called 'embedding' — models/text_transformer.py:633 — and matches no
    '/kernel$').  On the PyTorch surface: >= 2-D `*weight` of Linear / Conv / out_proj, the packed `in_proj_weight`, and the
    raw `proj` / `text_projection` matrices; `*embedding.weight` (nn.Embedding) is excluded."""
    parts = name.rsplit(".", 2)
    leaf = parts[-1]
    if leaf == "weight" and len(parts) >= 2 and parts[-2].endswith("embedding"):
        return False
    return p.dim() >= 2 and (leaf in ("weight", "in_proj_weight", "proj", "text_projection"))


def decay_mask_from_modules(model: torch.nn.Module) -> Callable[[str, torch.Tensor], bool]:
    """The same '/kernel$' rule derived from module TYPES instead of names (robust against renamed submodules): weights of
    nn.Linear / nn.ConvNd / nn.MultiheadAttention and raw >= 2-D projection parameters named proj / text_projection decay,
    everything owned by nn.Embedding / nn.LayerNorm does not."""
    decayed = set()
    for mod_name, mod in model.named_modules():
        prefix = mod_name + "." if mod_name else ""
        if isinstance(mod, (torch.nn.Linear, torch.nn.Conv1d, torch.nn.Conv2d)):
            decayed.add(prefix + "weight")
        elif isinstance(mod, torch.nn.MultiheadAttention):
            decayed.update(prefix + n for n in ("in_proj_weight", "q_proj_weight", "k_proj_weight", "v_proj_weight"))
        for n, q in mod.named_parameters(recurse=False):
            if n in ("proj", "text_projection") and q.dim() >= 2:
                decayed.add(prefix + n)
    return lambda name, p: name in decayed


class _Group:
    def __init__(self, params: List[Tuple[str, torch.nn.Parameter]], decay: bool):
        dtype, dev = params[0][1].dtype, params[0][1].device
        sizes = [(-(-p.numel() // 8)) * 8 for _, p in params]   # 16-byte aligned slices for bf16 and fp32 alike
        total = sum(sizes)
        self.decay = decay
        self.names = [n for n, _ in params]
        self.params = [p for _, p in params]
        self.flat_p = torch.zeros(total, dtype=dtype, device=dev)
        self.flat_g = torch.zeros(total, dtype=dtype, device=dev)
        self.mu = torch.zeros(total, dtype=torch.bfloat16, device=dev)
        self.nu = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for (_, p), sz in zip(params, sizes):
            view = self.flat_p[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_g[off:off + p.numel()].view_as(p)
            off += sz


class BucketedGradReducer:
    """Data-parallel gradient SUM overlapped with the backward pass (the reference averages the gradients inside its jitted
    step, main_clip.py:480-483, where XLA overlaps the collective with the remaining backward; on the PyTorch surface that is
    DDP's job).  The flat gradient buffer of every parameter group is cut into buckets of whole parameters; a
    post-accumulate-grad hook per parameter counts its bucket down, and the bucket's NCCL all-reduce is enqueued (async, on
    NCCL's own stream) the moment its last gradient has been accumulated - parameters are laid out in registration order
    and gradients arrive roughly in reverse, so the buckets of the last layers are already on the wire while the first
    layers are still in backward.  finish() launches whatever never fired (parameters without a gradient this step) and
    waits for everything.  Works on any backend (tests run it under gloo on CPU buffers)."""

    class _Bucket:
        __slots__ = ("flat", "lo", "hi", "n", "pending", "work")

    def __init__(self, groups, bucket_bytes: int = 64 << 20, group=None):
        self.process_group = group
        self.buckets = []
        self._handles = []
        for g in groups:
            per = max(1, bucket_bytes // g.flat_g.element_size())
            cur = None
            base = g.flat_g.data_ptr()
            for p in g.params:
                lo = (p.grad.data_ptr() - base) // g.flat_g.element_size()
                hi = lo + p.numel()
                if cur is None or hi - cur.lo > per and cur.n > 0:
                    cur = self._Bucket()
                    cur.flat, cur.lo, cur.hi, cur.n, cur.pending, cur.work = g.flat_g, lo, hi, 0, 0, None
                    self.buckets.append(cur)
                cur.hi = hi
                cur.n += 1
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(cur)))
        self.reset()

    def _make_hook(self, bucket):
        def hook(_param):
            bucket.pending -= 1
            if bucket.pending == 0 and bucket.work is None:
                self._launch(bucket)
        return hook

    def _launch(self, b):
        import torch.distributed as dist
        b.work = dist.all_reduce(b.flat[b.lo:b.hi], group=self.process_group, async_op=True)

    def reset(self) -> None:
        for b in self.buckets:
            b.pending, b.work = b.n, None

    def finish(self) -> None:
        """call after backward(): every bucket reduced and visible to the current stream; re-armed for the next step"""
        for b in self.buckets:
            if b.work is None:
                self._launch(b)
        for b in self.buckets:
            b.work.wait()
        self.reset()

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


class FlatAdamW:
    """scale_by_adam (bf16 first moment) + decoupled weight decay + lr schedule + optional global-norm clipping, on libovk
    kernels (ops.adamw_step / ops.sumsq).  `named_params`: iterable of (name, parameter) with requires_grad."""

    def __init__(self, named_params: Iterable[Tuple[str, torch.nn.Parameter]], lr: float, b1: float = 0.9, b2: float = 0.95,
                 eps: float = 1e-8, weight_decay: float = 0.2, grad_clip_norm: Optional[float] = None,
                 decay_mask: Callable[[str, torch.Tensor], bool] = default_decay_mask):
        named = [(n, p) for n, p in named_params if p.requires_grad]
        if not named:
            raise OvkError("FlatAdamW: no trainable parameters")
        if any(not p.is_cuda for _, p in named):
            raise OvkError("FlatAdamW runs on CUDA parameters only (there is no CPU path)")
        self.lr, self.b1, self.b2, self.eps, self.wd, self.clip = lr, b1, b2, eps, weight_decay, grad_clip_norm
        buckets: Dict[Tuple[torch.dtype, bool], List[Tuple[str, torch.nn.Parameter]]] = {}
        for n, p in named:
            if p.dtype not in (torch.float32, torch.bfloat16):
                raise OvkError(f"FlatAdamW: parameter {n} has dtype {p.dtype}; fp32 and bf16 are supported")
            buckets.setdefault((p.dtype, bool(decay_mask(n, p))), []).append((n, p))
        self.groups = [_Group(ps, decay) for (_, decay), ps in buckets.items()]
        self.step_count = 0
        self._gnorm_sq = torch.zeros(1, dtype=torch.float32, device=named[0][1].device)

    def zero_grad(self) -> None:
        """Gradients accumulate into the flat buffers (autograd adds into an existing .grad): clear them before backward."""
        for g in self.groups:
            g.flat_g.zero_()

    def grad_buffers(self) -> List[torch.Tensor]:
        """The flat gradient buffers, e.g. for `torch.distributed.all_reduce` (sum; pass grad_scale=1/world to step)."""
        return [g.flat_g for g in self.groups]

    def all_reduce_grads(self, bucket_bytes: int = 256 << 20, group=None) -> None:
        """Data-parallel gradient SUM over the ranks (main_clip.py:480-483 does a pmean inside the jitted step): NCCL
        all-reduces over the flat gradient buffers in buckets of `bucket_bytes`, all enqueued asynchronously and waited for
        together, so that NCCL pipelines them back to back (pass grad_scale = 1 / world to step())."""
        import torch.distributed as dist
        works = []
        for g in self.groups:
            per = max(1, bucket_bytes // g.flat_g.element_size())
            for off in range(0, g.flat_g.numel(), per):
                works.append(dist.all_reduce(g.flat_g[off:off + per], group=group, async_op=True))
        for w in works:
            w.wait()

    def overlap_all_reduce(self, bucket_bytes: int = 64 << 20, group=None) -> BucketedGradReducer:
        """Arm the overlapped, bucketed gradient all-reduce (BucketedGradReducer): call once after construction, then
        `reducer.finish()` after every backward() instead of all_reduce_grads()."""
        self.reducer = BucketedGradReducer(self.groups, bucket_bytes, group)
        return self.reducer

    def step(self, lr_mult: float = 1.0, grad_scale: float = 1.0) -> None:
        self.step_count += 1
        gnorm = None
        if self.clip:
            self._gnorm_sq.zero_()
            for g in self.groups:
                ops.sumsq(g.flat_g, self._gnorm_sq)
            gnorm = self._gnorm_sq
        for g in self.groups:
            ops.adamw_step(g.flat_p, g.flat_g, g.mu, g.nu, self.lr * lr_mult, self.b1, self.b2, self.eps,
                           self.wd if g.decay else 0.0, self.step_count, grad_scale, gnorm, self.clip or 0.0)
            # the kernel wrote through raw pointers: tell autograd (and the packed-weight caches keyed on ._version)
            for p in g.params:
                torch.autograd.graph.increment_version(p)

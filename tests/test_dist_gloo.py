"""CPU, world_size 2 and 4 under gloo: the multi-rank host logic of openvision_b200.loss.ClipLoss (feature all-gather,
row offsets, merging of per-rank column statistics, reduce-scatter of the text gradients, gradient weights of the
three reference modes) against the gradients of the UNMODIFIED reference run under gloo (tests/golden/loss_dist_*).
The libovk kernels are replaced by tests/kernel_emulator.py here (no GPU in this container); the same logic runs on
the real kernels in tests/test_gpu_loss.py."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    """A TCP port the OS just handed out on 127.0.0.1 (fixed numbers collide with sockets in TIME_WAIT)."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, local_loss, gwg, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import kernel_emulator
    from openvision_b200 import loss as loss_mod
    from oracle import synth
    loss_mod.ops = kernel_emulator                     # test seam: kernels -> CPU stand-ins
    n, e = 64, 32
    img, txt = synth.make_features(n, e, seed=n + e, dtype=torch.float64)
    nl = n // world
    img_l = img[rank * nl:(rank + 1) * nl].clone().requires_grad_(True)
    txt_l = txt[rank * nl:(rank + 1) * nl].clone().requires_grad_(True)
    ls = torch.tensor(np.log(1 / 0.07), dtype=torch.float64, requires_grad=True)
    crit = loss_mod.ClipLoss(local_loss=local_loss, gather_with_grad=gwg, rank=rank, world_size=world)
    loss = crit(img_l, txt_l, ls.exp())
    loss.backward()
    torch.save(dict(loss=loss.detach(), d_img=img_l.grad, d_txt=txt_l.grad, d_ls=ls.grad), os.path.join(outdir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("local_loss,gwg", [(True, True), (False, False), (False, True)])
def test_multirank_loss_matches_reference(golden, world, local_loss, gwg):
    g = golden(f"loss_dist_W{world}_N64_E32_local{int(local_loss)}_gwg{int(gwg)}.npz")
    port = _free_port()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, local_loss, gwg, port, d), nprocs=world, join=True)
        res = [torch.load(os.path.join(d, f"r{r}.pt")) for r in range(world)]
    # bf16 rounding of the features inside ClipLoss.forward bounds the agreement (features are cast as on the GPU)
    for r, x in enumerate(res):
        assert abs(float(x["loss"]) - float(g[f"loss_r{r}"])) < 2e-2 * abs(float(g[f"loss_r{r}"])) + 1e-3
        for k, kk in (("d_img", "d_img_r"), ("d_txt", "d_txt_r")):
            ref = g[f"{kk}{r}"]
            err = np.abs(x[k].numpy() - ref).max()
            assert err <= 3e-2 * np.abs(ref).max(), (k, r, err)
    # logit_scale: our per-rank value is the row-block share; the mean over ranks is what DDP sees
    ours = np.mean([float(x["d_ls"]) for x in res])
    ref = np.mean([float(g[f"d_logit_scale_r{r}"]) for r in range(world)])
    assert abs(ours - ref) <= 3e-2 * abs(ref) + 1e-4
    if not local_loss:   # global objective on every rank: per-rank values agree with the reference too
        for r, x in enumerate(res):
            ref_r = float(g[f"d_logit_scale_r{r}"])
            assert abs(float(x["d_ls"]) - ref_r) <= 3e-2 * abs(ref_r) + 1e-4


def test_loss_weights_table():
    from openvision_b200.loss import loss_weights
    assert loss_weights(1.0, 16, 64, 1, False, False) == 1 / 32
    assert loss_weights(1.0, 16, 64, 4, True, True) == 1 / 32        # per-rank objective: mean over local rows
    assert loss_weights(1.0, 16, 64, 4, False, False) == 1 / 128     # global objective, own-slice gradient
    assert loss_weights(1.0, 16, 64, 4, False, True) == 1 / 32       # W copies of the global objective
    assert loss_weights(2.0, 16, 64, 4, False, False) == 2 / 128


def _gather_worker(rank, world, local_loss, gwg, port, outdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openvision_b200.loss import gather_features
    from oracle import synth
    img, txt = synth.make_features(16, 8, seed=5, dtype=torch.float64)
    nl = 16 // world
    i_l = img[rank * nl:(rank + 1) * nl].clone().requires_grad_(True)
    t_l = txt[rank * nl:(rank + 1) * nl].clone().requires_grad_(True)
    ai, at = gather_features(i_l, t_l, local_loss, gwg, rank, world)
    # a loss that weights every gathered row differently, so the gradient shows which copies carry autograd
    w = torch.arange(1, 17, dtype=torch.float64)[:, None]
    total = (ai * w).sum() + 2.0 * (at * w).sum()
    if total.requires_grad:      # local_loss without gather_with_grad: the gathered copies are detached (loss.py:52-55)
        total.backward()
    torch.save(dict(ai=ai.detach(), at=at.detach(), gi=i_l.grad, gt=t_l.grad), os.path.join(outdir, f"g{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("local_loss,gwg", [(False, False), (True, False), (False, True)])
def test_gather_features_semantics(local_loss, gwg):
    """loss.py:19-63: every rank sees the concatenation in rank order; with gather_with_grad the gradient of ALL ranks'
    copies flows back (summed over ranks by torch.distributed.nn.all_gather), without it only the rank's own slice carries
    autograd, and only when local_loss is False (loss.py:56-61)."""
    from oracle import synth
    world = 2
    img, txt = synth.make_features(16, 8, seed=5, dtype=torch.float64)
    port = _free_port()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_gather_worker, args=(world, local_loss, gwg, port, d), nprocs=world, join=True)
        res = [torch.load(os.path.join(d, f"g{r}.pt")) for r in range(world)]
    w = torch.arange(1, 17, dtype=torch.float64)[:, None].expand(16, 8)
    for r, x in enumerate(res):
        assert torch.equal(x["ai"], img) and torch.equal(x["at"], txt)
        sl = slice(r * 8, (r + 1) * 8)
        if gwg:          # both ranks' losses differentiate this rank's rows
            assert torch.allclose(x["gi"], world * w[sl]) and torch.allclose(x["gt"], 2.0 * world * w[sl])
        elif not local_loss:
            assert torch.allclose(x["gi"], w[sl]) and torch.allclose(x["gt"], 2.0 * w[sl])
        else:            # no gradient reaches the local features through the gathered copies
            assert x["gi"] is None and x["gt"] is None


def _reducer_worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openvision_b200.optim import BucketedGradReducer, _Group
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.GELU(), torch.nn.Linear(40, 40), torch.nn.GELU(),
                                torch.nn.Linear(40, 8))
    unused = torch.nn.Parameter(torch.zeros(16))          # never receives a gradient: its bucket is launched by finish()
    named = list(model.named_parameters()) + [("unused", unused)]
    group = _Group(named, decay=False)
    red = BucketedGradReducer([group], bucket_bytes=4 * 1000)      # ~1000 floats per bucket -> several buckets
    assert len(red.buckets) >= 3
    launched_early = []
    out = {}
    for step in range(2):                                  # two steps: the reducer re-arms itself
        group.flat_g.zero_()
        g = torch.Generator().manual_seed(100 * step + rank)
        x = torch.randn(16, 24, generator=g)
        model(x).square().sum().backward()
        launched_early.append(sum(b.work is not None for b in red.buckets))
        red.finish()
        out[step] = group.flat_g.clone()
    out["early"] = launched_early
    out["nbuckets"] = len(red.buckets)
    torch.save(out, os.path.join(outdir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_gradient_all_reduce_overlaps_and_sums():
    """SURVEY.md §8f rank 2 (main_clip.py:480-483): the DP gradient sum is bucketed and launched from
    post-accumulate-grad hooks while backward is still running; result = sum of the ranks' gradients, on every rank."""
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_reducer_worker, args=(world, _free_port(), d), nprocs=world, join=True)
        res = [torch.load(os.path.join(d, f"r{r}.pt")) for r in range(world)]
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.GELU(), torch.nn.Linear(40, 40), torch.nn.GELU(),
                                torch.nn.Linear(40, 8))
    for step in range(2):
        total = None
        for rank in range(world):
            model.zero_grad()
            g = torch.Generator().manual_seed(100 * step + rank)
            model(torch.randn(16, 24, generator=g)).square().sum().backward()
            flat = torch.cat([torch.nn.functional.pad(p.grad.reshape(-1), (0, (-p.numel()) % 8)) for p in model.parameters()])
            total = flat if total is None else total + flat
        for rank in range(world):
            got = res[rank][step]
            assert torch.allclose(got[:total.numel()], total, rtol=1e-5, atol=1e-6), (step, rank)
            assert torch.equal(got[total.numel():], torch.zeros_like(got[total.numel():])), "unused parameter's slice stays zero"
    # every bucket except the one holding the gradient-less parameter was already on the wire when backward returned
    for rank in range(world):
        assert all(e >= res[rank]["nbuckets"] - 1 for e in res[rank]["early"]), res[rank]["early"]

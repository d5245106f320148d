"""GPU bring-up probe: runs each libovk kernel against a torch fp32 reference in its own subprocess (a device trap
in one kernel must not poison the others) and prints error statistics plus error *patterns* useful for remote
debugging of descriptor / swizzle mistakes.   Usage: python tools/gpu_probe.py [case ...]
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _stats(name, got, ref, tol):
    import torch
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    bad = (err > tol * denom)
    print(f"[{name}] max_abs_err={err.max().item():.4e} ref_max={denom:.4e} rel={err.max().item()/denom:.3e} "
          f"bad_frac={bad.float().mean().item():.4f} nan={torch.isnan(got).sum().item()}", flush=True)
    return err, bad


def _pattern(bad, rb, cb):
    """print fraction of bad entries per (row block, col block)"""
    import torch
    M, N = bad.shape
    rows = []
    for r0 in range(0, min(M, rb * 8), rb):
        row = []
        for c0 in range(0, min(N, cb * 8), cb):
            row.append(f"{bad[r0:r0+rb, c0:c0+cb].float().mean().item():.2f}")
        rows.append(" ".join(row))
    print("\n".join(rows), flush=True)


def case_gemm(M=512, N=512, K=256, flags="", seed=0):
    import torch
    from openvision_b200 import ops
    torch.manual_seed(seed)
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    bias = torch.randn(N, device="cuda") if "b" in flags else None
    res = torch.randn(M, N, device="cuda").bfloat16() if "r" in flags else None
    act = "gelu" if "g" in flags else None
    out = ops.gemm(a, w, bias=bias, residual=res, act=act)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    if bias is not None:
        ref = ref + bias
    if act:
        ref = torch.nn.functional.gelu(ref)
    if res is not None:
        ref = ref + res.float()
    err, bad = _stats(f"gemm M{M} N{N} K{K} flags='{flags}'", out, ref, 1e-2)
    if bad.any():
        _pattern(bad, 32, 32)
        print("first rows got:", out[0, :8].float().tolist(), "ref:", ref[0, :8].tolist())
    return not bad.any().item()


def case_gemm_basic():
    ok = case_gemm(128, 256, 64)
    ok &= case_gemm(128, 256, 256)
    ok &= case_gemm(512, 512, 256)
    return ok


def case_gemm_tails():
    ok = case_gemm(808, 576, 192, "b")      # Ti/16 QKV: M, N tails
    ok &= case_gemm(808, 192, 768, "br")    # Ti/16 fc2
    ok &= case_gemm(300, 72, 200, "bg")     # BN=128 path, all tails
    ok &= case_gemm(256, 1024, 592, "b")    # patch-embed K tail (592 = 9.25 k-blocks)
    return ok


def case_gemm_epilogues():
    ok = case_gemm(1024, 1024, 1024, "b")
    ok &= case_gemm(1024, 4096, 1024, "bg")
    ok &= case_gemm(1024, 1024, 4096, "br")
    return ok


def case_gemm_big():
    # multi-wave persistent scheduling: more tiles than SMs
    return case_gemm(128 * 150, 1024, 1024, "br", seed=1)


def case_gemm_inplace_residual():
    import torch
    from openvision_b200 import ops
    torch.manual_seed(0)
    M, N, K = 2048, 1024, 1024
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    x = torch.randn(M, N, device="cuda").bfloat16()
    ref = a.float() @ w.float().t() + x.float()
    ops.gemm(a, w, residual=x, out=x)
    torch.cuda.synchronize()
    _, bad = _stats("gemm in-place residual", x, ref, 1e-2)
    return not bad.any().item()


def case_layernorm():
    import torch
    from openvision_b200 import ops
    ok = True
    for rows, D in [(808, 192), (1000, 768), (4096, 1024), (777, 1280), (64, 1152)]:
        torch.manual_seed(0)
        x = (torch.randn(rows, D, device="cuda") * 2 + 0.5).bfloat16()
        g = torch.randn(D, device="cuda")
        b = torch.randn(D, device="cuda")
        y, mean, rstd = ops.layernorm(x, g, b, 1e-6, save_stats=True)
        ref = torch.nn.functional.layer_norm(x.float(), (D,), g, b, 1e-6)
        _, bad = _stats(f"layernorm {rows}x{D}", y, ref, 1e-2)
        ok &= not bad.any().item()
        mref = x.float().mean(-1)
        ok &= (mean - mref).abs().max().item() < 1e-4
    return ok


def case_layernorm_bwd():
    import torch
    from openvision_b200 import ops
    ok = True
    for rows, D in [(808, 192), (3000, 768), (4096, 1024), (500, 1280)]:
        torch.manual_seed(0)
        x = (torch.randn(rows, D, device="cuda") * 2 + 0.5).bfloat16()
        dy = torch.randn(rows, D, device="cuda").bfloat16()
        g = torch.randn(D, device="cuda")
        b = torch.randn(D, device="cuda")
        y, mean, rstd = ops.layernorm(x, g, b, 1e-6, save_stats=True)
        dg = torch.zeros(D, device="cuda")
        db = torch.zeros(D, device="cuda")
        dx = ops.layernorm_bwd(dy, x, g, mean, rstd, dg, db)
        xf = x.float().requires_grad_(True)
        gf = g.clone().requires_grad_(True)
        bf = b.clone().requires_grad_(True)
        torch.nn.functional.layer_norm(xf, (D,), gf, bf, 1e-6).backward(dy.float())
        _, bad1 = _stats(f"ln_bwd dx {rows}x{D}", dx, xf.grad, 1e-2)
        _, bad2 = _stats(f"ln_bwd dgamma {rows}x{D}", dg, gf.grad, 1e-3)
        _, bad3 = _stats(f"ln_bwd dbeta {rows}x{D}", db, bf.grad, 1e-3)
        ok &= not (bad1.any().item() or bad2.any().item() or bad3.any().item())
    return ok


def case_patch_embed():
    import torch
    from openvision_b200 import ops
    ok = True
    for B, H, P, D in [(4, 160, 16, 192), (3, 224, 14, 1024)]:
        torch.manual_seed(0)
        img = torch.randn(B, 3, H, H, device="cuda")
        wconv = torch.randn(D, 3, P, P, device="cuda") * 0.05
        K = 3 * P * P
        ldc = (K + 7) // 8 * 8
        cols = ops.im2col_patches(img, P, ldc)
        ref_cols = torch.nn.functional.unfold(img, P, stride=P).transpose(1, 2).reshape(-1, K)
        _, bad = _stats(f"im2col B{B} H{H} P{P}", cols[:, :K], ref_cols.bfloat16(), 1e-6)
        ok &= not bad.any().item()
        ok &= (cols[:, K:] == 0).all().item()
        wp = torch.zeros(D, ldc, device="cuda", dtype=torch.bfloat16)
        wp[:, :K] = wconv.reshape(D, K).bfloat16()
        tok = ops.gemm(cols, wp)
        N = (H // P) ** 2
        cls = torch.randn(D, device="cuda")
        pos = torch.randn(N + 1, D, device="cuda")
        x = ops.embed_assemble(tok, cls, pos, B, N)
        conv = torch.nn.functional.conv2d(img.bfloat16().float(), wconv.bfloat16().float(), stride=P)
        ref = conv.reshape(B, D, -1).permute(0, 2, 1)
        ref = torch.cat([cls.expand(B, 1, D), ref], 1) + pos
        _, bad = _stats(f"patch_embed B{B} H{H} P{P} D{D}", x, ref, 1e-2)
        ok &= not bad.any().item()
    return ok


def _attn_ref(qkv, B, L, H, hd):
    import torch
    q, k, v = qkv.float().view(B, L, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / hd ** 0.5
    p = torch.softmax(s, -1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * L, H * hd)
    lse = torch.logsumexp(s, -1)
    return o, lse


def case_attention(shapes=None):
    import torch
    from openvision_b200 import ops
    ok = True
    for B, L, H in shapes or [(2, 128, 2), (2, 64, 1), (3, 101, 3), (2, 257, 4), (2, 577, 3), (1, 16, 1), (1, 200, 2)]:
        torch.manual_seed(0)
        hd = 64
        qkv = torch.randn(B * L, 3 * H * hd, device="cuda").bfloat16()
        out, lse = ops.attention(qkv, B, L, H, hd, save_lse=True)
        torch.cuda.synchronize()
        ref, lse_ref = _attn_ref(qkv, B, L, H, hd)
        _, bad = _stats(f"attention B{B} L{L} H{H}", out, ref, 2e-2)
        _, bad2 = _stats(f"attention lse B{B} L{L} H{H}", lse, lse_ref, 1e-3)
        if bad.any():
            _pattern(bad, 32, 16)
        ok &= not (bad.any().item() or bad2.any().item())
    return ok


def case_attention_small():
    return case_attention([(1, 128, 1)])


def case_pool_norm():
    import torch
    from openvision_b200 import ops
    torch.manual_seed(0)
    ok = True
    x = torch.randn(5, 257, 1024, device="cuda").bfloat16()
    p = ops.pool_tokens(x, "avg")
    _, bad = _stats("pool avg", p, x.float()[:, 1:].mean(1), 1e-2)
    ok &= not bad.any().item()
    p = ops.pool_tokens(x, "tok")
    ok &= torch.equal(p, x[:, 0])
    e = torch.randn(100, 768, device="cuda").bfloat16() * 3
    y, n = ops.l2_normalize(e, return_norms=True)
    _, bad = _stats("l2norm", y, torch.nn.functional.normalize(e.float(), dim=-1), 1e-5)
    ok &= not bad.any().item()
    yb = ops.l2_normalize(e, out_dtype=torch.bfloat16)
    _, bad = _stats("l2norm bf16", yb, torch.nn.functional.normalize(e.float(), dim=-1), 1e-2)
    ok &= not bad.any().item()
    return ok



def _e2e(cfg_name, batch, check_text=True):
    import torch
    import openvision_b200 as ov
    from oracle import synth, vit_oracle as O
    cfg = synth.CONFIGS[cfg_name]
    sd = synth.make_state_dict(cfg_name)
    imgs = synth.make_images(cfg_name, batch)
    heads = synth.vision_heads(cfg_name)
    v = cfg["vision"]
    ref = O.vision_transformer(imgs, sd, heads, pool_type=v["pool_type"], final_ln_after_pool=v["final_ln_after_pool"])
    ref_n = O.l2_normalize(ref)
    m = ov.CLIP(cfg["embed_dim"], cfg["vision"], cfg["text"]).cuda().eval()
    m.load_state_dict(sd)
    ok = True
    with torch.no_grad():
        got = m.encode_image(imgs.cuda()).float().cpu()
        got_n = m.encode_image(imgs.cuda(), normalize=True).float().cpu()
    _, bad = _stats(f"e2e {cfg_name} raw", got, ref, 3e-2)
    err = (got_n - ref_n).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(got_n, ref_n, dim=-1).min().item()
    print(f"[e2e {cfg_name}] normalized max_abs={err:.3e} min_cos={cos:.6f}", flush=True)
    ok &= err <= 2e-2 and cos >= 0.9995
    if check_text:
        t = cfg["text"]
        text = synth.make_text(cfg_name, batch)
        if t.get("no_causal_mask", False):
            tref = O.l2_normalize(O.text_transformer(text, sd, t["heads"], causal=False, pool_type=t["pool_type"],
                                                     act="tanh" if (t.get("act_kwargs") or {}).get("approximate") == "tanh" else "erf"))
            with torch.no_grad():
                tgot = m.encode_text(text.cuda(), normalize=True).float().cpu()
            terr = (tgot - tref).abs().max().item()
            tcos = torch.nn.functional.cosine_similarity(tgot, tref, dim=-1).min().item()
            print(f"[e2e {cfg_name}] text normalized max_abs={terr:.3e} min_cos={tcos:.6f}", flush=True)
            ok &= terr <= 2e-2 and tcos >= 0.9995
    return ok


def case_e2e_mini():
    return _e2e("mini-ov", 3) & _e2e("mini-stock", 3)


def case_e2e_ti16():
    return _e2e("Ti16-160", 8)


def case_perf_l14():
    """per-kernel timings at ViT-L/14@224 shapes (batch from OVK_PERF_BATCH, default 256) + whole-tower forward."""
    import torch
    from openvision_b200 import ops
    Bt = int(os.environ.get("OVK_PERF_BATCH", "256"))
    L, D, H, MLP = 257, 1024, 16, 4096
    M = Bt * L
    dev = "cuda"
    torch.manual_seed(0)
    x = torch.randn(M, D, device=dev).bfloat16()
    wqkv = (torch.randn(3 * D, D, device=dev) * 0.03).bfloat16()
    wo = (torch.randn(D, D, device=dev) * 0.03).bfloat16()
    w1 = (torch.randn(MLP, D, device=dev) * 0.03).bfloat16()
    w2 = (torch.randn(D, MLP, device=dev) * 0.03).bfloat16()
    b3 = torch.randn(3 * D, device=dev)
    b1 = torch.randn(MLP, device=dev)
    bD = torch.randn(D, device=dev)
    g = torch.ones(D, device=dev)
    qkv = torch.randn(M, 3 * D, device=dev).bfloat16()
    f = torch.randn(M, MLP, device=dev).bfloat16()

    def timeit(name, fn, flops=None, bytes_=None, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        extra = ""
        if flops:
            extra += f" {flops / ms / 1e9:.1f} TFLOP/s"
        if bytes_:
            extra += f" {bytes_ / ms / 1e6:.1f} GB/s"
        print(f"[perf B={Bt}] {name}: {ms:.3f} ms{extra}", flush=True)
        return ms

    tot = 0.0
    tot += 2 * timeit("layernorm", lambda: ops.layernorm(x, g, bD, 1e-6), bytes_=2 * M * D * 2)
    tot += timeit("gemm qkv  (N=3072,K=1024,bias)", lambda: ops.gemm(x, wqkv, bias=b3), flops=2 * M * 3 * D * D)
    tot += timeit("attention (L=257,H=16)", lambda: ops.attention(qkv, Bt, L, H, 64), flops=4 * Bt * H * L * L * 64)
    tot += timeit("gemm out  (N=1024,K=1024,bias+res)", lambda: ops.gemm(x, wo, bias=bD, residual=x), flops=2 * M * D * D)
    tot += timeit("gemm fc1  (N=4096,K=1024,bias+gelu)", lambda: ops.gemm(x, w1, bias=b1, act="gelu"), flops=2 * M * MLP * D)
    tot += timeit("gemm fc1  (N=4096,K=1024,bias only)", lambda: ops.gemm(x, w1, bias=b1), flops=2 * M * MLP * D) * 0
    tot += timeit("gemm fc2  (N=1024,K=4096,bias+res)", lambda: ops.gemm(f, w2, bias=bD, residual=x), flops=2 * M * MLP * D)
    timeit("torch.matmul qkv (cuBLAS, no bias)", lambda: torch.matmul(x, wqkv.t()), flops=2 * M * 3 * D * D)
    timeit("torch.matmul fc1 (cuBLAS, no bias)", lambda: torch.matmul(x, w1.t()), flops=2 * M * MLP * D)
    timeit("torch.matmul fc2 (cuBLAS, no bias)", lambda: torch.matmul(f, w2.t()), flops=2 * M * MLP * D)
    print(f"[perf B={Bt}] per-layer sum {tot:.3f} ms -> 24 layers {24 * tot:.1f} ms -> {Bt / (24 * tot) * 1e3:.0f} img/s (blocks only)", flush=True)
    # whole tower
    import openvision_b200 as ov
    from oracle import synth
    cfg = synth.CONFIGS["L14-224"]
    vt = ov.model._build_vision_tower(cfg["embed_dim"], cfg["vision"]).cuda().eval()
    imgs = torch.randn(Bt, 3, 224, 224, device=dev)
    with torch.no_grad():
        ms = timeit("ViT-L/14 tower forward", lambda: vt(imgs), flops=Bt * 162.03e9, iters=5)
    print(f"[perf B={Bt}] tower: {Bt / ms * 1e3:.0f} img/s", flush=True)
    return True


CASES = {k[5:]: v for k, v in list(globals().items()) if k.startswith("case_") and callable(v)}


def main():
    names = sys.argv[1:]
    if len(names) == 2 and names[0] == "--one":
        ok = CASES[names[1]]()
        print(f"RESULT {names[1]}: {'PASS' if ok else 'FAIL'}", flush=True)
        sys.exit(0 if ok else 1)
    names = names or ["gemm_basic", "gemm_tails", "gemm_epilogues", "gemm_big", "gemm_inplace_residual", "layernorm",
                      "layernorm_bwd", "patch_embed", "attention_small", "attention", "pool_norm"]
    summary = {}
    for n in names:
        print(f"=== {n}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], timeout=240)
            summary[n] = "PASS" if r.returncode == 0 else f"FAIL(rc={r.returncode})"
        except subprocess.TimeoutExpired:
            summary[n] = "TIMEOUT"
    print("SUMMARY", summary, flush=True)


if __name__ == "__main__":
    main()

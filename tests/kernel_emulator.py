"""TEST-ONLY stand-ins for the libovk loss kernels (same signatures as openvision_b200.ops), written with plain torch
ops so that the multi-rank HOST logic of openvision_b200.loss (gathers, row offsets, column-statistic merging,
reduce-scatter, gradient weights) can run under gloo on CPU.  Never imported by the product package."""
import torch


def clip_loss_fwd(a_loc, b_all, row_offset, scale):
    z = scale * a_loc.double() @ b_all.double().t()
    n = a_loc.shape[0]
    row_lse = torch.logsumexp(z, dim=1)
    diag = z[torch.arange(n), torch.arange(n) + row_offset]
    col_max = z.max(dim=0).values
    col_sum = torch.exp(z - col_max).sum(dim=0)
    return row_lse.float(), diag.float(), col_max.float(), col_sum.float()


def clip_loss_combine(col_max_parts, col_sum_parts):
    m = col_max_parts.double()
    M = m.max(dim=0).values
    return (M + torch.log((col_sum_parts.double() * torch.exp(m - M)).sum(dim=0))).float()


def clip_loss_value(row_lse, col_lse, diag, row_offset):
    n = row_lse.numel()
    a = (row_lse - diag).double().sum()
    b = (col_lse[row_offset:row_offset + n] - diag).double().sum()
    return torch.stack([0.5 * (a + b) / n, a, b]).float()


def clip_loss_grad_logits(a_loc, b_all, row_offset, scale, row_lse, col_lse, w_row, w_col, d_scale, grad_out=None):
    z = scale * a_loc.double() @ b_all.double().t()
    n = a_loc.shape[0]
    if grad_out is not None:
        w_row, w_col = w_row * float(grad_out), w_col * float(grad_out)
    G = w_row * torch.exp(z - row_lse.double()[:, None]) + w_col * torch.exp(z - col_lse.double()[None, :])
    G[torch.arange(n), torch.arange(n) + row_offset] -= (w_row + w_col)
    d_scale += float((G * z).sum() / scale)
    return G


def gemm_nn(a, b, alpha=1.0, out_dtype=torch.float32, **kw):
    return (alpha * a.double() @ b.double()).to(out_dtype)


def gemm_tn(a, b, alpha=1.0, out_dtype=torch.float32, **kw):
    return (alpha * a.double().t() @ b.double()).to(out_dtype)


def gemm_scaled(a, b, a_mn=False, b_mn=False, alpha=1.0, alpha_dev=None, out_dtype=torch.float32, out=None):
    A = a.double().t() if a_mn else a.double()
    B = b.double() if b_mn else b.double().t()
    s = alpha * (float(alpha_dev) if alpha_dev is not None else 1.0)
    return (s * A @ B).to(out_dtype)

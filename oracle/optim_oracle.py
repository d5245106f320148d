"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's optimizer transformation for one flat parameter group.

Follows src/optim/build_optax.py:188-278 (the optax.chain built there) with the hyper-parameters of
src/configs/openvision.py:265-289.  The arithmetic lives in a third-party dependency that is ABSENT from /root/reference and
from this container: optax (requirements.txt of the reference pins the JAX stack; optax's `scale_by_adam`,
`add_decayed_weights`, `clip_by_global_norm` are restated here from their published definitions).  Parity w.r.t. optax itself is
therefore UNPINNED; the restatement is cross-checked against torch.optim.AdamW (same mathematics with an fp32 first moment) in
tests/test_optim.py.  Only tests/ may import this module.
"""
import numpy as np


def _bf16_round(x: np.ndarray) -> np.ndarray:
    """round-to-nearest-even to bfloat16, returned as float32 (optax casts mu to mu_dtype after the update)"""
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def clip_scale(grads, max_norm, gscale=1.0):
    """optax.clip_by_global_norm (build_optax.py:206-210): updates * max_norm / g_norm when g_norm >= max_norm."""
    norm = float(np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads))) * gscale
    return gscale * (max_norm / norm if norm > max_norm else 1.0)


def adamw_step(p, g, mu, nu, lr, b1, b2, eps, wd, step, gscale=1.0, mu_bf16=True):
    """One step on float32 numpy arrays; returns (p, mu, nu).
    scale_by_adam (build_optax.py:213-214 with config.optax = {mu_dtype: bfloat16, b1: 0.9, b2: 0.95}):
        mu = b1 mu + (1-b1) g ; nu = b2 nu + (1-b2) g^2 ; u = (mu / (1-b1^t)) / (sqrt(nu / (1-b2^t)) + eps)
    add_decayed_weights (:252-262): u += wd * p ; scale(lr), schedule (:219-226, 199-204): u *= lr ; scale(-1) (:273): p -= u."""
    g = g.astype(np.float32) * np.float32(gscale)
    mu_new = np.float32(b1) * mu + np.float32(1 - b1) * g
    nu_new = np.float32(b2) * nu + np.float32(1 - b2) * g * g
    mu_hat = mu_new / np.float32(1 - b1 ** step)
    nu_hat = nu_new / np.float32(1 - b2 ** step)
    u = mu_hat / (np.sqrt(nu_hat) + np.float32(eps)) + np.float32(wd) * p
    p_new = p - np.float32(lr) * u
    return p_new.astype(np.float32), (_bf16_round(mu_new) if mu_bf16 else mu_new.astype(np.float32)), nu_new.astype(np.float32)

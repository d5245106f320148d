// Activation functions of the reference surface (shared by the GEMM epilogues and the standalone elementwise kernels).
#pragma once
#include "../../include/ovk.h"
#include "ptx.cuh"

namespace ovk {

// Activations of the reference surface in ONE functional form:  act(x) = x * sigmoid(2 q(x)),
//   q(x) = xc (a0 + a1 xc^2 + a2 xc^4), xc = clamp(x, -8, 8)
//   nn.GELU() exact-erf (transformer.py:232-236, JAX approximate=False vit.py:198-202): Phi(x) = sigmoid(2 atanh(erf(x/sqrt2)))
//       with the odd function atanh(erf(x/sqrt2)) fitted by a0..a2 below: |act - gelu_erf| <= 3.0e-5 for all x
//       (an order of magnitude below bf16 rounding of the output).
//   nn.GELU(approximate='tanh') (text tower act_kwargs): exact, a = (sqrt(2/pi), 0.044715 sqrt(2/pi), 0).
//   QuickGELU x*sigmoid(1.702x) (transformer.py:33-36): exact, a = (0.851, 0, 0).
// sigmoid(2q) = 1 / (1 + 2^(t)), t = -2 log2(e) q  -> one ex2 + one rcp (MUFU) and 6 FMA-pipe ops per element.
struct ActCoef {
  float b0, b1, b2;  // t(x)  = xc (b0 + b1 x2 + b2 x2^2),  b_i = -2 log2(e) a_i
  float d0, d1, d2;  // 2q'(x) = d0 + d1 x2 + d2 x2^2,      d = (2 a0, 6 a1, 10 a2)   (0 outside the clamp)
};

static inline ActCoef act_coef(int act) {
  double a0 = 0, a1 = 0, a2 = 0;
  if (act == OVK_EPI_GELU_ERF) {
    a0 = 0.7974584707815301, a1 = 0.03705034510095251, a2 = -0.0003587323612208004;
  } else if (act == OVK_EPI_GELU_TANH) {
    a0 = 0.7978845608028654, a1 = 0.7978845608028654 * 0.044715;
  } else if (act == OVK_EPI_GELU_QUICK) {
    a0 = 0.851;
  }
  const double c = -2.0 * 1.4426950408889634;
  ActCoef k;
  k.b0 = (float)(c * a0), k.b1 = (float)(c * a1), k.b2 = (float)(c * a2);
  k.d0 = (float)(2 * a0), k.d1 = (float)(6 * a1), k.d2 = (float)(10 * a2);
  return k;
}

#ifndef OVK_ACT_TWO_MUFU
// x * sigmoid(2q) = 0.5 x (1 + tanh(q)): ONE MUFU op (tanh.approx, relative error 2^-11) per element, so that the fc1
// epilogue hides under the MMAs.  Absolute error <= 2.5e-4 * |x| on top of the polynomial fit: below the bf16 rounding
// of the output for x >= -1.5 and below 1e-3 everywhere (define OVK_ACT_TWO_MUFU for the ex2 + rcp form).
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float act_q(float xc, float x2, const ActCoef& k) {   // q(x) = -t(x) / (2 log2 e)
  return xc * fmaf(fmaf(k.d2 * 0.1f, x2, k.d1 * (1.f / 6.f)), x2, k.d0 * 0.5f);
}
__device__ __forceinline__ float act_fwd(float x, const ActCoef& k) {
  const float xc = fminf(fmaxf(x, -8.f), 8.f);
  const float t = fast_tanh(act_q(xc, xc * xc, k));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
// d/dx [0.5 x (1 + tanh q)] = 0.5 (1 + t) + 0.5 x (1 - t^2) q'(x)
__device__ __forceinline__ float act_bwd(float x, const ActCoef& k) {
  const float xc = fminf(fmaxf(x, -8.f), 8.f);
  const float x2 = xc * xc;
  const float t = fast_tanh(act_q(xc, x2, k));
  const float qp = (fabsf(x) < 8.f) ? 0.5f * fmaf(fmaf(k.d2, x2, k.d1), x2, k.d0) : 0.f;   // q'(x)
  return fmaf(0.5f * x * qp, fmaf(-t, t, 1.f), fmaf(0.5f, t, 0.5f));
}
#else
__device__ __forceinline__ float act_fwd(float x, const ActCoef& k) {
  const float xc = fminf(fmaxf(x, -8.f), 8.f);
  const float x2 = xc * xc;
  const float t = xc * fmaf(fmaf(k.b2, x2, k.b1), x2, k.b0);
  return x * fast_rcp(1.f + fast_exp2(t));
}
// d/dx [x sigmoid(2q(x))] = s (1 + x (1 - s) 2q'(x)),  1 - s = e s
__device__ __forceinline__ float act_bwd(float x, const ActCoef& k) {
  const float xc = fminf(fmaxf(x, -8.f), 8.f);
  const float x2 = xc * xc;
  const float t = xc * fmaf(fmaf(k.b2, x2, k.b1), x2, k.b0);
  const float e = fast_exp2(t);
  const float s = fast_rcp(1.f + e);
  const float qp = (fabsf(x) < 8.f) ? fmaf(fmaf(k.d2, x2, k.d1), x2, k.d0) : 0.f;
  return s * fmaf(x * qp, e * s, 1.f);
}


#endif

}  // namespace ovk

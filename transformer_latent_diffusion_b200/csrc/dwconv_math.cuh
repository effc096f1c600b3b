// Channel-pair math of the depthwise-conv + GELU kernels (rowwise.cu, gemm_dwconv.cu): packed fp32 (FFMA2) on the FMA pipe.
#pragma once
#include "ptx.cuh"

namespace tld {

__device__ __forceinline__ float2 unpack_bf16x2(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// erf-GELU on a channel pair, MUFU-free: erf(v/sqrt2) = v R(v^2) with R a degree-7 minimax polynomial on |v| <= 3.96
// (|erf err| <= 4.4e-5; beyond 3.96 the argument is clamped, erf -> 0.99988 instead of 1: |gelu err| <= 3e-4 out there and
// <= 9e-5 inside - the bf16 rounding of the result is 2^-9 relative).  All of it is packed FFMA2 on the FMA pipe: the
// Abramowitz-Stegun form needs a reciprocal and an exponential per element, and MUFU issues only 16 results/clk/SM.
// gelu(v) = v (1/2 + vc R(vc^2)/2) with the halved coefficients folded in: 10 packed FMA-pipe instructions per channel
// pair (the form v/2 + (v/2) erf took 12) - in the fused up-projection kernel this polynomial is over half of the CUDA-core work.
__device__ __forceinline__ float2 gelu2(float2 v) {
  const float2 vc = make_float2(fminf(fmaxf(v.x, -3.96f), 3.96f), fminf(fmaxf(v.y, -3.96f), 3.96f));
  const float2 u = fmul2(vc, vc);
  float2 r = ffma2(make_float2(-1.6720120843416453e-09f, -1.6720120843416453e-09f), u,
                   make_float2(1.2670039950535284e-07f, 1.2670039950535284e-07f));
  r = ffma2(r, u, make_float2(-4.2092156036233065e-06f, -4.2092156036233065e-06f));
  r = ffma2(r, u, make_float2(8.185514889191836e-05f, 8.185514889191836e-05f));
  r = ffma2(r, u, make_float2(-0.001055103144608438f, -0.001055103144608438f));
  r = ffma2(r, u, make_float2(0.009685170836746693f, 0.009685170836746693f));
  r = ffma2(r, u, make_float2(-0.06620126217603683f, -0.06620126217603683f));
  r = ffma2(r, u, make_float2(0.39885684847831726f, 0.39885684847831726f));
  const float2 s = ffma2(vc, r, make_float2(0.5f, 0.5f));           // Phi(v) = 1/2 + erf(v / sqrt 2) / 2
  return fmul2(v, s);
}

// erf-GELU through the logistic form: Phi(v) = 1 / (1 + exp(-2 v p(v^2))) with p(u) = a1 + a3 u + a5 u^2 fitted (minimax on
// |v| <= 8) to atanh(erf(v / sqrt 2)) / v: |gelu error| <= 2.6e-5, tighter than the degree-7 erf polynomial of gelu2 (4.7e-4 at its
// clamp edge).  6 packed FMA-pipe instructions + 2 ex2.approx + 2 rcp.approx (MUFU, 2^-22 relative error) per channel pair
// instead of 10 packed ones: in the fused up-projection kernel the FMA pipe is the bound and the MUFU unit is idle.
// u is clamped at 64 (the fit's range; beyond it exp() has long saturated and a5 < 0 would eventually flip the sign).
__device__ __forceinline__ float2 gelu2_logistic(float2 v) {
  float2 u = fmul2(v, v);
  u.x = fminf(u.x, 64.f);
  u.y = fminf(u.y, 64.f);
  // q(u) = -2 log2(e) p(u):  exp(-2 v p) = 2^(v q)
  float2 q = ffma2(make_float2(1.0142712e-03f, 1.0142712e-03f), u, make_float2(-0.10677578f, -0.10677578f));
  q = ffma2(q, u, make_float2(-2.3011214f, -2.3011214f));
  const float2 a = fmul2(v, q);
  const float2 d = fadd2(make_float2(ex2_approx(a.x), ex2_approx(a.y)), make_float2(1.f, 1.f));
  return fmul2(v, make_float2(rcp_approx(d.x), rcp_approx(d.y)));
}

}  // namespace tld

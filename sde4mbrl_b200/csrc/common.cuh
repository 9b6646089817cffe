// common.cuh -- device math, threefry2x32 and staging helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sde4b200.h"

#define SDE4_DEV __device__ __forceinline__
#define SDE4_HD __host__ __device__ __forceinline__

namespace sde4 {

// ----------------------------------------------------------------------------
// compile-time helpers
// ----------------------------------------------------------------------------
SDE4_HD constexpr int popc(uint32_t m) {
  int c = 0;
  for (int i = 0; i < 32; ++i) c += (m >> i) & 1u;
  return c;
}
// index of the k-th set bit of m (k < popc(m))
SDE4_HD constexpr int nth_bit(uint32_t m, int k) {
  int c = 0;
  for (int i = 0; i < 32; ++i) {
    if ((m >> i) & 1u) {
      if (c == k) return i;
      ++c;
    }
  }
  return -1;
}
SDE4_HD constexpr int up4(int n) { return (n + 3) & ~3; }

// ----------------------------------------------------------------------------
// activations.  value h = act(a), derivative d = act'(a)
// ----------------------------------------------------------------------------
// MUFU-based primitives (2 ulp): ex2.approx / rcp.approx without the denormal / range fix-up code
// that exp2f() and 1.0f/x expand to.
SDE4_DEV float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
SDE4_DEV float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

SDE4_DEV float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
SDE4_DEV float sigmoidf_(float a) { return fast_rcp(1.0f + fast_ex2(a * -1.44269504088896341f)); }

// tanh(a) = 1 - 2/(exp(2a) + 1): 2 MUFU + 3 FP32 ops, absolute error ~1.5e-7 (saturates correctly
// to +-1; for |a| << 1 the error is absolute, not relative -- fine for float32 parity).
SDE4_DEV float tanhf_(float a) {
  const float t = fast_ex2(a * 2.88539008177792681f);  // exp(2a)
  return fmaf(-2.0f, fast_rcp(t + 1.0f), 1.0f);
}

template <int ACT>
SDE4_DEV float act_fwd(float a) {
  if (ACT == SDE4_ACT_TANH) return tanhf_(a);
  if (ACT == SDE4_ACT_SWISH) return a * sigmoidf_(a);
  return a;
}
// value and derivative together (shares the transcendental)
template <int ACT>
SDE4_DEV void act_fwd_d(float a, float& h, float& d) {
  if (ACT == SDE4_ACT_TANH) {
    h = tanhf_(a);
    d = fmaf(-h, h, 1.0f);
  } else if (ACT == SDE4_ACT_SWISH) {
    float s = sigmoidf_(a);
    h = a * s;
    d = fmaf(h, 1.0f - s, s);  // s + a*s*(1-s)
  } else {
    h = a;
    d = 1.0f;
  }
}

// ----------------------------------------------------------------------------
// threefry2x32 (20 rounds) and the jax.random layouts
// ----------------------------------------------------------------------------
SDE4_HD uint32_t rotl32(uint32_t x, int r) {
#ifdef __CUDA_ARCH__
  return __funnelshift_l(x, x, r);
#else
  return (x << r) | (x >> (32 - r));
#endif
}

SDE4_HD void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t& o0,
                          uint32_t& o1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  uint32_t x0 = c0 + k0, x1 = c1 + k1;
#define SDE4_TF_R(r)   \
  x0 += x1;            \
  x1 = rotl32(x1, r);  \
  x1 ^= x0;
  SDE4_TF_R(13) SDE4_TF_R(15) SDE4_TF_R(26) SDE4_TF_R(6)
  x0 += k1; x1 += k2 + 1u;
  SDE4_TF_R(17) SDE4_TF_R(29) SDE4_TF_R(16) SDE4_TF_R(24)
  x0 += k2; x1 += k0 + 2u;
  SDE4_TF_R(13) SDE4_TF_R(15) SDE4_TF_R(26) SDE4_TF_R(6)
  x0 += k0; x1 += k1 + 3u;
  SDE4_TF_R(17) SDE4_TF_R(29) SDE4_TF_R(16) SDE4_TF_R(24)
  x0 += k1; x1 += k2 + 4u;
  SDE4_TF_R(13) SDE4_TF_R(15) SDE4_TF_R(26) SDE4_TF_R(6)
  x0 += k2; x1 += k0 + 5u;
#undef SDE4_TF_R
  o0 = x0;
  o1 = x1;
}

// element j of random_bits(key, (total,)) -- 32-bit, both jax layouts.  Branch free (ONE threefry instance,
// selects around it) so that it can be scheduled inside a straight-line block:
//   partitionable: counter (hi32(j), lo32(j)), word = o0 ^ o1
//   original:      counts iota(total) (padded with one 0 if odd) hashed by halves: element j < half is word 0 of
//                  the block (j, j + half), element j >= half is word 1 of the block (j - half, j)
SDE4_HD uint32_t jax_bits(uint32_t k0, uint32_t k1, uint64_t j, uint64_t total, int mode) {
  const bool part = mode == SDE4_RNG_PARTITIONABLE;
  const uint64_t half = (total + 1) >> 1;
  const bool low = j < half;
  const uint64_t hi = j + half;
  const uint32_t c0 = part ? (uint32_t)(j >> 32) : (low ? (uint32_t)j : (uint32_t)(j - half));
  const uint32_t c1 = (part || !low) ? (uint32_t)j : (hi < total ? (uint32_t)hi : 0u);
  uint32_t o0, o1;
  threefry2x32(k0, k1, c0, c1, o0, o1);
  return part ? (o0 ^ o1) : (low ? o0 : o1);
}

// child `i` of jax.random.split(key, num)
SDE4_HD void jax_split(uint32_t k0, uint32_t k1, uint32_t i, uint32_t num, int mode, uint32_t& c0,
                       uint32_t& c1) {
  if (mode == SDE4_RNG_PARTITIONABLE) {
    threefry2x32(k0, k1, 0u, i, c0, c1);
    return;
  }
  c0 = jax_bits(k0, k1, 2ull * i, 2ull * num, SDE4_RNG_ORIGINAL);
  c1 = jax_bits(k0, k1, 2ull * i + 1, 2ull * num, SDE4_RNG_ORIGINAL);
}

// XLA float32 erf_inv (Giles) -- see SURVEY appendix F.  Both polynomial branches are evaluated and selected
// (the tail one, |x| > 0.9966, is rare but a data-dependent branch would split the caller's basic block).
SDE4_DEV float erfinv_f32(float x) {
  // XLA evaluates -log1p(-(x*x)) with the product rounded first; keep that (no FMA contraction)
  const float w = -__logf(1.0f - __fmul_rn(x, x));
  const float wc = w - 2.5f;
  float p = 2.81022636e-08f;
  p = fmaf(p, wc, 3.43273939e-07f);
  p = fmaf(p, wc, -3.5233877e-06f);
  p = fmaf(p, wc, -4.39150654e-06f);
  p = fmaf(p, wc, 0.00021858087f);
  p = fmaf(p, wc, -0.00125372503f);
  p = fmaf(p, wc, -0.00417768164f);
  p = fmaf(p, wc, 0.246640727f);
  p = fmaf(p, wc, 1.50140941f);
  const float wt = fast_sqrt(fmaxf(w, 0.0f)) - 3.0f;  // MUFU.SQRT (<= 2 ulp); IEEE sqrtf carries a slow-path branch
  float r = -0.000200214257f;
  r = fmaf(r, wt, 0.000100950558f);
  r = fmaf(r, wt, 0.00134934322f);
  r = fmaf(r, wt, -0.00367342844f);
  r = fmaf(r, wt, 0.00573950773f);
  r = fmaf(r, wt, -0.0076224613f);
  r = fmaf(r, wt, 0.00943887047f);
  r = fmaf(r, wt, 1.00167406f);
  r = fmaf(r, wt, 2.83297682f);
  return (w < 5.0f ? p : r) * x;
}

// jax.random.normal transform of one 32-bit word
SDE4_DEV float bits_to_normal(uint32_t bits) {
  const float lo = -0.99999994f;  // nextafter(-1, 0)
  float u = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  float v = fmaxf(lo, fmaf(u, 2.0f, lo));  // (hi - lo) rounds to exactly 2.0f in float32
  return 1.41421356237309505f * erfinv_f32(v);
}

// Per-sample Brownian key: k_p = split(rng, N)[n] (nsde.py:1026,1581), then
// rng_brownian = split(k_p, 2)[0] (sde_solver.py:276).
// chain 1 (training loss, nsde.py:1187 -> :709 -> sde_solver.py:276): one more level,
// rng_brownian = split(k_p, 3)[0], before the solver's own split.
SDE4_DEV void brownian_key(uint32_t r0, uint32_t r1, uint32_t n, uint32_t N, int mode, uint32_t& b0,
                           uint32_t& b1, int chain = 0) {
  uint32_t p0, p1;
  jax_split(r0, r1, n, N, mode, p0, p1);
  if (chain == 1) {
    uint32_t q0, q1;
    jax_split(p0, p1, 0u, 3u, mode, q0, q1);
    p0 = q0;
    p1 = q1;
  }
  jax_split(p0, p1, 0u, 2u, mode, b0, b1);
}

// ----------------------------------------------------------------------------
// parameter staging: one 1-D bulk async copy (TMA engine, UBLKCP) global -> shared,
// completion on an mbarrier.  nbytes must be a multiple of 16, both pointers 16B aligned.
// ----------------------------------------------------------------------------
SDE4_DEV void stage_params_bulk(float* smem_dst, const float* gsrc, uint32_t nbytes, uint64_t* mbar) {
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nbytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(gsrc), "r"(nbytes), "r"(bar)
        : "memory");
  }
  // all threads wait for phase 0
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar)
      : "memory");
}

}  // namespace sde4

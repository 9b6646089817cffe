/* sde_oracle_impl.h -- per-ISA driver (OpenMP over chunks of VL particles) around
 * sde_oracle_core.h.  Test infrastructure / CPU baseline only. */
#include <omp.h>
#include "sde_oracle_core.h"

/* threefry2x32 + jax.random layouts (scalar; see oracle/threefry.py for the citations) */
static inline uint32_t FN(rotl)(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static void FN(tf)(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t* o0, uint32_t* o1) {
  static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
  for (int g = 0; g < 5; ++g) {
    for (int r = 0; r < 4; ++r) {
      x0 += x1;
      x1 = FN(rotl)(x1, R[g % 2][r]);
      x1 ^= x0;
    }
    x0 += ks[(g + 1) % 3];
    x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
  *o0 = x0;
  *o1 = x1;
}
static uint32_t FN(bits)(uint32_t k0, uint32_t k1, uint64_t j, uint64_t total, int mode) {
  uint32_t o0, o1;
  if (mode == SDE4_RNG_PARTITIONABLE) {
    FN(tf)(k0, k1, (uint32_t)(j >> 32), (uint32_t)j, &o0, &o1);
    return o0 ^ o1;
  }
  uint64_t half = (total + 1) >> 1;
  if (j < half) {
    uint64_t hi = j + half;
    FN(tf)(k0, k1, (uint32_t)j, hi < total ? (uint32_t)hi : 0u, &o0, &o1);
    return o0;
  }
  FN(tf)(k0, k1, (uint32_t)(j - half), (uint32_t)j, &o0, &o1);
  return o1;
}
static void FN(split)(uint32_t k0, uint32_t k1, uint32_t i, uint32_t num, int mode, uint32_t* c0, uint32_t* c1) {
  if (mode == SDE4_RNG_PARTITIONABLE) {
    FN(tf)(k0, k1, 0u, i, c0, c1);
    return;
  }
  *c0 = FN(bits)(k0, k1, 2ull * i, 2ull * num, 0);
  *c1 = FN(bits)(k0, k1, 2ull * i + 1, 2ull * num, 0);
}
static float FN(normal)(uint32_t b) {
  union { uint32_t u; float f; } cv;
  cv.u = (b >> 9) | 0x3F800000u;
  float u = cv.f - 1.0f;
  const float lo = -0.99999994f;
  float v = fmaxf(lo, u * 2.0f + lo);
  float w = -log1pf(-(v * v));
  float p;
  if (w < 5.0f) {
    w -= 2.5f;
    p = 2.81022636e-08f;
    p = 3.43273939e-07f + p * w;
    p = -3.5233877e-06f + p * w;
    p = -4.39150654e-06f + p * w;
    p = 0.00021858087f + p * w;
    p = -0.00125372503f + p * w;
    p = -0.00417768164f + p * w;
    p = 0.246640727f + p * w;
    p = 1.50140941f + p * w;
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = 0.000100950558f + p * w;
    p = 0.00134934322f + p * w;
    p = -0.00367342844f + p * w;
    p = 0.00573950773f + p * w;
    p = -0.0076224613f + p * w;
    p = 0.00943887047f + p * w;
    p = 1.00167406f + p * w;
    p = 2.83297682f + p * w;
  }
  return 1.41421356237309505f * (p * v);
}

/* value (+ gradient) of the mean-over-particles MPC cost for ONE problem.
 * noise: NULL -> in-line threefry (keys as nsde.py:1026 -> sde_solver.py:276), else [N][H][13].
 * xtraj: optional [N][H+1][13].  Returns 0. */
int FN(sde4o_rotor_cost_grad)(const sde4_model_desc* d, const sde4_cost_desc* cd, const float* params,
                              const float* y0, const float* opt, const float* dt, const float* xref, int N, int H,
                              const uint32_t* key, int rng_mode, const float* noise, int noisy, int want_grad,
                              int nthreads, float* cost_out, float* grad_out, float* xtraj) {
  rotor_ctx ctx;
  FN(ctx_init)(&ctx, d, params);
  const int nu = d->n_u;
  const int nchunks = (N + VL - 1) / VL;
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  double* gacc = (double*)calloc((size_t)nthreads * (H * nu + 1), sizeof(double));
#pragma omp parallel num_threads(nthreads)
  {
    const int tid = omp_get_thread_num();
    double* my = gacc + (size_t)tid * (H * nu + 1);
    vf* xs = (vf*)aligned_alloc(64, sizeof(vf) * (size_t)(H + 1) * 13);
    vf* aux = (vf*)aligned_alloc(64, sizeof(vf) * (size_t)H);
    vf* dw = (vf*)aligned_alloc(64, sizeof(vf) * (size_t)H * 13);
#pragma omp for schedule(dynamic, 1)
    for (int ch = 0; ch < nchunks; ++ch) {
      const int n0 = ch * VL;
      const int nvalid = N - n0 < VL ? N - n0 : VL;
      if (noisy) {
        for (int l = 0; l < VL; ++l) {
          const int n = n0 + (l < nvalid ? l : 0);
          if (noise) {
            for (int t = 0; t < H; ++t)
              for (int i = 0; i < 13; ++i) dw[t * 13 + i][l] = noise[((size_t)n * H + t) * 13 + i];
          } else {
            uint32_t p0, p1, b0, b1;
            FN(split)(key[0], key[1], (uint32_t)n, (uint32_t)N, rng_mode, &p0, &p1);
            FN(split)(p0, p1, 0u, 2u, rng_mode, &b0, &b1);
            for (int t = 0; t < H; ++t)
              for (int i = 0; i < 13; ++i)
                dw[t * 13 + i][l] = d->noise_prior[i] != 0.0f
                                        ? FN(normal)(FN(bits)(b0, b1, (uint64_t)t * 13 + i, (uint64_t)H * 13, rng_mode))
                                        : 0.0f;
          }
        }
      }
      FN(chunk_cost_grad)(&ctx, cd, y0, opt, dt, xref, H, dw, nvalid, want_grad, xs, aux, &my[H * nu], my,
                          xtraj ? xtraj + (size_t)n0 * (H + 1) * 13 : NULL, noisy);
    }
    free(xs);
    free(aux);
    free(dw);
  }
  double c = 0.0;
  for (int t = 0; t < nthreads; ++t) c += gacc[(size_t)t * (H * nu + 1) + H * nu];
  c /= N;
  /* control slew term of augmented_cost (sde_mpc_design.py:141-148) */
  if (cd->has_slew) {
    double sl = 0.0;
    for (int t = 0; t + 1 < H; ++t)
      for (int j = 0; j < nu; ++j) {
        double dv = ((double)opt[(t + 1) * nu + j] - opt[t * nu + j]) / dt[t];
        sl += cd->u_slew_coeff[j] * dv * dv;
      }
    c += sl * cd->res_mult;
  }
  *cost_out = (float)c;
  if (want_grad && grad_out) {
    for (int k = 0; k < H * nu; ++k) {
      double g = 0.0;
      for (int t = 0; t < nthreads; ++t) g += gacc[(size_t)t * (H * nu + 1) + k];
      g /= N;
      if (cd->has_slew) {
        const int t = k / nu, j = k % nu;
        double gs = 0.0;
        if (t > 0) gs += 2.0 * cd->u_slew_coeff[j] * ((double)opt[k] - opt[k - nu]) / ((double)dt[t - 1] * dt[t - 1]);
        if (t + 1 < H) gs -= 2.0 * cd->u_slew_coeff[j] * ((double)opt[k + nu] - opt[k]) / ((double)dt[t] * dt[t]);
        g += gs * cd->res_mult;
      }
      grad_out[k] = (float)g;
    }
  }
  free(gacc);
  return 0;
}

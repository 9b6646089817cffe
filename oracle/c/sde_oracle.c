/* sde_oracle.c -- runtime ISA dispatch for the C oracle (test infrastructure / CPU baseline only). */
#include <stdint.h>
#include "../../include/sde4b200.h"

#define DECL(sfx)                                                                                                     \
  int sde4o_rotor_cost_grad_vl##sfx(const sde4_model_desc*, const sde4_cost_desc*, const float*, const float*,       \
                                    const float*, const float*, const float*, int, int, const uint32_t*, int,        \
                                    const float*, int, int, int, float*, float*, float*);
DECL(8)
DECL(16)

int sde4o_simd_width(void) { return __builtin_cpu_supports("avx512f") ? 16 : 8; }

int sde4o_rotor_cost_grad(const sde4_model_desc* d, const sde4_cost_desc* cd, const float* params, const float* y0,
                          const float* opt, const float* dt, const float* xref, int N, int H, const uint32_t* key,
                          int rng_mode, const float* noise, int noisy, int want_grad, int nthreads, float* cost_out,
                          float* grad_out, float* xtraj) {
  if (__builtin_cpu_supports("avx512f"))
    return sde4o_rotor_cost_grad_vl16(d, cd, params, y0, opt, dt, xref, N, H, key, rng_mode, noise, noisy, want_grad,
                                      nthreads, cost_out, grad_out, xtraj);
  return sde4o_rotor_cost_grad_vl8(d, cd, params, y0, opt, dt, xref, N, H, key, rng_mode, noise, noisy, want_grad,
                                   nthreads, cost_out, grad_out, xtraj);
}

// kernels.cuh -- the rollout kernel (K1), the fused rollout + cost + adjoint kernel (K2) and
// the particle reduction (K3).
//
// K1  euler_maruyama under create_sampling_fn.multi_sampling
//       sde4mbrl/sde_solver.py:243-317, sde4mbrl/nsde.py:1024-1028
// K2  compute_cost_function under create_online_cost_sampling_fn.multi_sampling and
//     jax.value_and_grad of it   sde4mbrl/nsde.py:1485-1527,1576-1586, sde4mbrl/apg.py:82,224
// K3  jnp.mean over particles (nsde.py:1586) + control-slew term of augmented_cost
//       sde4mbrlExamples/rotor_uav/sde_mpc_design.py:134-167
//
// Thread mapping: one thread per sample g = p*N + n (problem p, particle n); state, controls and
// cotangents live in registers across the whole horizon; network weights are staged once per CTA
// into shared memory by one bulk async copy; every per-sample array is sample-minor so a warp
// reads / writes 128 B lines.
#pragma once
#include "cost.cuh"

namespace sde4 {

struct RolloutK {
  const float* params;
  const float* y0;
  const float* u;
  long long u_sp, u_st, u_sj;
  const float* dt;
  const uint32_t* key;
  long long key_sp;
  const float* noise;
  float* xevol;
  int P, N, H, G;
  int x_rows, rng_mode, noise_mode;
};

struct CostK {
  const float* params;
  const float* y0;
  const float* opt;
  long long o_sp, o_st, o_sj;
  const float* dt;
  const float* xref;
  long long xref_sp;
  const uint32_t* key;
  long long key_sp;
  const float* noise;  // given noise or nullptr
  float* xbuf;         // [H+1][x_rows][G] checkpoints (user xtraj or workspace) or nullptr (value only)
  float* dwbuf;        // [H][NX][G] generated noise kept for the backward sweep
  float* auxbuf;       // [H][G] projection aux
  float* gpart;        // [H][n_opt][PW] per-sample / per-warp gradient partials
  float* cpart;        // [PW] cost partials
  int P, N, H, G;
  int x_rows, rng_mode, noise_mode;
  int n_slack, warp_reduce, PW, has_terminal, write_stage_cost;
};

struct ReduceK {
  const float* gpart;
  const float* cpart;
  const float* opt;
  long long o_sp, o_st, o_sj;
  const float* dt;
  float* cost;
  float* grad;
  int P, N, H, n_u, n_opt, PW, NP;  // NP partials per problem
  int has_slew;
  float res_mult;
  float slew[SDE4_MAX_NU];
};

// dw_t for one sample; components with zero prior diffusion are never generated.
template <int NX>
SDE4_DEV void draw_noise(const sde4_model_desc& d, int noise_mode, int rng_mode, const float* __restrict__ noise,
                         uint32_t b0, uint32_t b1, int t, int H, int G, int g, float (&dw)[NX]) {
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    dw[i] = 0.0f;
    if (d.noise_prior[i] != 0.0f) {
      if (noise_mode == SDE4_NOISE_GIVEN) {
        dw[i] = __ldg(noise + ((size_t)t * NX + i) * G + g);
      } else {
        dw[i] = bits_to_normal(jax_bits(b0, b1, (uint64_t)t * NX + i, (uint64_t)H * NX, rng_mode));
      }
    }
  }
}

// One Euler-Maruyama step in place: x <- proj(x + f dt + sigma dw)   (sde_solver.py:301-304)
template <class M>
SDE4_DEV void em_step(const float* __restrict__ W, const sde4_model_desc& d, float (&x)[M::NX],
                      const float (&u)[M::NU], const float (&dw)[M::NX], float dt, bool noisy, float& aux) {
  float f[M::NX];
  M::drift(W, d, x, u, f);
  if (noisy) {
    float sig[M::NX];
    Diffusion<M>::fwd(W, d, x, sig);
#pragma unroll
    for (int i = 0; i < M::NX; ++i) x[i] = fmaf(sig[i], dw[i], fmaf(f[i], dt, x[i]));
  } else {
#pragma unroll
    for (int i = 0; i < M::NX; ++i) x[i] = fmaf(f[i], dt, x[i]);
  }
  M::project(x, aux);
}

// ----------------------------------------------------------------------------
// K1: forward rollout
// ----------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(256) k_rollout(const __grid_constant__ sde4_model_desc d,
                                                  const __grid_constant__ RolloutK a) {
  using A = typename M::A;
  constexpr int NX = M::NX, NU = M::NU;
  extern __shared__ __align__(16) float W[];
  __shared__ __align__(8) uint64_t mbar;
  stage_params_bulk(W, a.params, A::PARAM_TOTAL * sizeof(float), &mbar);
  Diffusion<M>::derive(W);
  __syncthreads();

  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.G) return;
  const int p = g / a.N, n = g - p * a.N;
  float x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    x[i] = __ldg(a.y0 + (size_t)p * NX + i);
    a.xevol[(size_t)i * a.G + g] = x[i];
  }
  const bool noisy = a.noise_mode != SDE4_NOISE_NONE;
  uint32_t b0 = 0, b1 = 0;
  if (a.noise_mode == SDE4_NOISE_THREEFRY) {
    const uint32_t* k = a.key + (size_t)p * a.key_sp;
    brownian_key(__ldg(k), __ldg(k + 1), (uint32_t)n, (uint32_t)a.N, a.rng_mode, b0, b1);
  }
  const float* up = a.u + (size_t)p * a.u_sp;
  for (int t = 0; t < a.H; ++t) {
    float u[NU], dw[NX];
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = __ldg(up + (size_t)t * a.u_st + (size_t)j * a.u_sj);
    if (noisy) draw_noise<NX>(d, a.noise_mode, a.rng_mode, a.noise, b0, b1, t, a.H, a.G, g, dw);
    float aux;
    em_step<M>(W, d, x, u, dw, __ldg(a.dt + t), noisy, aux);
    float* out = a.xevol + (size_t)(t + 1) * a.x_rows * a.G + g;
#pragma unroll
    for (int i = 0; i < NX; ++i) out[(size_t)i * a.G] = x[i];
  }
}

// ----------------------------------------------------------------------------
// K2: rollout + cost (+ adjoint)
// ----------------------------------------------------------------------------
SDE4_DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <class M, bool GRAD, bool CONSTR>
__global__ void __launch_bounds__(256) k_cost(const __grid_constant__ sde4_model_desc d,
                                               const __grid_constant__ sde4_cost_desc cd,
                                               const __grid_constant__ sde4_cost_desc td,
                                               const __grid_constant__ CostK a) {
  using A = typename M::A;
  constexpr int NX = M::NX, NU = M::NU;
  extern __shared__ __align__(16) float W[];
  __shared__ __align__(8) uint64_t mbar;
  stage_params_bulk(W, a.params, A::PARAM_TOTAL * sizeof(float), &mbar);
  Diffusion<M>::derive(W);
  __syncthreads();

  const int g0 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = g0 < a.G;
  const int g = active ? g0 : a.G - 1;  // inactive lanes shadow a valid sample, contribute nothing
  const int p = g / a.N, n = g - p * a.N;
  const int G = a.G, H = a.H;
  const float* op = a.opt + (size_t)p * a.o_sp;
  const float* xr = a.xref + (size_t)p * a.xref_sp;
  const bool noisy = a.noise_mode != SDE4_NOISE_NONE;
  const bool keep = GRAD || a.xbuf != nullptr;

  float x[NX], x0s[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    x[i] = __ldg(a.y0 + (size_t)p * NX + i);
    x0s[i] = x[i];
  }
  uint32_t b0 = 0, b1 = 0;
  if (a.noise_mode == SDE4_NOISE_THREEFRY) {
    const uint32_t* k = a.key + (size_t)p * a.key_sp;
    brownian_key(__ldg(k), __ldg(k + 1), (uint32_t)n, (uint32_t)a.N, a.rng_mode, b0, b1);
  }

  // ---------------- forward sweep ----------------
  float u[NU], unext[NU];
#pragma unroll
  for (int j = 0; j < NU; ++j) u[j] = __ldg(op + (size_t)j * a.o_sj);
  float dummy_gx[NX], dummy_gu[NU], dummy_gs[NX];
  float cost = 0.0f;
  float c_prev;
  {
    c_prev = stage_cost<NX, NU, false>(cd, x, u, xr, true, 0.0f, dummy_gx, dummy_gu);
    if (CONSTR) c_prev += constr_cost<NX, false>(cd, x, x0s, 0.0f, dummy_gx, dummy_gs);  // slack_0 := x0 (nsde.py:1510)
    if (keep && active && a.xbuf) {
#pragma unroll
      for (int i = 0; i < NX; ++i) a.xbuf[(size_t)i * G + g] = x[i];
      if (a.write_stage_cost) a.xbuf[(size_t)NX * G + g] = c_prev;
    }
  }
  for (int t = 0; t < H; ++t) {
    const float dt = __ldg(a.dt + t);
    float dw[NX];
    if (noisy) {
      draw_noise<NX>(d, a.noise_mode, a.rng_mode, a.noise, b0, b1, t, H, G, g, dw);
      if (GRAD && active && a.dwbuf) {
#pragma unroll
        for (int i = 0; i < NX; ++i)
          if (d.noise_prior[i] != 0.0f) a.dwbuf[((size_t)t * NX + i) * G + g] = dw[i];
      }
    }
    float aux = 0.0f;
    em_step<M>(W, d, x, u, dw, dt, noisy, aux);
    if (GRAD && M::NAUX && active) a.auxbuf[(size_t)t * G + g] = aux;
    // control of the next stage: u_{t+1}, the last one repeated (nsde.py:1509)
    const int tn = t + 1 < H ? t + 1 : H - 1;
#pragma unroll
    for (int j = 0; j < NU; ++j) unext[j] = __ldg(op + (size_t)tn * a.o_st + (size_t)j * a.o_sj);
    float c_next = stage_cost<NX, NU, false>(cd, x, unext, xr + (size_t)(t + 1) * NX, true, 0.0f, dummy_gx, dummy_gu);
    if (CONSTR) {
      float sl[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i)
        sl[i] = (cd.slack_col[i] >= 0 && cd.constr_mode == SDE4_CONSTR_WITHPROX)
                    ? __ldg(op + (size_t)t * a.o_st + (size_t)(NU + cd.slack_col[i]) * a.o_sj)
                    : 0.0f;
      c_next += constr_cost<NX, false>(cd, x, sl, 0.0f, dummy_gx, dummy_gs);
    }
    cost = fmaf(0.5f * dt, c_prev + c_next, cost);  // trapezoid (nsde.py:1515)
    c_prev = c_next;
    if (keep && active && a.xbuf) {
      float* out = a.xbuf + (size_t)(t + 1) * a.x_rows * G + g;
#pragma unroll
      for (int i = 0; i < NX; ++i) out[(size_t)i * G] = x[i];
      if (a.write_stage_cost) out[(size_t)NX * G] = c_next;
    }
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = unext[j];
  }
  if (a.has_terminal) cost += stage_cost<NX, NU, false>(td, x, u, xr + (size_t)H * NX, false, 0.0f, dummy_gx, dummy_gu);

  // cost partial
  {
    float cv = active ? cost : 0.0f;
    if (a.warp_reduce) {
      cv = warp_sum(cv);
      if ((threadIdx.x & 31) == 0 && active) a.cpart[g >> 5] = cv;
    } else if (active) {
      a.cpart[g] = cv;
    }
  }
  if (!GRAD) return;

  // ---------------- backward sweep ----------------
  // x holds x_H, u holds u_{H-1}
  float lam[NX], gu_pend[NU], gs_pend[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    lam[i] = 0.0f;
    gs_pend[i] = 0.0f;
  }
#pragma unroll
  for (int j = 0; j < NU; ++j) gu_pend[j] = 0.0f;
  {
    const float wH = 0.5f * __ldg(a.dt + H - 1);
    stage_cost<NX, NU, true>(cd, x, u, xr + (size_t)H * NX, true, wH, lam, gu_pend);
    if (CONSTR) {
      float sl[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i)
        sl[i] = (cd.slack_col[i] >= 0 && cd.constr_mode == SDE4_CONSTR_WITHPROX)
                    ? __ldg(op + (size_t)(H - 1) * a.o_st + (size_t)(NU + cd.slack_col[i]) * a.o_sj)
                    : 0.0f;
      constr_cost<NX, true>(cd, x, sl, wH, lam, gs_pend);
    }
    if (a.has_terminal) stage_cost<NX, NU, true>(td, x, u, xr + (size_t)H * NX, false, 1.0f, lam, gu_pend);
  }
  float xn[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) xn[i] = x[i];

  for (int t = H - 1; t >= 0; --t) {
    const float dt = __ldg(a.dt + t);
    // checkpoint x_t, control u_t, noise dw_t
    {
      const float* in = a.xbuf + (size_t)t * a.x_rows * G + g;
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = in[(size_t)i * G];
    }
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = __ldg(op + (size_t)t * a.o_st + (size_t)j * a.o_sj);
    float gu[NU], gs[NX], gx[NX];
#pragma unroll
    for (int j = 0; j < NU; ++j) gu[j] = gu_pend[j];
#pragma unroll
    for (int j = 0; j < NU; ++j) gu_pend[j] = 0.0f;
#pragma unroll
    for (int i = 0; i < NX; ++i) gs[i] = gs_pend[i];
    // through the projection
    float aux = 0.0f;
    if (M::NAUX) aux = a.auxbuf[(size_t)t * G + g];
    M::project_vjp(xn, aux, lam);
    // through x + f dt + sigma dw
    float vf[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      gx[i] = lam[i];
      vf[i] = dt * lam[i];
    }
    M::drift_vjp(W, d, x, u, vf, gx, gu);
    if (noisy) {
      float vs[NX];
      const float* nz = a.noise_mode == SDE4_NOISE_GIVEN ? a.noise : a.dwbuf;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        float dwi = 0.0f;
        if (d.noise_prior[i] != 0.0f) dwi = nz[((size_t)t * NX + i) * G + g];
        vs[i] = dwi * lam[i];
      }
      Diffusion<M>::vjp(W, d, x, vs, gx);
    }
    // stage cost at t: trapezoid weight w_t
    const float wt = t > 0 ? 0.5f * (__ldg(a.dt + t - 1) + dt) : 0.5f * dt;
    stage_cost<NX, NU, true>(cd, x, u, xr + (size_t)t * NX, true, wt, gx, gu);
    if (CONSTR) {
      float sl[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        sl[i] = 0.0f;
        if (cd.slack_col[i] >= 0 && cd.constr_mode == SDE4_CONSTR_WITHPROX)
          sl[i] = t > 0 ? __ldg(op + (size_t)(t - 1) * a.o_st + (size_t)(NU + cd.slack_col[i]) * a.o_sj) : x0s[i];
      }
      constr_cost<NX, true>(cd, x, sl, wt, gx, gs_pend);  // slack of stage t lives in opt row t-1
    }
    // emit row t of the per-sample gradient
    {
      float* gp = a.gpart + (size_t)t * (NU + a.n_slack) * a.PW;
      const int pw = a.warp_reduce ? (g >> 5) : g;
      const bool writer = active && (!a.warp_reduce || (threadIdx.x & 31) == 0);
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        float v = active ? gu[j] : 0.0f;
        if (a.warp_reduce) v = warp_sum(v);
        if (writer) gp[(size_t)j * a.PW + pw] = v;
      }
      if (CONSTR) {
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          if (cd.slack_col[i] >= 0 && cd.constr_mode == SDE4_CONSTR_WITHPROX) {
            float v = active ? gs[i] : 0.0f;
            if (a.warp_reduce) v = warp_sum(v);
            if (writer) gp[(size_t)(NU + cd.slack_col[i]) * a.PW + pw] = v;
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      lam[i] = gx[i];
      xn[i] = x[i];
    }
  }
}

}  // namespace sde4

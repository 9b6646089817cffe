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
  int n_off, n_tot;
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
  int n_off, n_tot;
};

struct ReduceK {
  const float* gpart;
  const float* cpart;
  const float* opt;
  long long o_sp, o_st, o_sj;
  const float* dt;
  float* cost;
  float* grad;
  int P, N, H, n_u, n_opt, PW, NP;  // NP partials per problem; N = total particles (divisor)
  int has_slew;
  float res_mult;
  float slew[SDE4_MAX_NU];
};

// dw_t for one sample, written to the per-thread scratch column `dwc` (element i at dwc[i*bs]);
// components with zero prior diffusion are never generated.  ROLLED over the state index (the
// threefry block is ~100 instructions; unrolling it 13x only bloats the instruction stream).
// `keep` (optional) receives a copy for the backward sweep: keep[i*G] = dw_i.
template <int NX>
SDE4_DEV void draw_noise(const sde4_model_desc& d, int noise_mode, int rng_mode, const float* __restrict__ noise,
                         uint32_t b0, uint32_t b1, int t, int H, int G, int g, float* dwc, int bs, float* keep) {
  // 4 elements per iteration: the threefry rounds are one serial dependency chain per element, so
  // independent elements are interleaved for ILP (a single resident warp cannot hide it otherwise)
#pragma unroll 1
  for (int i0 = 0; i0 < NX; i0 += 4) {
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + q;
      v[q] = 0.0f;
      if (i < NX && d.noise_prior[i < NX ? i : 0] != 0.0f) {
        if (noise_mode == SDE4_NOISE_GIVEN) {
          v[q] = __ldg(noise + ((size_t)t * NX + i) * G + g);
        } else {
          v[q] = bits_to_normal(jax_bits(b0, b1, (uint64_t)t * NX + i, (uint64_t)H * NX, rng_mode));
          if (keep) keep[(size_t)i * G] = v[q];
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (i0 + q < NX) dwc[(size_t)(i0 + q) * bs] = v[q];
  }
}

// One Euler-Maruyama step in place: x <- proj(x + f dt + sigma dw)   (sde_solver.py:301-304)
template <class M>
SDE4_DEV void em_step(const float* W, const Scratch& sc, const sde4_model_desc& d, float (&x)[M::NX],
                      const float (&u)[M::NU], float dt, bool noisy, float& aux) {
  float f[M::NX];
  M::drift(W, sc, d, x, u, f);
  if (noisy) {
    float sig[M::NX];
    Diffusion<M>::fwd(W, sc, d, x, sig);
    const float* dwc = sc.slot(Scratch::SLOT_X);
#pragma unroll
    for (int i = 0; i < M::NX; ++i) x[i] = fmaf(sig[i], dwc[(size_t)i * sc.bs], fmaf(f[i], dt, x[i]));
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
  const Scratch sc{W + A::SMEM_FLOATS + threadIdx.x, (int)blockDim.x};

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
    brownian_key(__ldg(k), __ldg(k + 1), (uint32_t)(n + a.n_off), (uint32_t)a.n_tot, a.rng_mode, b0, b1);
  }
  const float* up = a.u + (size_t)p * a.u_sp;
  for (int t = 0; t < a.H; ++t) {
    float u[NU];
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = __ldg(up + (size_t)t * a.u_st + (size_t)j * a.u_sj);
    if (noisy)
      draw_noise<NX>(d, a.noise_mode, a.rng_mode, a.noise, b0, b1, t, a.H, a.G, g, sc.slot(Scratch::SLOT_X), sc.bs,
                     nullptr);
    float aux;
    em_step<M>(W, sc, d, x, u, __ldg(a.dt + t), noisy, aux);
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
  const Scratch sc{W + A::SMEM_FLOATS + threadIdx.x, (int)blockDim.x};

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
    brownian_key(__ldg(k), __ldg(k + 1), (uint32_t)(n + a.n_off), (uint32_t)a.n_tot, a.rng_mode, b0, b1);
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
    if (noisy)
      draw_noise<NX>(d, a.noise_mode, a.rng_mode, a.noise, b0, b1, t, H, G, g, sc.slot(Scratch::SLOT_X), sc.bs,
                     (GRAD && active && a.dwbuf) ? a.dwbuf + (size_t)t * NX * G + g : nullptr);
    float aux = 0.0f;
    em_step<M>(W, sc, d, x, u, dt, noisy, aux);
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
  // x holds x_H, u holds u_{H-1}.  Live state is kept small on purpose (x, lam, gx, u, gu): the
  // projected state and the noise are re-read from the checkpoint buffers where needed.
  float lam[NX], gu[NU], gs_pend[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    lam[i] = 0.0f;
    gs_pend[i] = 0.0f;
  }
#pragma unroll
  for (int j = 0; j < NU; ++j) gu[j] = 0.0f;  // row H-1 also receives the u-gradient of stage H
  {
    const float wH = 0.5f * __ldg(a.dt + H - 1);
    stage_cost<NX, NU, true>(cd, x, u, xr + (size_t)H * NX, true, wH, lam, gu);
    if (CONSTR) {
      float sl[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i)
        sl[i] = (cd.slack_col[i] >= 0 && cd.constr_mode == SDE4_CONSTR_WITHPROX)
                    ? __ldg(op + (size_t)(H - 1) * a.o_st + (size_t)(NU + cd.slack_col[i]) * a.o_sj)
                    : 0.0f;
      constr_cost<NX, true>(cd, x, sl, wH, lam, gs_pend);
    }
    if (a.has_terminal) {
      float gu_unused[NU];
      stage_cost<NX, NU, true>(td, x, u, xr + (size_t)H * NX, false, 1.0f, lam, gu_unused);
    }
  }
  const float* nz = a.noise_mode == SDE4_NOISE_GIVEN ? a.noise : a.dwbuf;

  for (int t = H - 1; t >= 0; --t) {
    const float dt = __ldg(a.dt + t);
    // through the projection (needs the projected x_{t+1}, still in the checkpoint buffer)
    {
      float aux = 0.0f;
      if (M::NAUX) aux = a.auxbuf[(size_t)t * G + g];
      M::project_vjp(a.xbuf + (size_t)(t + 1) * a.x_rows * G + g, (size_t)G, aux, lam);
    }
    // checkpoint x_t and control u_t
    {
      const float* in = a.xbuf + (size_t)t * a.x_rows * G + g;
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = in[(size_t)i * G];
    }
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = __ldg(op + (size_t)t * a.o_st + (size_t)j * a.o_sj);
    float gs[NX], gx[NX];
    if (CONSTR) {
#pragma unroll
      for (int i = 0; i < NX; ++i) gs[i] = gs_pend[i];
    }
    // through x + f dt + sigma dw
#pragma unroll
    for (int i = 0; i < NX; ++i) gx[i] = lam[i];
    {
      const Scaled<NX> vf{lam, dt};
      M::drift_vjp(W, sc, d, x, u, vf, gx, gu);
    }
    if (noisy) {
      const float* nzt = nz + (size_t)t * NX * G + g;
      Diffusion<M>::vjp(W, sc, d, x, lam, [&](int i) { return nzt[(size_t)i * G]; }, gx);
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
        gu[j] = 0.0f;
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
    for (int i = 0; i < NX; ++i) lam[i] = gx[i];
  }
}

}  // namespace sde4

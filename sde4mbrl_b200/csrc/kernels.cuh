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
#pragma nv_diag_suppress 128  // the value-only instantiation returns before the backward sweep

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
  int solver;  // sde4_solver
  // simulator summary (sde4_sim_rollout): xevol may be null, the last state goes to last[i*G + g], the accumulated
  // reward of the sample to racc[g]
  float* last;
  float* racc;
  int reward_kind;
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
  const int32_t* mask;  // optional [P]: problems with mask 0 are skipped
  int data_mod;  // y0 / xref / key of problem p are those of (p mod data_mod)
  // training loss (sde4_loss_grad): plain sum over t >= 1 instead of the trapezoid, training key chain,
  // parameter-gradient streams of the three networks ([rows][H*G], column = t*G + g)
  int sum_mode, rng_chain;
  float* pd_net[3];
  const float* step_w;  // training loss only: optional [H][G] weight of the data term at t+1 (num_sample2consider)
  float* pd_phys;  // [NPHYS][pd_stride]: per-sample cotangents of the physical scalars in column g (t = 0)
  unsigned long long pd_stride;
};

struct XgpuK {  // sde4_xgpu_reduce, by value in the kernel parameters
  int world, rank, n;
  unsigned epoch;
  float* buf[8];
  unsigned* flag[8];
  float* out;
  unsigned* counter;
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
  const int32_t* mask;  // optional [P]: skipped problems keep their old outputs
  int lpi;  // lanes per reduced item: 1 or 32
  float res_mult;
  float slew[SDE4_MAX_NU];
  XgpuK x;  // x.world > 1: the last CTA sums the world's partial buffers (see sde4_xgpu_reduce)
};

// ------------------------------------------------------------------------------------------------
// Per-thread kernel context: which sample this thread works on, its lane within the sample's lane
// group (EvL kernels) and the evaluator context.
// ------------------------------------------------------------------------------------------------
template <class M>
struct Tctx {
  static constexpr int L = M::Ev::L;
  typename M::Ev::Ctx ev;
  int l;  // lane within the sample (0 for one-thread-per-sample kernels)
};

// Prologue shared by K1/K2: stage the parameter image with one bulk async copy, derive the
// state-independent constants and set up the per-thread context.  Returns the thread's global sample slot (may be >= G: inactive).
template <class M>
SDE4_DEV int kernel_prologue(float* W, const float* params, uint64_t* mbar, Tctx<M>& tc) {
  using A = typename M::A;
  constexpr int L = M::Ev::L;
  stage_params_bulk(W, params, A::PARAM_TOTAL * sizeof(float), mbar);
  Diffusion<M>::derive(W);
  __syncthreads();
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  if constexpr (L == 1) {
    tc.ev = Scratch{W + A::smem_floats(1) + threadIdx.x, (int)blockDim.x};
    tc.l = 0;
    return gt;
  } else {
    tc.l = threadIdx.x & (L - 1);
    tc.ev.l = tc.l;
    return gt / L;
  }
}

// dw_t for one sample.  Components with zero prior diffusion are never generated.
//   one thread per sample: ROLLED over the state index in groups of 4 (ILP across the serial
//     threefry chains), values parked in the thread's scratch column;
//   lane-split: lane l generates elements l, l+L, ... and the group exchanges them by SHFL.
// `keep` (optional) receives a copy for the backward sweep: keep[i*G] = dw_i.
template <class M>
SDE4_DEV void draw_noise(const sde4_model_desc& d, int noise_mode, int rng_mode, const float* __restrict__ noise,
                         uint32_t b0, uint32_t b1, int t, int H, unsigned G, int g, uint32_t nzmask, Tctx<M>& tc,
                         float (&dw)[M::NX],
                         float* keep) {
  constexpr int NX = M::NX, L = M::Ev::L;
  if constexpr (L == 1) {
    float* dwc = tc.ev.slot(Scratch::SLOT_X);
    const int bs = tc.ev.bs;
#pragma unroll 1
    for (int i0 = 0; i0 < NX; i0 += 4) {
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + q;
        v[q] = 0.0f;
        if (i < NX && ((nzmask >> i) & 1u)) {
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
#pragma unroll
    for (int i = 0; i < NX; ++i) dw[i] = dwc[(size_t)i * bs];
  } else {
    constexpr int PER = (NX + L - 1) / L;
    float v[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int i = tc.l + q * L;
      v[q] = 0.0f;
      if (i < NX && ((nzmask >> i) & 1u)) {
        if (noise_mode == SDE4_NOISE_GIVEN) {
          v[q] = __ldg(noise + ((size_t)t * NX + i) * G + g);
        } else {
          v[q] = bits_to_normal(jax_bits(b0, b1, (uint64_t)t * NX + i, (uint64_t)H * NX, rng_mode));
          if (keep) keep[(size_t)i * G] = v[q];
        }
      }
    }
    const int base = (threadIdx.x & 31) & ~(L - 1);
#pragma unroll
    for (int i = 0; i < NX; ++i) dw[i] = __shfl_sync(0xffffffffu, v[i / L], base + (i % L));
  }
}

// ---- lane-split kernels with in-kernel threefry noise (template flag TFP): the noise of step t+1 is generated
// while step t is evaluated.  Everything below is branch free, so that the integer threefry chain and the
// floating-point step end up in ONE basic block and ptxas interleaves them (the ALU and FMA pipes run side by side).
template <class M>
struct NoiseLanes {
  static constexpr int NX = M::NX, L = M::Ev::L, PER = (NX + L - 1) / L;
  // lane l's elements l, l+L, ... of dw_t
  SDE4_DEV static void gen(uint32_t b0, uint32_t b1, int rng_mode, uint32_t nzmask, int t, int H, int l,
                           float (&v)[PER]) {
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int i = l + q * L;
      const bool valid = i < NX && ((nzmask >> i) & 1u);
      const int ii = valid ? i : 0;
      const float z = bits_to_normal(jax_bits(b0, b1, (uint64_t)t * NX + ii, (uint64_t)H * NX, rng_mode));
      v[q] = valid ? z : 0.0f;
    }
  }
  SDE4_DEV static void distribute(const float (&v)[PER], float (&dw)[NX]) {
    const int base = (threadIdx.x & 31) & ~(L - 1);
#pragma unroll
    for (int i = 0; i < NX; ++i) dw[i] = __shfl_sync(0xffffffffu, v[i / L], base + (i % L));
  }
  // keep[i*G] = dw_i for the backward sweep, written by the owning lane
  SDE4_DEV static void keep(const float (&v)[PER], float* kp, unsigned G, uint32_t nzmask, int l, bool active) {
    float* row = kp + (size_t)l * G;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int i = l + q * L;
      if (active && i < NX && ((nzmask >> i) & 1u)) row[(size_t)(q * L) * G] = v[q];
    }
  }
};

// One Euler-Maruyama step in place: x <- proj(x + f dt + sigma dw)   (sde_solver.py:301-304)
template <class M>
SDE4_DEV void em_step(const float* W, Tctx<M>& tc, const sde4_model_desc& d, float (&x)[M::NX],
                      const float (&u)[M::NU], const float (&dw)[M::NX], float dt, bool noisy, float& aux) {
  float f[M::NX];
  M::drift(W, tc.ev, d, x, u, f);
  if (noisy) {
    float sig[M::NX];
    Diffusion<M>::fwd(W, tc.ev, d, x, u, sig);
#pragma unroll
    for (int i = 0; i < M::NX; ++i) x[i] = fmaf(sig[i], dw[i], fmaf(f[i], dt, x[i]));
  } else {
#pragma unroll
    for (int i = 0; i < M::NX; ++i) x[i] = fmaf(f[i], dt, x[i]);
  }
  M::project(x, aux);
}

// x[base + l] for l in [0, L) by a select tree over the bits of l (all indices compile-time)
template <int L, int N>
SDE4_DEV float lane_pick(const float (&x)[N], int base, int l) {
  float v[L];
#pragma unroll
  for (int j = 0; j < L; ++j) v[j] = x[base + j < N ? base + j : N - 1];
#pragma unroll
  for (int s = 1; s < L; s <<= 1) {
    const bool bit = (l & s) != 0;
#pragma unroll
    for (int j = 0; j + s < L; j += 2 * s) v[j] = bit ? v[j + s] : v[j];
  }
  return v[0];
}

// The two Stratonovich schemes of the reference's commented-out code, offered as an extension of K1 (xi = the
// step's standard normals).  Heun (sde_solver.py:37-50): dW = sqrt(dt) xi, predictor xb = x + f0 dt + g0 dW,
// x+ = x + (f0+f1)/2 dt + (g0+g1)/2 dW.  Derivative-free Milstein (:146-169): xb = x + f0 dt + g0 sqrt(dt),
// x+ = x + f0 dt + g0 sqrt(dt) xi + (0.5/sqrt(dt)) (g(xb) - g0) xi^2.  Projection after the step, as in the comments.
template <class M>
SDE4_DEV void strat_step(int solver, const float* W, Tctx<M>& tc, const sde4_model_desc& d, float (&x)[M::NX],
                         const float (&u)[M::NU], const float (&xi)[M::NX], float dt, bool noisy, float& aux) {
  constexpr int NX = M::NX;
  float f0[NX], g0[NX], xb[NX];
  M::drift(W, tc.ev, d, x, u, f0);
#pragma unroll
  for (int i = 0; i < NX; ++i) g0[i] = 0.0f;
  if (noisy) Diffusion<M>::fwd(W, tc.ev, d, x, u, g0);
  const float sq = sqrtf(dt);
  if (solver == SDE4_SOLVER_STRAT_HEUN) {
#pragma unroll
    for (int i = 0; i < NX; ++i) xb[i] = fmaf(g0[i], noisy ? sq * xi[i] : 0.0f, fmaf(f0[i], dt, x[i]));
    float f1[NX], g1[NX];
    M::drift(W, tc.ev, d, xb, u, f1);
#pragma unroll
    for (int i = 0; i < NX; ++i) g1[i] = 0.0f;
    if (noisy) Diffusion<M>::fwd(W, tc.ev, d, xb, u, g1);
#pragma unroll
    for (int i = 0; i < NX; ++i)
      x[i] = fmaf(0.5f * (g0[i] + g1[i]), noisy ? sq * xi[i] : 0.0f, fmaf(0.5f * (f0[i] + f1[i]), dt, x[i]));
  } else {
    float g1[NX], xt[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      xt[i] = fmaf(f0[i], dt, x[i]);
      xb[i] = fmaf(g0[i], sq, xt[i]);
      g1[i] = 0.0f;
    }
    if (noisy) Diffusion<M>::fwd(W, tc.ev, d, xb, u, g1);
    const float c = 0.5f / sq;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const float z = noisy ? xi[i] : 0.0f;
      x[i] = fmaf(c * (g1[i] - g0[i]), z * z, fmaf(g0[i], sq * z, xt[i]));
    }
  }
  M::project(x, aux);
}

// state rows of one time slice: element i is written by lane (i mod L) of the sample's group (the state is
// replicated over the group; every lane selects its ceil(NX/L) elements and issues that many stores)
template <class M>
SDE4_DEV void store_state(float* out, unsigned G, const float (&x)[M::NX], int l, bool active) {
  constexpr int L = M::Ev::L, NX = M::NX;
  if constexpr (L == 1) {
#pragma unroll
    for (int i = 0; i < NX; ++i)
      if (active) out[(size_t)i * G] = x[i];
  } else {
    float* row = out + (size_t)l * G;
#pragma unroll
    for (int k = 0; k * L < NX; ++k) {
      const float v = lane_pick<L, NX>(x, k * L, l);
      if (active && k * L + l < NX) row[(size_t)(k * L) * G] = v;
    }
  }
}

// bit i set <=> state i has a non-zero prior diffusion (its noise sample exists)
template <int NX>
SDE4_DEV uint32_t noise_mask(const sde4_model_desc& d) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < NX; ++i) m |= d.noise_prior[i] != 0.0f ? (1u << i) : 0u;
  return m;
}

// mbrl-lib reward / termination of the cartpole swing-up task on the SDE state [x, xd, th, thd]
// (mbrlLibUtils/reward_functions.py:3-9 on obs = [x, xd, sin th, cos th, thd]: cos(obs[3]) - 0.1 |obs[0]| -- sic, the
// cosine of cos th; termination_functions.py:3-14: done when |x| >= 120)
template <int NX>
SDE4_HD float sim_reward_cartpole(const float (&x)[NX]) {
  return cosf(cosf(x[2])) - 0.1f * fabsf(x[0]);
}
template <int NX>
SDE4_HD bool sim_done_cartpole(const float (&x)[NX]) {
  return !(x[0] > -120.0f && x[0] < 120.0f);
}

// ----------------------------------------------------------------------------
// K1: forward rollout
// ----------------------------------------------------------------------------
// STRAT: the instantiation that carries the Stratonovich extension steps (kept out of the Euler-Maruyama kernels:
// the extra inlined drift / diffusion evaluations cost the throughput kernels 20 % through code size alone)
// SIM: the simulator-summary instantiation (sde4_sim_rollout): trajectory stores only on request, per-step reward with
// termination, last state / accumulated reward of every sample at the end.  Kept out of the plain rollout kernels: the
// few extra instructions per step cost the 2^20-start-state rollout 5 % when they were run-time branches of one kernel.
template <class M, bool TFP = false, bool STRAT = false, bool SIM = false>
__global__ void __launch_bounds__(256) k_rollout(const __grid_constant__ sde4_model_desc d,
                                                  const __grid_constant__ RolloutK a) {
  constexpr int NX = M::NX, NU = M::NU;
  static_assert(!TFP || M::Ev::L > 1, "the pipelined-noise variant is a lane-split kernel");
  static_assert(!(TFP && STRAT), "the extension steps use the generic noise path");
  extern __shared__ __align__(16) float W[];
  __shared__ __align__(8) uint64_t mbar;
  Tctx<M> tc;
  const int g0 = kernel_prologue<M>(W, a.params, &mbar, tc);
  const bool active = g0 < a.G;
  const int g = active ? g0 : a.G - 1;  // inactive lanes shadow a valid sample (warp collectives stay uniform)
  const int p = g / a.N, n = g - p * a.N;
  float x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = __ldg(a.y0 + (size_t)p * NX + i);
  const bool traj = !SIM || a.xevol != nullptr;  // warp uniform
  if (traj) store_state<M>(a.xevol + g, a.G, x, tc.l, active);
  [[maybe_unused]] float racc = 0.0f;
  [[maybe_unused]] bool alive = true;
  const bool noisy = TFP ? true : a.noise_mode != SDE4_NOISE_NONE;
  const uint32_t nzmask = noise_mask<NX>(d);
  uint32_t b0 = 0, b1 = 0;
  if (TFP || a.noise_mode == SDE4_NOISE_THREEFRY) {
    const uint32_t* k = a.key + (size_t)p * a.key_sp;
    brownian_key(__ldg(k), __ldg(k + 1), (uint32_t)(n + a.n_off), (uint32_t)a.n_tot, a.rng_mode, b0, b1);
  }
  const float* up = a.u + (size_t)p * a.u_sp;
  using NL = NoiseLanes<M>;
  float vn[NL::PER];
  if constexpr (TFP) NL::gen(b0, b1, a.rng_mode, nzmask, 0, a.H, tc.l, vn);
  for (int t = 0; t < a.H; ++t) {
    float u[NU], dw[NX];
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = __ldg(up + (size_t)t * a.u_st + (size_t)j * a.u_sj);
    if constexpr (TFP) {
      NL::distribute(vn, dw);
      NL::gen(b0, b1, a.rng_mode, nzmask, t + 1 < a.H ? t + 1 : t, a.H, tc.l, vn);  // overlaps the step below
    } else {
      if (noisy) draw_noise<M>(d, a.noise_mode, a.rng_mode, a.noise, b0, b1, t, a.H, (unsigned)a.G, g, nzmask, tc, dw, nullptr);
    }
    float aux;
    if constexpr (STRAT) strat_step<M>(a.solver, W, tc, d, x, u, dw, __ldg(a.dt + t), noisy, aux);
    else em_step<M>(W, tc, d, x, u, dw, __ldg(a.dt + t), noisy, aux);
    if (traj) store_state<M>(a.xevol + (size_t)(t + 1) * a.x_rows * a.G + g, a.G, x, tc.l, active);
    if constexpr (SIM && NX == 4) {  // cartpole: reward of the step, counted up to and including the terminal one
      if (a.reward_kind == SDE4_REWARD_CARTPOLE_SWINGUP) {
        const float r = sim_reward_cartpole(x);
        racc += alive ? r : 0.0f;
        alive = alive && !sim_done_cartpole(x);
      }
    }
  }
  if constexpr (SIM) {  // the last state (and the accumulated reward) of every sample
    if (active && tc.l == 0) {
#pragma unroll
      for (int i = 0; i < NX; ++i) a.last[(size_t)i * a.G + g] = x[i];
      if (a.racc) a.racc[g] = racc;
    }
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

// MODE 0: generic.  MODE 1 (lane-split only): in-kernel threefry noise generated one step ahead, branch free.
// MODE 2: MODE 1 + the model's plain (unshaped) stage cost, also branch free.
// MODE 3: generic noise path + the plain stage cost (one-thread-per-sample kernels: less code, fewer branches).
// The multirotor one-thread-per-sample gradient kernel is compiled for 3 CTAs of 128 threads per SM (168 registers +
// local-memory spills instead of 255 + spills; its four scratch columns allow exactly 3 CTAs): 12 instead of 8 warps
// hide more of the shared-memory latency of the rolled dense loops (C4 / N=32 gradient launch 2.06 -> 1.96 ms).
template <class M, bool GRAD, bool CONSTR, int MODE = 0>
__global__ void __launch_bounds__((M::Ev::L == 1 && GRAD && !M::Ev::PGRAD && M::NX == 13) ? 384 : 256) k_cost(const __grid_constant__ sde4_model_desc d,
                                               const __grid_constant__ sde4_cost_desc cd,
                                               const __grid_constant__ sde4_cost_desc td,
                                               const __grid_constant__ CostK a) {
  constexpr int NX = M::NX, NU = M::NU, L = M::Ev::L, SPW = 32 / L;  // SPW: samples per warp
  extern __shared__ __align__(16) float W[];
  __shared__ __align__(8) uint64_t mbar;
  Tctx<M> tc;
  const int g0 = kernel_prologue<M>(W, a.params, &mbar, tc);
  const bool active = g0 < a.G;
  const int g = active ? g0 : a.G - 1;  // inactive lanes shadow a valid sample, contribute nothing
  const bool contrib = active && tc.l == 0;  // one lane per sample feeds the particle reductions
  const int p = g / a.N, n = g - p * a.N;
  const unsigned G = (unsigned)a.G;
  const int H = a.H;
  const int pd = p % a.data_mod;  // stacked candidate evaluations share the problem data
  if (a.mask) {  // skipped problems: a warp leaves when none of its samples is wanted (warp-uniform exit; a partly
    const bool on = __ldg(a.mask + p) != 0;  // wanted warp computes everything, its unwanted results are ignored)
    if (!__any_sync(0xffffffffu, on)) return;
  }
  const float* op = a.opt + (size_t)p * a.o_sp;
  const float* xr = a.xref + (size_t)pd * a.xref_sp;
  constexpr bool TFP = MODE == 1 || MODE == 2, PLAINC = MODE >= 2;
  static_assert(!TFP || L > 1, "the pipelined-noise variant is a lane-split kernel");
  const bool noisy = TFP ? true : a.noise_mode != SDE4_NOISE_NONE;
  const uint32_t nzmask = noise_mask<NX>(d);
  const bool keep = GRAD || a.xbuf != nullptr;

  float x[NX], x0s[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    x[i] = __ldg(a.y0 + (size_t)pd * NX + i);
    x0s[i] = x[i];
  }
  uint32_t b0 = 0, b1 = 0;
  if (TFP || a.noise_mode == SDE4_NOISE_THREEFRY) {
    const uint32_t* k = a.key + (size_t)pd * a.key_sp;
    brownian_key(__ldg(k), __ldg(k + 1), (uint32_t)(n + a.n_off), (uint32_t)a.n_tot, a.rng_mode, b0, b1, a.rng_chain);
  }

  // ---------------- forward sweep ----------------
  float u[NU], unext[NU];
#pragma unroll
  for (int j = 0; j < NU; ++j) u[j] = __ldg(op + (size_t)j * a.o_sj);
  float dummy_gx[NX], dummy_gu[NU], dummy_gs[NX];
  float cost = 0.0f;
  float c_prev = 0.0f;
  // With GRAD the stage costs are evaluated once, in the backward sweep (value and gradient together);
  // the value-only kernel accumulates the trapezoid here, in the reference's order.
  {
    if constexpr (!GRAD) {
      c_prev = stage_cost<NX, NU, false, PLAINC>(cd, x, u, xr, true, 0.0f, dummy_gx, dummy_gu);
      if (CONSTR) c_prev += constr_cost<NX, false>(cd, x, x0s, 0.0f, dummy_gx, dummy_gs);  // slack_0 := x0 (nsde.py:1510)
    }
    if (keep && a.xbuf) {
      store_state<M>(a.xbuf + g, G, x, tc.l, active);
      if (!GRAD && a.write_stage_cost && contrib) a.xbuf[(size_t)NX * G + g] = c_prev;
    }
  }
  using NL = NoiseLanes<M>;
  float vn[NL::PER];
  if constexpr (TFP) NL::gen(b0, b1, a.rng_mode, nzmask, 0, H, tc.l, vn);
  for (int t = 0; t < H; ++t) {
    const float dt = __ldg(a.dt + t);
    float dw[NX];
    if constexpr (TFP) {
      NL::distribute(vn, dw);
      if constexpr (GRAD) NL::keep(vn, a.dwbuf + (size_t)t * NX * G + g, G, nzmask, tc.l, active);  // dwbuf is never null here
      NL::gen(b0, b1, a.rng_mode, nzmask, t + 1 < H ? t + 1 : t, H, tc.l, vn);  // overlaps the step below
    } else {
      if (noisy)
        draw_noise<M>(d, a.noise_mode, a.rng_mode, a.noise, b0, b1, t, H, G, g, nzmask, tc, dw,
                      (GRAD && active && a.dwbuf) ? a.dwbuf + (size_t)t * NX * G + g : nullptr);
    }
    float aux = 0.0f;
    em_step<M>(W, tc, d, x, u, dw, dt, noisy, aux);
    if (GRAD && M::NAUX && contrib) a.auxbuf[(size_t)t * G + g] = aux;
    // control of the next stage: u_{t+1}, the last one repeated (nsde.py:1509)
    const int tn = t + 1 < H ? t + 1 : H - 1;
#pragma unroll
    for (int j = 0; j < NU; ++j) unext[j] = __ldg(op + (size_t)tn * a.o_st + (size_t)j * a.o_sj);
    float c_next = 0.0f;
    if constexpr (!GRAD) {
      c_next = stage_cost<NX, NU, false, PLAINC>(cd, x, unext, xr + (size_t)(t + 1) * NX, true, 0.0f, dummy_gx, dummy_gu);
      if (CONSTR) {
        float sl[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i)
          sl[i] = (cd.slack_col[i] >= 0 && cd.constr_mode == SDE4_CONSTR_WITHPROX)
                      ? __ldg(op + (size_t)t * a.o_st + (size_t)(NU + cd.slack_col[i]) * a.o_sj)
                      : 0.0f;
        c_next += constr_cost<NX, false>(cd, x, sl, 0.0f, dummy_gx, dummy_gs);
      }
      if (a.sum_mode) cost += c_next;  // data term of the training loss: sum over t >= 1 (nsde.py:766)
      else cost = fmaf(0.5f * dt, c_prev + c_next, cost);  // trapezoid (nsde.py:1515)
      c_prev = c_next;
    }
    if (GRAD || (keep && a.xbuf)) {  // GRAD: the checkpoint buffer always exists (no branch in the loop body)
      float* out = a.xbuf + (size_t)(t + 1) * a.x_rows * G + g;
      store_state<M>(out, G, x, tc.l, active);
      if (!GRAD && a.write_stage_cost && contrib) out[(size_t)NX * G] = c_next;
    }
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = unext[j];
  }
  if (!GRAD && a.has_terminal)
    cost += stage_cost<NX, NU, false>(td, x, u, xr + (size_t)H * NX, false, 0.0f, dummy_gx, dummy_gu);

  // partial index of this thread's sample: per warp (all its samples in one problem) or per sample
  const int pw = a.warp_reduce ? (g / SPW) : g;
  const bool writer = a.warp_reduce ? (active && (threadIdx.x & 31) == 0) : contrib;
  auto emit_cost = [&]() {
    float cv = contrib ? cost : 0.0f;
    if (a.warp_reduce) cv = warp_sum(cv);
    if (writer) a.cpart[pw] = cv;
  };
  if (!GRAD) {
    emit_cost();
    return;
  }
  if constexpr (L > 1) __syncwarp();  // checkpoints written by sibling lanes are read back below

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
    float wH = a.sum_mode ? 1.0f : 0.5f * __ldg(a.dt + H - 1);
    if constexpr (M::Ev::PGRAD) {  // training instantiations only: the MPC kernels do not carry the load
      if (a.step_w) wH = __ldg(a.step_w + (size_t)(H - 1) * G + g);
    }
    float cH = stage_cost<NX, NU, true, PLAINC>(cd, x, u, xr + (size_t)H * NX, true, wH, lam, gu);
    if (CONSTR) {
      float sl[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i)
        sl[i] = (cd.slack_col[i] >= 0 && cd.constr_mode == SDE4_CONSTR_WITHPROX)
                    ? __ldg(op + (size_t)(H - 1) * a.o_st + (size_t)(NU + cd.slack_col[i]) * a.o_sj)
                    : 0.0f;
      cH += constr_cost<NX, true>(cd, x, sl, wH, lam, gs_pend);
    }
    cost = wH * cH;
    if (a.write_stage_cost && contrib) a.xbuf[((size_t)H * a.x_rows + NX) * G + g] = cH;
    if (a.has_terminal) {
      float gu_unused[NU];
      cost += stage_cost<NX, NU, true>(td, x, u, xr + (size_t)H * NX, false, 1.0f, lam, gu_unused);
    }
  }
  const float* nz = a.noise_mode == SDE4_NOISE_GIVEN ? a.noise : a.dwbuf;

  if constexpr (M::Ev::PGRAD) {
    tc.ev.dnet[0] = a.pd_net[0];
    tc.ev.dnet[1] = a.pd_net[1];
    tc.ev.dnet[2] = a.pd_net[2];
    tc.ev.dstride = a.pd_stride;
#pragma unroll
    for (int i = 0; i < 16; ++i) tc.ev.gS[i] = 0.0f;
  }
  // The lane-split kernels have registers to spare: they load the checkpoint x_t one iteration ahead so
  // that the L2 / HBM latency hides behind the previous step's arithmetic.
  constexpr bool PREFETCH = L > 1;
  float xpre[PREFETCH ? NX : 1], auxpre = 0.0f;
  if constexpr (PREFETCH) {
    const float* in = a.xbuf + (size_t)(H - 1) * a.x_rows * G + g;
#pragma unroll
    for (int i = 0; i < NX; ++i) xpre[i] = in[(size_t)i * G];
    if (M::NAUX) auxpre = a.auxbuf[(size_t)(H - 1) * G + g];
  }
  // software pipeline (straight-line lane-split kernels): the cotangent-independent half of the diffusion adjoint of
  // step t-1 is evaluated inside the block that applies step t
  constexpr bool PIPE = TFP && PREFETCH && !M::Ev::PGRAD && M::A::HAS_DENSITY;
  typename Diffusion<M>::Tape dtape, dtape_next;
  [[maybe_unused]] float upre[NU];  // controls of the pipelined step (read by the density net under control_dependent_noise only)
  if constexpr (PIPE) {
#pragma unroll
    for (int j = 0; j < NU; ++j) upre[j] = Diffusion<M>::CDN ? __ldg(op + (size_t)(H - 1) * a.o_st + (size_t)j * a.o_sj) : 0.0f;
    Diffusion<M>::prep(W, tc.ev, d, xpre, upre, dtape);
  }
  for (int t = H - 1; t >= 0; --t) {
    const float dt = __ldg(a.dt + t);
    if constexpr (M::Ev::PGRAD) tc.ev.dcol = (size_t)t * G + g;
    // noise of this step, requested before the long drift adjoint that precedes its use
    float nzv[PREFETCH ? NX : 1];
    if constexpr (PREFETCH) {
      if (noisy) {
        const float* nzt = nz + (size_t)t * NX * G + g;
#pragma unroll
        for (int i = 0; i < NX; ++i)
          if ((M::A::NOUT_MASK >> i) & 1u) nzv[i] = ((nzmask >> i) & 1u) ? nzt[(size_t)i * G] : 0.0f;
      }
    }
    // through the projection: x still holds the projected x_{t+1}
    {
      float aux = auxpre;
      if (!PREFETCH && M::NAUX) aux = a.auxbuf[(size_t)t * G + g];
      M::project_vjp(x, aux, lam);
    }
    // checkpoint x_t and control u_t
    if constexpr (PREFETCH) {
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = xpre[i];
      {  // unconditional (t = 0 re-reads row 0): no branch in the loop body
        const int tp = t > 0 ? t - 1 : 0;
        const float* in = a.xbuf + (size_t)tp * a.x_rows * G + g;
#pragma unroll
        for (int i = 0; i < NX; ++i) xpre[i] = in[(size_t)i * G];
        if (M::NAUX) auxpre = a.auxbuf[(size_t)tp * G + g];
      }
    } else {
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
      M::drift_vjp(W, tc.ev, d, x, u, vf, gx, gu);
    }
    if constexpr (PIPE) {
      Diffusion<M>::apply(W, d, x, lam, [&](int i) { return nzv[i]; }, dtape, gx, gu);
      if constexpr (Diffusion<M>::CDN) {
        const int tp = t > 0 ? t - 1 : 0;
#pragma unroll
        for (int j = 0; j < NU; ++j) upre[j] = __ldg(op + (size_t)tp * a.o_st + (size_t)j * a.o_sj);
      }
      Diffusion<M>::prep(W, tc.ev, d, xpre, upre, dtape_next);  // xpre = x_{t-1}, requested at the top of this iteration
    } else if (noisy) {
      if constexpr (PREFETCH) Diffusion<M>::vjp(W, tc.ev, d, x, u, lam, [&](int i) { return nzv[i]; }, gx, gu);
      else {
        const float* nzt = nz + (size_t)t * NX * G + g;
        Diffusion<M>::vjp(W, tc.ev, d, x, u, lam, [&](int i) { return nzt[(size_t)i * G]; }, gx, gu);
      }
    }
    // stage cost at t: trapezoid weight w_t
    float wt = a.sum_mode ? (t > 0 ? 1.0f : 0.0f) : (t > 0 ? 0.5f * (__ldg(a.dt + t - 1) + dt) : 0.5f * dt);
    if constexpr (M::Ev::PGRAD) {
      if (a.step_w && t > 0) wt = __ldg(a.step_w + (size_t)(t - 1) * G + g);
    }
    float ct = stage_cost<NX, NU, true, PLAINC>(cd, x, u, xr + (size_t)t * NX, true, wt, gx, gu);
    if (CONSTR) {
      float sl[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        sl[i] = 0.0f;
        if (cd.slack_col[i] >= 0 && cd.constr_mode == SDE4_CONSTR_WITHPROX)
          sl[i] = t > 0 ? __ldg(op + (size_t)(t - 1) * a.o_st + (size_t)(NU + cd.slack_col[i]) * a.o_sj) : x0s[i];
      }
      ct += constr_cost<NX, true>(cd, x, sl, wt, gx, gs_pend);  // slack of stage t lives in opt row t-1
    }
    cost = fmaf(wt, ct, cost);
    if (a.write_stage_cost && contrib) a.xbuf[((size_t)t * a.x_rows + NX) * G + g] = ct;
    // emit row t of the per-sample gradient
    {
      float* gp = a.gpart + (size_t)t * (NU + a.n_slack) * a.PW;
      if constexpr (L > 1 && NU <= L) {
        // gu is replicated over the sample's lanes: lane j of every group carries gu[j], the
        // groups of the warp are summed with log2(32/L) butterflies, lanes 0..NU-1 store
        float v = 0.0f;
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          if (tc.l == j) v = gu[j];
          gu[j] = 0.0f;
        }
        if (!active) v = 0.0f;
        // branch free: the butterflies always run, their sum is used only when the warp's samples share a problem
        float vs = v;
#pragma unroll
        for (int o = 16; o >= L; o >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, o);
        const bool wr = a.warp_reduce != 0;
        const bool st = active && (wr ? (int)(threadIdx.x & 31) < NU : tc.l < NU);
        if (st) gp[(size_t)tc.l * a.PW + pw] = wr ? vs : v;
      } else {
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          float v = contrib ? gu[j] : 0.0f;
          if (a.warp_reduce) v = warp_sum(v);
          if (writer) gp[(size_t)j * a.PW + pw] = v;
          gu[j] = 0.0f;
        }
      }
      if (CONSTR) {
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          if (cd.slack_col[i] >= 0 && cd.constr_mode == SDE4_CONSTR_WITHPROX) {
            float v = contrib ? gs[i] : 0.0f;
            if (a.warp_reduce) v = warp_sum(v);
            if (writer) gp[(size_t)(NU + cd.slack_col[i]) * a.PW + pw] = v;
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) lam[i] = gx[i];
    if constexpr (PIPE) dtape = dtape_next;
  }
  if constexpr (M::Ev::PGRAD && M::NPHYS > 0) {
    if (a.pd_phys && contrib) {  // one lane per sample (the accumulators are replicated over a sample's lanes)
#pragma unroll
      for (int i = 0; i < M::NPHYS; ++i) a.pd_phys[(size_t)i * a.pd_stride + g] = tc.ev.gS[i];
    }
  }
  emit_cost();
}

}  // namespace sde4

// kernels_ws.cuh -- warp-specialised variant of the fused rollout + cost + adjoint kernel (K2) for launches that are
// bound by the serial instruction chain of a step (one MPC problem with a few thousand particles: BASELINE configs[2]).
//
// Same mathematics as k_cost<M, GRAD = true> (kernels.cuh; sde4mbrl/nsde.py:1485-1527,1576-1586 under
// jax.value_and_grad, sde4mbrl/apg.py:82,224), different mapping: the two halves of a step that are independent given
// x_t run in DIFFERENT WARPS of one CTA and meet at one named barrier per step, exchanging a few dozen floats through
// shared memory (double buffered):
//
//   forward sweep    warp role A  dw_t (in-kernel threefry, one step ahead), density net -> sigma(x_t, u_t) o dw_t
//                    warp role B  physics + residual networks -> f(x_t, u_t); checkpoints x_{t+1}, 1/|q|
//                    both apply the Euler-Maruyama update + projection to their own copy of x (bitwise identical)
//   backward sweep   warp role P (= A's warp)  everything that does NOT depend on the cotangent, one step ahead of C:
//                                 density value + input gradient, forward passes of the residual nets (act' of the
//                                 hidden units), stage cost and its gradient
//                    warp role C (= B's warp)  the cotangent chain: projection VJP, physics adjoint, backward passes
//                                 of the residual nets from P's tape, diffusion adjoint, gradient-row emission
//
// ncu of the one-role kernel (profiles/r1_m_kcost_ncu.txt): 1024 warps on 592 schedulers, every warp issuing one
// instruction per ~3 cycles whether or not it shares its scheduler -- the launch time was the serial instruction count
// of ONE warp.  Splitting by role halves that count and doubles the resident warps without replicating any more scalar
// code than the lane split already does.
#pragma once
#include "kernels.cuh"

namespace sde4 {

template <class MA, class MB>
struct WsCfg {
  using A = typename MA::A;
  using NetA = typename A::NetA;
  using NetB = typename A::NetB;
  using DNet = typename A::DNet;
  static constexpr int LA = MA::Ev::L, LB = MB::Ev::L;
  static_assert(LA > 1 && LB > 1, "warp-specialised kernels are lane-split kernels");
  static constexpr int LMIN = LA < LB ? LA : LB;
  static constexpr int NPART = 32 / LMIN;  // particles per CTA
  static constexpr int WA = NPART * LA / 32, WB = NPART * LB / 32;
  static constexpr int THREADS = 32 * (WA + WB);
  static constexpr int NX = MA::NX, NU = MA::NU;
  static constexpr int DIN = DNet::IN > 0 ? DNet::IN : 0;
  // forward exchange (values x NPART): f[NX] | sigma o dw [NX]
  static constexpr int F_F = 0, F_S = NX, FWD_VALS = 2 * NX;
  // backward item (values x NPART)
  static constexpr int O_FB = 0, O_DENS = 3, O_J = 4, O_GXC = O_J + DIN, O_GUC = O_GXC + NX, O_WCT = O_GUC + NU,
                       O_UA1 = O_WCT + 1, O_UA2 = O_UA1 + NetA::H1P, O_UB1 = O_UA2 + NetA::H2P, O_UB2 = O_UB1 + NetB::H1P,
                       ITEM_VALS = O_UB2 + NetB::H2P;
  static constexpr int BUF_VALS = ITEM_VALS > FWD_VALS ? ITEM_VALS : FWD_VALS;
  static constexpr int PRE_VALS = NX + NU + 1;  // stage-H item: lam | gu | cost
  static constexpr int XCH_FLOATS = (2 * BUF_VALS + PRE_VALS) * NPART;
  static constexpr size_t smem_bytes() { return sizeof(float) * ((size_t)A::SMEM_FLOATS + XCH_FLOATS); }
};

// all threads of the CTA, one hardware barrier (the roles run different code, so this is spelled as a named
// barrier rather than __syncthreads())
template <int THREADS>
SDE4_DEV void ws_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
}

// ---- role A / P --------------------------------------------------------------------------------------------------
template <class MA, class MB, bool PLAINC>
SDE4_DEV void ws_role_ap(const float* W, float* xch, const sde4_model_desc& d, const sde4_cost_desc& cd,
                         const sde4_cost_desc& td, const CostK& a, int pl, int g, int p, int n, int pd, bool active) {
  using Cfg = WsCfg<MA, MB>;
  using A = typename MA::A;
  using NL = NoiseLanes<MA>;
  constexpr int NX = Cfg::NX, NU = Cfg::NU, NPART = Cfg::NPART, BV = Cfg::BUF_VALS * NPART;
  constexpr int LA = Cfg::LA;
  const unsigned G = (unsigned)a.G;
  const int H = a.H;
  Tctx<MA> tc;
  tc.l = threadIdx.x & (LA - 1);
  tc.ev.l = tc.l;
  const bool contrib = active && tc.l == 0;
  const float* op = a.opt + (size_t)p * a.o_sp;
  const float* xr = a.xref + (size_t)pd * a.xref_sp;
  const uint32_t nzmask = noise_mask<NX>(d);
  float x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = __ldg(a.y0 + (size_t)pd * NX + i);
  uint32_t b0, b1;
  {
    const uint32_t* k = a.key + (size_t)pd * a.key_sp;
    brownian_key(__ldg(k), __ldg(k + 1), (uint32_t)(n + a.n_off), (uint32_t)a.n_tot, a.rng_mode, b0, b1, a.rng_chain);
  }
  // ---------------- forward sweep: sigma(x_t, u_t) o dw_t ----------------
  float vn[NL::PER];
  NL::gen(b0, b1, a.rng_mode, nzmask, 0, H, tc.l, vn);
  float u[NU];
#pragma unroll
  for (int j = 0; j < NU; ++j) u[j] = 0.0f;
  for (int t = 0; t < H; ++t) {
    const float dt = __ldg(a.dt + t);
    float dw[NX], sig[NX];
    if constexpr (Diffusion<MA>::CDN) {
#pragma unroll
      for (int j = 0; j < NU; ++j) u[j] = __ldg(op + (size_t)t * a.o_st + (size_t)j * a.o_sj);
    }
    NL::distribute(vn, dw);
    NL::keep(vn, a.dwbuf + (size_t)t * NX * G + g, G, nzmask, tc.l, active);
    NL::gen(b0, b1, a.rng_mode, nzmask, t + 1 < H ? t + 1 : t, H, tc.l, vn);  // overlaps the density net below
    Diffusion<MA>::fwd(W, tc.ev, d, x, u, sig);
    float* buf = xch + (t & 1) * BV;
    float sdw[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      sdw[i] = sig[i] * dw[i];
      buf[(Cfg::F_S + i) * NPART + pl] = sdw[i];
    }
    ws_sync<Cfg::THREADS>();
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = fmaf(buf[(Cfg::F_F + i) * NPART + pl], dt, x[i]) + sdw[i];
    float aux;
    MA::project(x, aux);
  }
  ws_sync<Cfg::THREADS>();  // checkpoints (written by role B) are read back below

  // ---------------- backward sweep: cotangent-independent work, one step ahead of role C ----------------
  float* pre = xch + 2 * BV;
  {
    // stage H: x holds x_H, its control is u_{H-1} repeated (nsde.py:1509)
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = __ldg(op + (size_t)(H - 1) * a.o_st + (size_t)j * a.o_sj);
    float lamH[NX], guH[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) lamH[i] = 0.0f;
#pragma unroll
    for (int j = 0; j < NU; ++j) guH[j] = 0.0f;
    const float wH = 0.5f * __ldg(a.dt + H - 1);
    const float cH = stage_cost<NX, NU, true, PLAINC>(cd, x, u, xr + (size_t)H * NX, true, wH, lamH, guH);
    float cost = wH * cH;
    if (a.write_stage_cost && contrib) a.xbuf[((size_t)H * a.x_rows + NX) * G + g] = cH;
    if (a.has_terminal) {
      float gu_unused[NU];
      cost += stage_cost<NX, NU, true>(td, x, u, xr + (size_t)H * NX, false, 1.0f, lamH, gu_unused);
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) pre[i * NPART + pl] = lamH[i];
#pragma unroll
    for (int j = 0; j < NU; ++j) pre[(NX + j) * NPART + pl] = guH[j];
    pre[(NX + NU) * NPART + pl] = cost;
  }
  float xpre[NX];
  {
    const float* in = a.xbuf + (size_t)(H - 1) * a.x_rows * G + g;
#pragma unroll
    for (int i = 0; i < NX; ++i) xpre[i] = in[(size_t)i * G];
  }
  // item of step t (checkpoint x_t in xpre) -> buffer t & 1; the checkpoint of the step after is requested first
  auto produce = [&](int t) {
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = xpre[i];
    {
      const int tp = t > 0 ? t - 1 : 0;
      const float* in = a.xbuf + (size_t)tp * a.x_rows * G + g;
#pragma unroll
      for (int i = 0; i < NX; ++i) xpre[i] = in[(size_t)i * G];
    }
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = __ldg(op + (size_t)t * a.o_st + (size_t)j * a.o_sj);
    float* buf = xch + (t & 1) * BV;
    if constexpr (A::HAS_DENSITY) {
      typename Diffusion<MA>::Tape dtp;
      Diffusion<MA>::prep(W, tc.ev, d, x, u, dtp);
      buf[Cfg::O_DENS * NPART + pl] = dtp.dens;
#pragma unroll
      for (int k = 0; k < Cfg::DIN; ++k) buf[(Cfg::O_J + k) * NPART + pl] = dtp.J[k];
    }
    {
      typename MA::DriftTape tp;
      MA::drift_prep(W, tc.ev, d, x, tp);
#pragma unroll
      for (int k = 0; k < 3; ++k) buf[(Cfg::O_FB + k) * NPART + pl] = tp.Fb[k];
      // act' of the hidden units, by unit index (role C may split the layers over a different number of lanes)
#pragma unroll
      for (int m = 0; m < MA::TB1A; ++m) buf[(Cfg::O_UA1 + tc.l * MA::TB1A + m) * NPART + pl] = tp.dA1[m];
#pragma unroll
      for (int m = 0; m < MA::TB2A; ++m) buf[(Cfg::O_UA2 + tc.l * MA::TB2A + m) * NPART + pl] = tp.dA2[m];
#pragma unroll
      for (int m = 0; m < MA::TB1B; ++m) buf[(Cfg::O_UB1 + tc.l * MA::TB1B + m) * NPART + pl] = tp.dB1[m];
#pragma unroll
      for (int m = 0; m < MA::TB2B; ++m) buf[(Cfg::O_UB2 + tc.l * MA::TB2B + m) * NPART + pl] = tp.dB2[m];
    }
    {
      // stage cost at t with its trapezoid weight (nsde.py:1515), value and gradient
      const float dt = __ldg(a.dt + t);
      const float wt = t > 0 ? 0.5f * (__ldg(a.dt + t - 1) + dt) : 0.5f * dt;
      float gxc[NX], guc[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) gxc[i] = 0.0f;
#pragma unroll
      for (int j = 0; j < NU; ++j) guc[j] = 0.0f;
      const float ct = stage_cost<NX, NU, true, PLAINC>(cd, x, u, xr + (size_t)t * NX, true, wt, gxc, guc);
      if (a.write_stage_cost && contrib) a.xbuf[((size_t)t * a.x_rows + NX) * G + g] = ct;
#pragma unroll
      for (int i = 0; i < NX; ++i) buf[(Cfg::O_GXC + i) * NPART + pl] = gxc[i];
#pragma unroll
      for (int j = 0; j < NU; ++j) buf[(Cfg::O_GUC + j) * NPART + pl] = guc[j];
      buf[Cfg::O_WCT * NPART + pl] = wt * ct;
    }
  };
  for (int t = H; t >= 1; --t) {
    produce(t - 1);  // role C consumes item t meanwhile (nothing in the first round)
    ws_sync<Cfg::THREADS>();
  }
  ws_sync<Cfg::THREADS>();  // role C's last iteration
}

// ---- role B / C --------------------------------------------------------------------------------------------------
template <class MA, class MB, bool PLAINC>
SDE4_DEV void ws_role_bc(const float* W, float* xch, const sde4_model_desc& d, const sde4_cost_desc& cd,
                         const CostK& a, int pl, int g, int p, int pd, bool active) {
  using Cfg = WsCfg<MA, MB>;
  using A = typename MB::A;
  constexpr int NX = Cfg::NX, NU = Cfg::NU, NPART = Cfg::NPART, BV = Cfg::BUF_VALS * NPART;
  constexpr int L = Cfg::LB, SPW = 32 / L;
  const unsigned G = (unsigned)a.G;
  const int H = a.H;
  Tctx<MB> tc;
  tc.l = threadIdx.x & (L - 1);
  tc.ev.l = tc.l;
  const bool contrib = active && tc.l == 0;
  const float* op = a.opt + (size_t)p * a.o_sp;
  const uint32_t nzmask = noise_mask<NX>(d);
  float x[NX], u[NU];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = __ldg(a.y0 + (size_t)pd * NX + i);
  store_state<MB>(a.xbuf + g, G, x, tc.l, active);
  // ---------------- forward sweep: f(x_t, u_t), checkpoints ----------------
  for (int t = 0; t < H; ++t) {
    const float dt = __ldg(a.dt + t);
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = __ldg(op + (size_t)t * a.o_st + (size_t)j * a.o_sj);
    float f[NX];
    MB::drift(W, tc.ev, d, x, u, f);
    float* buf = xch + (t & 1) * BV;
#pragma unroll
    for (int i = 0; i < NX; ++i) buf[(Cfg::F_F + i) * NPART + pl] = f[i];
    ws_sync<Cfg::THREADS>();
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = fmaf(f[i], dt, x[i]) + buf[(Cfg::F_S + i) * NPART + pl];
    float aux = 0.0f;
    MB::project(x, aux);
    if (MB::NAUX && contrib) a.auxbuf[(size_t)t * G + g] = aux;
    store_state<MB>(a.xbuf + (size_t)(t + 1) * a.x_rows * G + g, G, x, tc.l, active);
  }
  ws_sync<Cfg::THREADS>();  // sibling lanes' checkpoints and role A's noise are read back below

  // ---------------- backward sweep: the cotangent chain ----------------
  const int pw = a.warp_reduce ? (g / SPW) : g;
  const bool writer = a.warp_reduce ? (active && (threadIdx.x & 31) == 0) : contrib;
  float xpre[NX], auxpre = 0.0f;
  {
    const float* in = a.xbuf + (size_t)(H - 1) * a.x_rows * G + g;
#pragma unroll
    for (int i = 0; i < NX; ++i) xpre[i] = in[(size_t)i * G];
    if (MB::NAUX) auxpre = a.auxbuf[(size_t)(H - 1) * G + g];
  }
  ws_sync<Cfg::THREADS>();  // stage-H item and item H-1 are there
  float lam[NX], gu[NU], cost;
  {
    const float* pre = xch + 2 * BV;
#pragma unroll
    for (int i = 0; i < NX; ++i) lam[i] = pre[i * NPART + pl];
#pragma unroll
    for (int j = 0; j < NU; ++j) gu[j] = pre[(NX + j) * NPART + pl];
    cost = pre[(NX + NU) * NPART + pl];
  }
  const float* nz = a.dwbuf;
  for (int t = H - 1; t >= 0; --t) {
    const float dt = __ldg(a.dt + t);
    // noise of this step, requested before the drift adjoint that precedes its use
    float nzv[NX];
    {
      const float* nzt = nz + (size_t)t * NX * G + g;
#pragma unroll
      for (int i = 0; i < NX; ++i)
        if ((A::NOUT_MASK >> i) & 1u) nzv[i] = ((nzmask >> i) & 1u) ? nzt[(size_t)i * G] : 0.0f;
    }
    MB::project_vjp(x, auxpre, lam);  // x still holds the projected x_{t+1}
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = xpre[i];
    {  // unconditional (t = 0 re-reads row 0): no branch in the loop body
      const int tp = t > 0 ? t - 1 : 0;
      const float* in = a.xbuf + (size_t)tp * a.x_rows * G + g;
#pragma unroll
      for (int i = 0; i < NX; ++i) xpre[i] = in[(size_t)i * G];
      if (MB::NAUX) auxpre = a.auxbuf[(size_t)tp * G + g];
    }
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = __ldg(op + (size_t)t * a.o_st + (size_t)j * a.o_sj);
    const float* buf = xch + (t & 1) * BV;
    float gx[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) gx[i] = lam[i];
    {
      typename MB::DriftTape tp;
#pragma unroll
      for (int k = 0; k < 3; ++k) tp.Fb[k] = buf[(Cfg::O_FB + k) * NPART + pl];
#pragma unroll
      for (int m = 0; m < MB::TB1A; ++m) tp.dA1[m] = buf[(Cfg::O_UA1 + tc.l * MB::TB1A + m) * NPART + pl];
#pragma unroll
      for (int m = 0; m < MB::TB2A; ++m) tp.dA2[m] = buf[(Cfg::O_UA2 + tc.l * MB::TB2A + m) * NPART + pl];
#pragma unroll
      for (int m = 0; m < MB::TB1B; ++m) tp.dB1[m] = buf[(Cfg::O_UB1 + tc.l * MB::TB1B + m) * NPART + pl];
#pragma unroll
      for (int m = 0; m < MB::TB2B; ++m) tp.dB2[m] = buf[(Cfg::O_UB2 + tc.l * MB::TB2B + m) * NPART + pl];
      const Scaled<NX> vf{lam, dt};
      MB::drift_apply(W, tc.ev, d, x, u, vf, tp, gx, gu);
    }
    if constexpr (A::HAS_DENSITY) {
      typename Diffusion<MB>::Tape dtp;
      dtp.dens = buf[Cfg::O_DENS * NPART + pl];
#pragma unroll
      for (int k = 0; k < Cfg::DIN; ++k) dtp.J[k] = buf[(Cfg::O_J + k) * NPART + pl];
      Diffusion<MB>::apply(W, d, x, lam, [&](int i) { return nzv[i]; }, dtp, gx, gu);
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) gx[i] += buf[(Cfg::O_GXC + i) * NPART + pl];
#pragma unroll
    for (int j = 0; j < NU; ++j) gu[j] += buf[(Cfg::O_GUC + j) * NPART + pl];
    cost += buf[Cfg::O_WCT * NPART + pl];
    // emit row t of the per-sample gradient (same partial layout as k_cost)
    {
      float* gp = a.gpart + (size_t)t * NU * a.PW;
      if constexpr (NU <= L) {
        float v = 0.0f;
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          if (tc.l == j) v = gu[j];
          gu[j] = 0.0f;
        }
        if (!active) v = 0.0f;
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
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) lam[i] = gx[i];
    ws_sync<Cfg::THREADS>();
  }
  {
    float cv = contrib ? cost : 0.0f;
    if (a.warp_reduce) cv = warp_sum(cv);
    if (writer) a.cpart[pw] = cv;
  }
}

// One CTA = NPART particles: WA warps of role A / P (LA lanes per particle) followed by WB warps of role B / C
// (LB lanes per particle).  Requirements checked by the launcher: value + gradient, in-kernel threefry noise,
// no state constraints, Euler-Maruyama.
#ifndef SDE4_WS_MIN_BLOCKS
#define SDE4_WS_MIN_BLOCKS 7
#endif
template <class MA, class MB, bool PLAINC>
__global__ void __launch_bounds__((WsCfg<MA, MB>::THREADS), SDE4_WS_MIN_BLOCKS)
    k_cost_ws(const __grid_constant__ sde4_model_desc d, const __grid_constant__ sde4_cost_desc cd,
              const __grid_constant__ sde4_cost_desc td, const __grid_constant__ CostK a) {
  using Cfg = WsCfg<MA, MB>;
  using A = typename MA::A;
  extern __shared__ __align__(16) float W[];
  __shared__ __align__(8) uint64_t mbar;
  stage_params_bulk(W, a.params, A::PARAM_TOTAL * sizeof(float), &mbar);
  Diffusion<MA>::derive(W);
  __syncthreads();
  float* xch = W + A::SMEM_FLOATS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool role_a = warp < Cfg::WA;
  const int pl = role_a ? (warp * 32 + lane) / Cfg::LA : ((warp - Cfg::WA) * 32 + lane) / Cfg::LB;
  const int g0 = blockIdx.x * Cfg::NPART + pl;
  const bool active = g0 < a.G;
  const int g = active ? g0 : a.G - 1;  // inactive lanes shadow a valid sample, contribute nothing
  const int p = g / a.N, n = g - p * a.N;
  const int pd = p % a.data_mod;
  if (a.mask) {  // CTA-uniform exit (the roles meet at CTA-wide barriers)
    const int on = __ldg(a.mask + p) != 0;
    if (!__syncthreads_or(on)) return;
  }
  if (role_a) ws_role_ap<MA, MB, PLAINC>(W, xch, d, cd, td, a, pl, g, p, n, pd, active);
  else ws_role_bc<MA, MB, PLAINC>(W, xch, d, cd, a, pl, g, p, pd, active);
}

}  // namespace sde4

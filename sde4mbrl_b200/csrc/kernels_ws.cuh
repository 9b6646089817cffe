// kernels_ws.cuh -- warp-specialised variant of the fused rollout + cost + adjoint kernel (K2) of the multirotor SDE
// for launches that are bound by the serial instruction chain of a step (one MPC problem with a few thousand particles:
// BASELINE configs[2]).
//
// Same mathematics as k_cost<RotorModel, GRAD = true> (kernels.cuh; sde4mbrl/nsde.py:1485-1527,1576-1586 under
// jax.value_and_grad, sde4mbrl/apg.py:82,224; drift sde_rotor_model.py:427-458), different mapping.  In the lane-split
// kernels every one of the 8 lanes of a particle repeats the scalar work (quaternion physics, projection, stage cost,
// address arithmetic): 40 % of the 289 M warp instructions of a C3 launch, eight times over
// (profiles/r1_m_kcost_ncu.txt).  Here a CTA owns NPART particles and its warps have two ROLES:
//
//   PHY   one warp, ONE LANE PER PARTICLE: the state x, the Euler-Maruyama update, projection, checkpoints, stage cost,
//         and in the backward sweep the cotangent lambda and the whole physics adjoint -- nothing is replicated;
//   DENS  NPART/4 warps, 8 lanes per particle (EvL<8>): the diffusion density network, the in-kernel threefry noise and
//         sigma o dw; in the backward sweep the density value + input gradient (cotangent independent, so it runs one
//         step ahead of the cotangent chain and never sits on it);
//   RES   NPART/8 warps, 4 lanes per particle: the residual force and moment networks (forward; in the backward sweep
//         their forward passes one step ahead, then the backward passes for the cotangents PHY hands over).
//
// PHY and the network roles alternate within a step -- v_b -> networks -> forces is a true dependency -- and meet at
// two CTA-wide barriers per step, exchanging 12 + 19 floats (forward) / 6 + 19 floats (backward) per particle through
// shared memory.
//
// (A first attempt with two symmetric 8-lane roles was slower than the one-role kernel: profiles/r2_a_ws_symmetric_ncu.txt.)
#pragma once
#include "kernels.cuh"

namespace sde4 {

// ----------------------------------------------------------------------------------------------------------------
// Multirotor physics split around the network evaluations, one thread per particle.
// Transcribed from RotorModel::drift / drift_vjp (models.cuh), which stay the reference for the other kernels.
// ----------------------------------------------------------------------------------------------------------------
template <class A_>
struct RotorSplit {
  using A = A_;
  using NetA = typename A::NetA;
  using NetB = typename A::NetB;
  static constexpr int NX = 13, NU = A::NU;
  static constexpr bool DRAG = (A::FLAGS & SDE4_ROTOR_AERO_DRAG) != 0;
  static constexpr bool HAS_F = NetA::PRESENT, HAS_M = NetB::PRESENT;
  static constexpr int NIA = HAS_F ? NetA::IN : 1, NIB = HAS_M ? NetB::IN : 1;

  SDE4_DEV static float thrust(const float* __restrict__ S, float u) { return fmaf(fmaf(fmaf(S[7], u, S[8]), u, S[9]), u, S[10]); }
  SDE4_DEV static float thrust_du(const float* __restrict__ S, float u) { return fmaf(fmaf(3.0f * S[7], u, 2.0f * S[8]), u, S[9]); }

  // v_b = q^-1 (x) (v (x) q) and the scaled network inputs (sde_rotor_model.py:311-315, :439)
  SDE4_DEV static void features(const float (&x)[NX], const sde4_model_desc& d, Quat& t1, float (&vb)[3], float (&in6)[6]) {
    const Quat q{x[6], x[7], x[8], x[9]};
    t1 = qmul_vq(x[3], x[4], x[5], q);
    const Quat r1 = qmul(qconj(q), t1);
    vb[0] = r1.x;
    vb[1] = r1.y;
    vb[2] = r1.z;
    in6[0] = vb[0] * d.inv_state_scaling[3];
    in6[1] = vb[1] * d.inv_state_scaling[4];
    in6[2] = vb[2] * d.inv_state_scaling[5];
    in6[3] = x[10] * d.inv_state_scaling[10];
    in6[4] = x[11] * d.inv_state_scaling[11];
    in6[5] = x[12] * d.inv_state_scaling[12];
  }

  // total body force [0,0,T] + Fres (+ aero drag) from the residual-force network output
  SDE4_DEV static void body_force(const float* __restrict__ S, const float (&u)[NU], const float (&vb)[3],
                                  const float (&Fn)[3], float (&Fb)[3]) {
    Fb[0] = Fn[0];
    Fb[1] = Fn[1];
    Fb[2] = Fn[2];
    if constexpr (DRAG) {  // :264-277
      Fb[0] += -S[11] * vb[0];
      Fb[1] += -S[12] * vb[1];
      Fb[2] += -S[13] * vb[2] + S[14] * (vb[0] * vb[0] + vb[1] * vb[1]);
    }
    float T = 0.0f;
#pragma unroll
    for (int i = 0; i < NU; ++i) T = fmaf(Mixer<NU>::m(3, i), thrust(S, u[i]), T);
    Fb[2] += T;
  }

  // f(x, u) given the network outputs Fn (residual force), Mn (residual moment)
  SDE4_DEV static void drift_post(const float* __restrict__ S, const sde4_model_desc& d, const float (&x)[NX],
                                  const float (&u)[NU], const float (&vb)[3], const float (&Fn)[3], const float (&Mn)[3],
                                  float (&f)[NX]) {
    const Quat q{x[6], x[7], x[8], x[9]};
    float Fb[3];
    body_force(S, u, vb, Fn, Fb);
    float Mm[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int i = 0; i < NU; ++i) {  // motor_models.py:74-95
      const float th = thrust(S, u[i]);
      Mm[0] = fmaf(Mixer<NU>::m(0, i) * S[4], th, Mm[0]);
      Mm[1] = fmaf(Mixer<NU>::m(1, i) * S[5], th, Mm[1]);
      Mm[2] = fmaf(Mixer<NU>::m(2, i) * S[6], th, Mm[2]);
    }
    const Quat t2 = qmul_qv(q, Fb[0], Fb[1], Fb[2]);
    const Quat r2 = qmul(t2, qconj(q));
    const float inv_m = fast_rcp(S[0]), g = d.cfloat[0];
    f[0] = x[3];
    f[1] = x[4];
    f[2] = x[5];
    f[3] = r2.x * inv_m;
    f[4] = r2.y * inv_m;
    f[5] = fmaf(r2.z, inv_m, -g);
    const float I0 = S[1], I1 = S[2], I2 = S[3];
    const float w0 = x[10], w1 = x[11], w2 = x[12];
    const float Iw0 = I0 * w0, Iw1 = I1 * w1, Iw2 = I2 * w2;
    f[10] = (Mm[0] + Mn[0] - (w1 * Iw2 - w2 * Iw1)) * fast_rcp(I0);
    f[11] = (Mm[1] + Mn[1] - (w2 * Iw0 - w0 * Iw2)) * fast_rcp(I1);
    f[12] = (Mm[2] + Mn[2] - (w0 * Iw1 - w1 * Iw0)) * fast_rcp(I2);
    const Quat qd = qmul_qv(q, w0, w1, w2);
    f[6] = 0.5f * qd.w;
    f[7] = 0.5f * qd.x;
    f[8] = 0.5f * qd.y;
    f[9] = 0.5f * qd.z;
  }

  // Adjoint, first half: the cotangents of the two network outputs (vf = lambda * dt is the cotangent of f)
  struct AdjTmp {
    Quat t2bar;
    float R2[3], gM[3], gFb[3];
  };
  SDE4_DEV static void vjp_pre(const float* __restrict__ S, const float (&x)[NX], const Scaled<NX>& vf, AdjTmp& a) {
    const Quat q{x[6], x[7], x[8], x[9]};
    a.gM[0] = vf[10] * fast_rcp(S[1]);
    a.gM[1] = vf[11] * fast_rcp(S[2]);
    a.gM[2] = vf[12] * fast_rcp(S[3]);
    const float inv_m = fast_rcp(S[0]);
    a.R2[0] = vf[3] * inv_m;
    a.R2[1] = vf[4] * inv_m;
    a.R2[2] = vf[5] * inv_m;
    a.t2bar = qmul_vq(a.R2[0], a.R2[1], a.R2[2], q);  // r2 = t2 (x) conj(q)  =>  t2bar = R2 (x) q
    const Quat gFq = qmul(qconj(q), a.t2bar);          // t2 = q (x) Fq        =>  Fqbar = conj(q) (x) t2bar
    a.gFb[0] = gFq.x;
    a.gFb[1] = gFq.y;
    a.gFb[2] = gFq.z;
  }
  // second half: everything else, given the input cotangents giA / giB the networks returned for gFb / gM
  SDE4_DEV static void vjp_post(const float* __restrict__ S, const sde4_model_desc& d, const float (&x)[NX],
                                const float (&u)[NU], const Scaled<NX>& vf, const Quat& t1, const float (&vb)[3],
                                const float (&Fn)[3], const AdjTmp& a, const float (&giA)[NIA], const float (&giB)[NIB],
                                float (&gx)[NX], float (&gu)[NU]) {
    const Quat q{x[6], x[7], x[8], x[9]};
    const Quat qc = qconj(q);
    const float I0 = S[1], I1 = S[2], I2 = S[3];
    const float w0 = x[10], w1 = x[11], w2 = x[12];
    Quat gq{0.0f, 0.0f, 0.0f, 0.0f};
    float gw[3] = {0.0f, 0.0f, 0.0f};
    gx[3] += vf[0];
    gx[4] += vf[1];
    gx[5] += vf[2];
    {  // alpha = (Mm + Mres - w x (I w)) / I
      const float b0 = I0 * w0, b1 = I1 * w1, b2 = I2 * w2;
      const float c0 = -a.gM[0], c1 = -a.gM[1], c2 = -a.gM[2];
      gw[0] += b1 * c2 - b2 * c1;
      gw[1] += b2 * c0 - b0 * c2;
      gw[2] += b0 * c1 - b1 * c0;
      const float bb0 = c1 * w2 - c2 * w1, bb1 = c2 * w0 - c0 * w2, bb2 = c0 * w1 - c1 * w0;
      gw[0] = fmaf(I0, bb0, gw[0]);
      gw[1] = fmaf(I1, bb1, gw[1]);
      gw[2] = fmaf(I2, bb2, gw[2]);
    }
    {  // qdot = 0.5 q (x) [0, w]
      const Quat gqd{0.5f * vf[6], 0.5f * vf[7], 0.5f * vf[8], 0.5f * vf[9]};
      qacc(gq, qmul_qv(gqd, -w0, -w1, -w2));
      const Quat gOm = qmul(qc, gqd);
      gw[0] += gOm.x;
      gw[1] += gOm.y;
      gw[2] += gOm.z;
    }
    float gvb[3] = {0.0f, 0.0f, 0.0f};
    float gin6[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if constexpr (HAS_F) {
#pragma unroll
      for (int k = 0; k < NetA::IN; ++k) gin6[nth_bit(A::A_MASK, k)] += giA[k];
    }
    if constexpr (DRAG) {
      gvb[0] += -S[11] * a.gFb[0] + 2.0f * S[14] * vb[0] * a.gFb[2];
      gvb[1] += -S[12] * a.gFb[1] + 2.0f * S[14] * vb[1] * a.gFb[2];
      gvb[2] += -S[13] * a.gFb[2];
    }
    {
      float Fb[3];
      body_force(S, u, vb, Fn, Fb);
      const Quat t2 = qmul_qv(q, Fb[0], Fb[1], Fb[2]);
      qacc(gq, qconj(qmul_qv(qconj(t2), a.R2[0], a.R2[1], a.R2[2])));  // qcbar = conj(t2) (x) R2
      qacc(gq, qmul_qv(a.t2bar, -Fb[0], -Fb[1], -Fb[2]));               // qbar += t2bar (x) conj(Fq)
    }
    {  // motors
      const float gT = a.gFb[2];
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        float gth = Mixer<NU>::m(3, i) * gT;
        gth = fmaf(Mixer<NU>::m(0, i) * S[4], a.gM[0], gth);
        gth = fmaf(Mixer<NU>::m(1, i) * S[5], a.gM[1], gth);
        gth = fmaf(Mixer<NU>::m(2, i) * S[6], a.gM[2], gth);
        gu[i] = fmaf(gth, thrust_du(S, u[i]), gu[i]);
      }
    }
    if constexpr (HAS_M) {
#pragma unroll
      for (int k = 0; k < NetB::IN; ++k) gin6[nth_bit(A::B_MASK, k)] += giB[k];
    }
    gvb[0] = fmaf(gin6[0], d.inv_state_scaling[3], gvb[0]);
    gvb[1] = fmaf(gin6[1], d.inv_state_scaling[4], gvb[1]);
    gvb[2] = fmaf(gin6[2], d.inv_state_scaling[5], gvb[2]);
    gw[0] = fmaf(gin6[3], d.inv_state_scaling[10], gw[0]);
    gw[1] = fmaf(gin6[4], d.inv_state_scaling[11], gw[1]);
    gw[2] = fmaf(gin6[5], d.inv_state_scaling[12], gw[2]);
    {  // v_b = vec(qc (x) t1), t1 = V (x) q
      qacc(gq, qconj(qmul_vq(gvb[0], gvb[1], gvb[2], qconj(t1))));
      const Quat t1bar = qmul_qv(q, gvb[0], gvb[1], gvb[2]);
      const Quat gV = qmul(t1bar, qc);
      gx[3] += gV.x;
      gx[4] += gV.y;
      gx[5] += gV.z;
      qacc(gq, qmul_vq(-x[3], -x[4], -x[5], t1bar));
    }
    gx[6] += gq.w;
    gx[7] += gq.x;
    gx[8] += gq.y;
    gx[9] += gq.z;
    gx[10] += gw[0];
    gx[11] += gw[1];
    gx[12] += gw[2];
  }
};

// ----------------------------------------------------------------------------------------------------------------
template <class A_, int NPART_>
struct WsCfg {
  using A = A_;
  using M8 = RotorModel<A, EvL<8>>;  // evaluator, diffusion constants and projection of the lane-split model
  using RS = RotorSplit<A>;
  using DNet = typename A::DNet;
  using NetA = typename A::NetA;
  using NetB = typename A::NetB;
  static_assert(A::MODEL == SDE4_MODEL_ROTOR && A::HAS_DENSITY && !Diffusion<M8>::CDN, "multirotor SDE with a density net");
  static constexpr int NPART = NPART_;  // particles per CTA: 8, 16 or 32
  static_assert(NPART == 8 || NPART == 16 || NPART == 32, "particles per CTA");
  static constexpr int LN = 8;                       // lanes per particle of the DENS role (32-wide layers: 4 units per lane)
  static constexpr int LR = 4;                       // lanes per particle of the RES role (8 / 16-wide layers: 2 / 4 units per lane)
  using MR = RotorModel<A, EvL<LR>>;
  static constexpr int WN = NPART * LN / 32;         // DENS warps
  static constexpr int WR = NPART * LR / 32;         // RES warps
  static constexpr int THREADS = 32 * (1 + WN + WR); // warp 0: PHY, then the DENS warps, then the RES warps
  static constexpr int NX = 13, NU = A::NU, NZ = DNet::IN, NO = A::NO;
  static constexpr int NIA = RS::NIA, NIB = RS::NIB;
  // exchange buffer, [value][NPART]
  static constexpr int X_IN6 = 0, X_Z = 6, X_FN = X_Z + NZ, X_MN = X_FN + 3, X_SDW = X_MN + 3, FWD_VALS = X_SDW + NX;
  static constexpr int B_GF = 0, B_GM = 3, B_GIA = 6, B_GIB = B_GIA + NIA, B_FN = B_GIB + NIB, B_DENS = B_FN + 3,
                       B_J = B_DENS + 1, BWD_VALS = B_J + NZ;
  static constexpr int B_GU = BWD_VALS;  // control-gradient row of the step PHY has just finished (summed by the RES role)
  static constexpr int BWD_ALL = B_GU + NU;
  static constexpr int VALS = FWD_VALS > BWD_ALL ? FWD_VALS : BWD_ALL;
  static constexpr int N_PRIOR = 16;  // noise prior staged in shared memory (indexed by a run-time state index)
  static constexpr int XCH_FLOATS = VALS * NPART + N_PRIOR;
  static constexpr size_t smem_bytes() { return sizeof(float) * ((size_t)A::SMEM_FLOATS + XCH_FLOATS); }
};

template <int THREADS>
SDE4_DEV void ws_sync() {  // CTA-wide; spelled as a named barrier because the roles run different code
  asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
}

// ---- PHY role: one lane per particle ---------------------------------------------------------------------------------
template <class Cfg, bool PLAINC>
SDE4_DEV void ws_role_phy(const float* W, float* xch, const sde4_model_desc& d, const sde4_cost_desc& cd,
                          const sde4_cost_desc& td, const CostK& a, float* ftape, int pl, int g, int p, int pd, bool active) {
  using A = typename Cfg::A;
  using RS = typename Cfg::RS;
  using M8 = typename Cfg::M8;
  constexpr int NX = Cfg::NX, NU = Cfg::NU, NPART = Cfg::NPART, NZ = Cfg::NZ;
  const float* __restrict__ S = W + A::OFF_SCALARS;
  const unsigned G = (unsigned)a.G;
  const int H = a.H;
  const float* op = a.opt + (size_t)p * a.o_sp;
  const float* xr = a.xref + (size_t)pd * a.xref_sp;
  const uint32_t nzmask = noise_mask<NX>(d);
  float x[NX], u[NU];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = __ldg(a.y0 + (size_t)pd * NX + i);
  if (active) {
#pragma unroll
    for (int i = 0; i < NX; ++i) a.xbuf[(size_t)i * G + g] = x[i];
  }
  // ---------------- forward sweep ----------------
  for (int t = 0; t < H; ++t) {
    const float dt = __ldg(a.dt + t);
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = __ldg(op + (size_t)t * a.o_st + (size_t)j * a.o_sj);
    Quat t1;
    float vb[3], in6[6];
    RS::features(x, d, t1, vb, in6);
#pragma unroll
    for (int k = 0; k < 6; ++k) xch[(Cfg::X_IN6 + k) * NPART + pl] = in6[k];
#pragma unroll
    for (int k = 0; k < NZ; ++k) xch[(Cfg::X_Z + k) * NPART + pl] = x[nth_bit(A::NIN_MASK, k)] * d.inv_red_scale;
    ws_sync<Cfg::THREADS>();  // B1: the networks may start
    if (active) {             // network inputs of the residual nets, kept for the backward sweep
      float* ft = ftape + (size_t)t * 6 * G + g;
#pragma unroll
      for (int k = 0; k < 6; ++k) ft[(size_t)k * G] = in6[k];
    }
    ws_sync<Cfg::THREADS>();  // B2: network outputs and sigma o dw are there
    float Fn[3], Mn[3], f[NX];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      Fn[k] = xch[(Cfg::X_FN + k) * NPART + pl];
      Mn[k] = xch[(Cfg::X_MN + k) * NPART + pl];
    }
    RS::drift_post(S, d, x, u, vb, Fn, Mn, f);
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = fmaf(f[i], dt, x[i]) + xch[(Cfg::X_SDW + i) * NPART + pl];
    float aux;
    M8::project(x, aux);
    if (active) {
      a.auxbuf[(size_t)t * G + g] = aux;
      float* out = a.xbuf + (size_t)(t + 1) * a.x_rows * G + g;
#pragma unroll
      for (int i = 0; i < NX; ++i) out[(size_t)i * G] = x[i];
    }
  }

  // ---------------- backward sweep ----------------
  float lam[NX], gu[NU], cost;
#pragma unroll
  for (int i = 0; i < NX; ++i) lam[i] = 0.0f;
#pragma unroll
  for (int j = 0; j < NU; ++j) gu[j] = 0.0f;  // row H-1 also receives the u-gradient of stage H
  {
    // stage H: x holds x_H, u holds u_{H-1} (the last control repeated, nsde.py:1509)
    const float wH = 0.5f * __ldg(a.dt + H - 1);
    const float cH = stage_cost<NX, NU, true, PLAINC>(cd, x, u, xr + (size_t)H * NX, true, wH, lam, gu);
    cost = wH * cH;
    if (a.write_stage_cost && active) a.xbuf[((size_t)H * a.x_rows + NX) * G + g] = cH;
    if (a.has_terminal) {
      float gu_unused[NU];
      cost += stage_cost<NX, NU, true>(td, x, u, xr + (size_t)H * NX, false, 1.0f, lam, gu_unused);
    }
  }
  constexpr int SPW = NPART;  // samples per partial when the CTA's particles share a problem
  const int pw = a.warp_reduce ? (g / SPW) : g;
  const bool writer = a.warp_reduce ? (active && (threadIdx.x & 31) == 0) : active;
  float xpre[NX], auxpre;
  {
    const float* in = a.xbuf + (size_t)(H - 1) * a.x_rows * G + g;
#pragma unroll
    for (int i = 0; i < NX; ++i) xpre[i] = in[(size_t)i * G];
    auxpre = a.auxbuf[(size_t)(H - 1) * G + g];
  }
  for (int t = H - 1; t >= 0; --t) {
    const float dt = __ldg(a.dt + t);
    float nzv[NX];
    {
      const float* nzt = a.dwbuf + (size_t)t * NX * G + g;
#pragma unroll
      for (int i = 0; i < NX; ++i)
        if ((A::NOUT_MASK >> i) & 1u) nzv[i] = ((nzmask >> i) & 1u) ? nzt[(size_t)i * G] : 0.0f;
    }
    M8::project_vjp(x, auxpre, lam);  // x still holds the projected x_{t+1}
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = xpre[i];
    {  // checkpoint of the next iteration, one step ahead (t = 0 re-reads row 0)
      const int tp = t > 0 ? t - 1 : 0;
      const float* in = a.xbuf + (size_t)tp * a.x_rows * G + g;
#pragma unroll
      for (int i = 0; i < NX; ++i) xpre[i] = in[(size_t)i * G];
      auxpre = a.auxbuf[(size_t)tp * G + g];
    }
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = __ldg(op + (size_t)t * a.o_st + (size_t)j * a.o_sj);
    const Scaled<NX> vf{lam, dt};
    typename RS::AdjTmp at;
    RS::vjp_pre(S, x, vf, at);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      xch[(Cfg::B_GF + k) * NPART + pl] = at.gFb[k];
      xch[(Cfg::B_GM + k) * NPART + pl] = at.gM[k];
    }
    ws_sync<Cfg::THREADS>();  // B3: the network backward passes may start
    // meanwhile: what does not need them -- features of x_t and the stage cost with its gradient
    Quat t1;
    float vb[3], in6[6];
    RS::features(x, d, t1, vb, in6);
    float gx[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) gx[i] = lam[i];
    {
      const float wt = t > 0 ? 0.5f * (__ldg(a.dt + t - 1) + dt) : 0.5f * dt;  // trapezoid weight (nsde.py:1515)
      const float ct = stage_cost<NX, NU, true, PLAINC>(cd, x, u, xr + (size_t)t * NX, true, wt, gx, gu);
      cost = fmaf(wt, ct, cost);
      if (a.write_stage_cost && active) a.xbuf[((size_t)t * a.x_rows + NX) * G + g] = ct;
    }
    ws_sync<Cfg::THREADS>();  // B4: input cotangents, residual force, density value + gradient are there
    float giA[Cfg::NIA], giB[Cfg::NIB], Fn[3];
#pragma unroll
    for (int k = 0; k < Cfg::NIA; ++k) giA[k] = xch[(Cfg::B_GIA + k) * NPART + pl];
#pragma unroll
    for (int k = 0; k < Cfg::NIB; ++k) giB[k] = xch[(Cfg::B_GIB + k) * NPART + pl];
#pragma unroll
    for (int k = 0; k < 3; ++k) Fn[k] = xch[(Cfg::B_FN + k) * NPART + pl];
    typename Diffusion<M8>::Tape dtp;
    dtp.dens = xch[Cfg::B_DENS * NPART + pl];
#pragma unroll
    for (int k = 0; k < NZ; ++k) dtp.J[k] = xch[(Cfg::B_J + k) * NPART + pl];
    RS::vjp_post(S, d, x, u, vf, t1, vb, Fn, at, giA, giB, gx, gu);
    Diffusion<M8>::apply(W, d, x, lam, [&](int i) { return nzv[i]; }, dtp, gx, gu);
    // row t of the gradient.  When the CTA's particles share a problem the row is summed over them -- by the RES
    // role, which has the time (30 dependent shuffle / add pairs here were half of this role's backward sweep);
    // otherwise one partial per particle, stored directly.
    if (a.warp_reduce) {
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        xch[(Cfg::B_GU + j) * NPART + pl] = active ? gu[j] : 0.0f;
        gu[j] = 0.0f;
      }
    } else {
      float* gp = a.gpart + (size_t)t * NU * a.PW;
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        if (active) gp[(size_t)j * a.PW + g] = gu[j];
        gu[j] = 0.0f;
      }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) lam[i] = gx[i];
  }
  ws_sync<Cfg::THREADS>();  // row 0 is there
  {
    float cv = active ? cost : 0.0f;
    if (a.warp_reduce) cv = warp_sum(cv);
    if (writer) a.cpart[pw] = cv;
  }
}

// ---- DENS role: 8 lanes per particle; density network, noise, sigma o dw -------------------------------------------------
template <class Cfg>
SDE4_DEV void ws_role_dens(const float* W, float* xch, const sde4_model_desc& d, const CostK& a, int pl, int g, int n,
                           int pd, bool active) {
  using A = typename Cfg::A;
  using M8 = typename Cfg::M8;
  using Ev = typename M8::Ev;
  using DNet = typename A::DNet;
  constexpr int NX = Cfg::NX, NPART = Cfg::NPART, NZ = Cfg::NZ, NO = Cfg::NO, L = Cfg::LN;
  constexpr int PER = (NX + L - 1) / L;
  const unsigned G = (unsigned)a.G;
  const int H = a.H;
  typename Ev::Ctx ctx;
  const int l = threadIdx.x & (L - 1);
  ctx.l = l;
  const uint32_t nzmask = noise_mask<NX>(d);
  const float* prior = xch + Cfg::VALS * NPART;  // staged by the kernel prologue
  uint32_t b0, b1;
  {
    const uint32_t* k = a.key + (size_t)pd * a.key_sp;
    brownian_key(__ldg(k), __ldg(k + 1), (uint32_t)(n + a.n_off), (uint32_t)a.n_tot, a.rng_mode, b0, b1, a.rng_chain);
  }
  // this lane's elements l, l + 8 of dw_t: validity, prior and (for eta-scaled states) the index of the scaler entry
  bool el_ok[PER];
  int el_k[PER];
  float el_prior[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int i = l + q * L;
    el_ok[q] = i < NX && ((nzmask >> i) & 1u);
    el_prior[q] = i < NX ? prior[i] : 0.0f;
    el_k[q] = (i < NX && ((A::NOUT_MASK >> i) & 1u)) ? __popc(A::NOUT_MASK & ((1u << i) - 1u)) : -1;
  }
  auto gen = [&](int t, float (&v)[PER]) {
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int i = el_ok[q] ? l + q * L : 0;
      const float z = bits_to_normal(jax_bits(b0, b1, (uint64_t)t * NX + i, (uint64_t)H * NX, a.rng_mode));
      v[q] = el_ok[q] ? z : 0.0f;
    }
  };
  const NetRef rD = A::template ref<0, L>(W);
  // ---------------- forward sweep ----------------
  float vn[PER];
  gen(0, vn);
  for (int t = 0; t < H; ++t) {
    float dwt[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) dwt[q] = vn[q];
    {  // kept for the backward sweep (the owning lane writes)
      float* row = a.dwbuf + (size_t)t * NX * G + g + (size_t)l * G;
#pragma unroll
      for (int q = 0; q < PER; ++q)
        if (active && el_ok[q]) row[(size_t)(q * L) * G] = dwt[q];
    }
    gen(t + 1 < H ? t + 1 : t, vn);  // integer work of the next step's noise, overlaps the wait for PHY
    ws_sync<Cfg::THREADS>();  // B1
    float z[NZ];
#pragma unroll
    for (int k = 0; k < NZ; ++k) z[k] = xch[(Cfg::X_Z + k) * NPART + pl];
    float dens[1];
    Ev::template fwd<DNet, false>(rD, ctx, z, dens);
    // sigma o dw for the own elements: sigma_i = prior_i * eta_k(i) (eta = 1 for states outside the noise-output mask)
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      float s = el_prior[q] * dwt[q];
      if (el_k[q] >= 0) {
        float coeff = 0.0f, offs = 0.0f;
        if constexpr (A::HAS_SCALER) {
          coeff = W[A::OFF_DERIVED + el_k[q]];
          offs = W[A::OFF_DERIVED + NO + el_k[q]];
        }
        s *= sigmoidf_((d.aggr + coeff) * dens[0] - 7.0f + offs);  // nsde.py:426
      }
      if (l + q * L < NX) xch[(Cfg::X_SDW + l + q * L) * NPART + pl] = s;
    }
    ws_sync<Cfg::THREADS>();  // B2
  }

  // ---------------- backward sweep: density value + input gradient of step t, one step ahead of its use ----------------
  float dval = 0.0f, J[NZ], z[NZ];
  auto load_inputs = [&](int t) {
    const float* xin = a.xbuf + (size_t)t * a.x_rows * G + g;
#pragma unroll
    for (int k = 0; k < NZ; ++k) z[k] = xin[(size_t)nth_bit(A::NIN_MASK, k) * G] * d.inv_red_scale;
  };
  load_inputs(H - 1);  // the forward sweep's last B2 ordered PHY's checkpoint stores before these loads
  for (int t = H - 1; t >= 0; --t) {
    Ev::template scalar_value_grad<DNet>(rD, ctx, z, dval, J, [](float) { return 1.0f; });
    load_inputs(t > 0 ? t - 1 : 0);
    ws_sync<Cfg::THREADS>();  // B3 (PHY has read the previous item by now)
    {
      float* out = xch + pl;
      if (l == 7) out[Cfg::B_DENS * NPART] = dval;
#pragma unroll
      for (int k = 0; k < NZ; ++k)
        if (l == (k & (L - 1))) out[(Cfg::B_J + k) * NPART] = J[k];
    }
    ws_sync<Cfg::THREADS>();  // B4
  }
  ws_sync<Cfg::THREADS>();  // (PHY's last gradient row for the RES role)
}

// ---- RES role: 4 lanes per particle; residual force / moment networks ---------------------------------------------------
template <class Cfg>
SDE4_DEV void ws_role_res(const float* W, float* xch, const sde4_model_desc& d, const CostK& a, const float* ftape, int pl,
                          int g, bool active) {
  using A = typename Cfg::A;
  using M8 = typename Cfg::MR;  // this role's lane-split model (4 lanes per particle)
  using Ev = typename M8::Ev;
  using NetA = typename A::NetA;
  using NetB = typename A::NetB;
  using RS = typename Cfg::RS;
  constexpr int NPART = Cfg::NPART, L = Cfg::LR;
  const unsigned G = (unsigned)a.G;
  const int H = a.H;
  typename Ev::Ctx ctx;
  const int l = threadIdx.x & (L - 1);
  ctx.l = l;
  const NetRef rA = A::template ref<1, L>(W), rB = A::template ref<2, L>(W);
  // ---------------- forward sweep ----------------
  for (int t = 0; t < H; ++t) {
    ws_sync<Cfg::THREADS>();  // B1
    float in6[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) in6[k] = xch[(Cfg::X_IN6 + k) * NPART + pl];
    float Fn[3] = {0.0f, 0.0f, 0.0f}, Mn[3] = {0.0f, 0.0f, 0.0f};
    if constexpr (RS::HAS_F) {
      float in[RS::NIA];
#pragma unroll
      for (int k = 0; k < NetA::IN; ++k) in[k] = in6[nth_bit(A::A_MASK, k)];
      Ev::template fwd<NetA, false>(rA, ctx, in, Fn);
    }
    if constexpr (RS::HAS_M) {
      float in[RS::NIB];
#pragma unroll
      for (int k = 0; k < NetB::IN; ++k) in[k] = in6[nth_bit(A::B_MASK, k)];
      Ev::template fwd<NetB, false>(rB, ctx, in, Mn);
    }
    if (l < 3) {  // outputs are replicated over the lanes: lanes 0..2 publish one component each
      xch[(Cfg::X_FN + l) * NPART + pl] = l == 0 ? Fn[0] : (l == 1 ? Fn[1] : Fn[2]);
      xch[(Cfg::X_MN + l) * NPART + pl] = l == 0 ? Mn[0] : (l == 1 ? Mn[1] : Mn[2]);
    }
    ws_sync<Cfg::THREADS>();  // B2
  }

  // ---------------- backward sweep ----------------
  float Fn[3] = {0.0f, 0.0f, 0.0f};
  float dA1[M8::TB1A], dA2[M8::TB2A], dB1[M8::TB1B], dB2[M8::TB2B];
  float in6[6];
  auto load_inputs = [&](int t) {
    const float* ft = ftape + (size_t)t * 6 * G + g;
#pragma unroll
    for (int k = 0; k < 6; ++k) in6[k] = ft[(size_t)k * G];
  };
  // forward passes of step t keeping act' of the own hidden units (cotangent independent: one step ahead)
  auto indep = [&]() {
    if constexpr (RS::HAS_F) {
      float in[RS::NIA];
#pragma unroll
      for (int k = 0; k < NetA::IN; ++k) in[k] = in6[nth_bit(A::A_MASK, k)];
      Ev::template fwd<NetA, true>(rA, ctx, in, Fn);
#pragma unroll
      for (int m = 0; m < M8::TB1A; ++m) dA1[m] = ctx.d1[m];
#pragma unroll
      for (int m = 0; m < M8::TB2A; ++m) dA2[m] = ctx.d2[m];
    }
    if constexpr (RS::HAS_M) {
      float in[RS::NIB], out[3];
#pragma unroll
      for (int k = 0; k < NetB::IN; ++k) in[k] = in6[nth_bit(A::B_MASK, k)];
      Ev::template fwd<NetB, true>(rB, ctx, in, out);
#pragma unroll
      for (int m = 0; m < M8::TB1B; ++m) dB1[m] = ctx.d1[m];
#pragma unroll
      for (int m = 0; m < M8::TB2B; ++m) dB2[m] = ctx.d2[m];
    }
  };
  // Gradient row of step t: fixed-order sum over the CTA's particles of what PHY left in the exchange buffer, by the
  // first RES warp: lane = (control j, quarter of the particles); serial over the quarter, two butterfly rounds.
  constexpr int NU = Cfg::NU, QP = NPART / 4;
  const int rlane = threadIdx.x & 31, rj = rlane >> 2, rq = rlane & 3;
  const bool sum_warp = a.warp_reduce && (int)(threadIdx.x >> 5) == 1 + Cfg::WN;
  auto sum_row = [&](int t) {
    if (!sum_warp) return;  // warp-uniform
    float v = 0.0f;
    if (rj < NU) {
#pragma unroll
      for (int k = 0; k < QP; ++k) v += xch[(Cfg::B_GU + rj) * NPART + rq * QP + k];
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (rj < NU && rq == 0) a.gpart[((size_t)t * NU + rj) * a.PW + g / NPART] = v;
  };
  load_inputs(H - 1);  // the forward sweep's last B2 ordered PHY's feature stores before these loads
  for (int t = H - 1; t >= 0; --t) {
    indep();
    load_inputs(t > 0 ? t - 1 : 0);  // requested one step ahead
    ws_sync<Cfg::THREADS>();  // B3: cotangents of the network outputs are there
    if (t + 1 < H) sum_row(t + 1);  // the row PHY finished before this barrier
    float gF[3], gM[3], giA[RS::NIA], giB[RS::NIB];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      gF[k] = xch[(Cfg::B_GF + k) * NPART + pl];
      gM[k] = xch[(Cfg::B_GM + k) * NPART + pl];
    }
    if constexpr (RS::HAS_F) {
#pragma unroll
      for (int m = 0; m < M8::TB1A; ++m) ctx.d1[m] = dA1[m];
#pragma unroll
      for (int m = 0; m < M8::TB2A; ++m) ctx.d2[m] = dA2[m];
      Ev::template bwd<NetA>(rA, ctx, gF, giA);
    }
    if constexpr (RS::HAS_M) {
#pragma unroll
      for (int m = 0; m < M8::TB1B; ++m) ctx.d1[m] = dB1[m];
#pragma unroll
      for (int m = 0; m < M8::TB2B; ++m) ctx.d2[m] = dB2[m];
      Ev::template bwd<NetB>(rB, ctx, gM, giB);
    }
    {  // replicated results: lane k publishes component k
      float* out = xch + pl;
      if constexpr (RS::HAS_F) {
#pragma unroll
        for (int k = 0; k < NetA::IN; ++k)
          if (l == (k & (L - 1))) out[(Cfg::B_GIA + k) * NPART] = giA[k];
      }
      if constexpr (RS::HAS_M) {
#pragma unroll
        for (int k = 0; k < NetB::IN; ++k)
          if (l == ((k + 3) & (L - 1))) out[(Cfg::B_GIB + k) * NPART] = giB[k];
      }
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (l == ((k + 1) & (L - 1))) out[(Cfg::B_FN + k) * NPART] = Fn[k];
    }
    ws_sync<Cfg::THREADS>();  // B4
  }
  ws_sync<Cfg::THREADS>();
  sum_row(0);
}

// One CTA = NPART particles: warp 0 is the PHY role, then NPART/4 DENS warps, then NPART/8 RES warps.
// Requirements checked by the launcher: value + gradient, in-kernel threefry noise, no state constraints.
template <class A, int NPART, bool PLAINC>
__global__ void __launch_bounds__((WsCfg<A, NPART>::THREADS), (32 / NPART))  // 32 particles per SM resident (<= 16 warps: 128 registers)
    k_cost_ws(const __grid_constant__ sde4_model_desc d, const __grid_constant__ sde4_cost_desc cd,
              const __grid_constant__ sde4_cost_desc td, const __grid_constant__ CostK a, float* ftape) {
  using Cfg = WsCfg<A, NPART>;
  extern __shared__ __align__(16) float W[];
  __shared__ __align__(8) uint64_t mbar;
  stage_params_bulk(W, a.params, A::PARAM_TOTAL * sizeof(float), &mbar);
  Diffusion<typename Cfg::M8>::derive(W);
  float* xch = W + A::SMEM_FLOATS;
  if (threadIdx.x < Cfg::N_PRIOR) xch[Cfg::VALS * NPART + threadIdx.x] = threadIdx.x < 13 ? d.noise_prior[threadIdx.x] : 0.0f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool phy = warp == 0, dens = warp <= Cfg::WN;
  const int pl = phy ? (lane & (NPART - 1))
                     : (dens ? ((warp - 1) * 32 + lane) / Cfg::LN : ((warp - 1 - Cfg::WN) * 32 + lane) / Cfg::LR);
  const int g0 = blockIdx.x * NPART + pl;
  // inactive lanes (beyond G; PHY lanes beyond NPART) shadow a valid sample, contribute nothing
  const bool active = g0 < a.G && (!phy || lane < NPART);
  const int g = g0 < a.G ? g0 : a.G - 1;
  const int p = g / a.N, n = g - p * a.N;
  const int pd = p % a.data_mod;
  if (a.mask) {  // CTA-uniform exit (the roles meet at CTA-wide barriers)
    const int on = __ldg(a.mask + p) != 0;
    if (!__syncthreads_or(on)) return;
  }
  if (phy) ws_role_phy<Cfg, PLAINC>(W, xch, d, cd, td, a, ftape, pl, g, p, pd, active);
  else if (dens) ws_role_dens<Cfg>(W, xch, d, a, pl, g, n, pd, active);
  else ws_role_res<Cfg>(W, xch, d, a, ftape, pl, g, active);
}

}  // namespace sde4

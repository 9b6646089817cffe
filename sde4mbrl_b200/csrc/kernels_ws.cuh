// kernels_ws.cuh -- warp-specialised variant of the fused rollout + cost + adjoint kernel (K2) of the multirotor SDE
// for launches that are bound by the serial instruction chain of a step (one MPC problem with a few thousand particles:
// BASELINE configs[2]).
//
// Same mathematics as k_cost<RotorModel, GRAD = true> (kernels.cuh; sde4mbrl/nsde.py:1485-1527,1576-1586 under
// jax.value_and_grad, sde4mbrl/apg.py:82,224; drift sde_rotor_model.py:427-458), different mapping.  In the lane-split
// kernels every one of the 8 lanes of a particle repeats the scalar work (quaternion physics, projection, stage cost,
// address arithmetic): 40 % of the 289 M warp instructions of a C3 launch, eight times over
// (profiles/r1_m_kcost_ncu.txt).  Here a CTA owns 16 particles and its three warps have ROLES:
//
//   PHY   one warp, ONE LANE PER PARTICLE: the state x, the Euler-Maruyama update, projection, checkpoints, stage cost,
//         and in the backward sweep the cotangent lambda and the whole physics adjoint -- nothing is replicated;
//   DENS  one warp per 16 particles: the diffusion density network on the TENSOR CORES (mma.sync m16n8k8, 3xTF32,
//         the 16 particles are the M rows; see MmaNet below); in the backward sweep the density value + input gradient
//         (cotangent independent, so it runs one step ahead of the cotangent chain and never sits on it);
//   RES   one warp per 16 particles: the residual force and moment networks on the tensor cores (forward; in the
//         backward sweep their forward passes one step ahead, then the backward passes for the cotangents PHY hands
//         over), the in-kernel threefry noise, and the fixed-order sum of the control-gradient rows over the CTA.
//
// PHY and the network roles alternate within a step -- v_b -> networks -> forces is a true dependency -- and meet at
// two CTA-wide barriers per step, exchanging a few dozen floats per particle through shared memory.
//
// (Earlier attempts: two symmetric 8-lane roles, profiles/r2_a_ws_symmetric_ncu.txt; PHY + FP32 lane-split network roles,
// profiles/r2_b_ws_fp32roles_ncu.txt -- the second one was bound by the LSU data pipe, hence the tensor cores.)
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

  // The same drift in two parts for the PHY role: drift_post_a is everything that needs no network output (motor thrust and
  // moments, gyroscopic term, quaternion kinematics) and runs while the networks are evaluated; drift_post_b adds the
  // residual force / moment after the hand-over barrier.
  struct DriftPart {
    float T, Ma[3];  // total thrust; (motor moments - w x (I w)) per axis
    Quat qd;
  };
  SDE4_DEV static void drift_post_a(const float* __restrict__ S, const float (&x)[NX], const float (&u)[NU], DriftPart& p) {
    const Quat q{x[6], x[7], x[8], x[9]};
    float T = 0.0f, Mm[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int i = 0; i < NU; ++i) {  // motor_models.py:74-95
      const float th = thrust(S, u[i]);
      T = fmaf(Mixer<NU>::m(3, i), th, T);
      Mm[0] = fmaf(Mixer<NU>::m(0, i) * S[4], th, Mm[0]);
      Mm[1] = fmaf(Mixer<NU>::m(1, i) * S[5], th, Mm[1]);
      Mm[2] = fmaf(Mixer<NU>::m(2, i) * S[6], th, Mm[2]);
    }
    const float w0 = x[10], w1 = x[11], w2 = x[12];
    const float Iw0 = S[1] * w0, Iw1 = S[2] * w1, Iw2 = S[3] * w2;
    p.T = T;
    p.Ma[0] = Mm[0] - (w1 * Iw2 - w2 * Iw1);
    p.Ma[1] = Mm[1] - (w2 * Iw0 - w0 * Iw2);
    p.Ma[2] = Mm[2] - (w0 * Iw1 - w1 * Iw0);
    p.qd = qmul_qv(q, w0, w1, w2);
  }
  SDE4_DEV static void drift_post_b(const float* __restrict__ S, const sde4_model_desc& d, const float (&x)[NX],
                                    const float (&vb)[3], const float (&Fn)[3], const float (&Mn)[3], const DriftPart& p,
                                    float (&f)[NX]) {
    const Quat q{x[6], x[7], x[8], x[9]};
    float Fb[3] = {Fn[0], Fn[1], Fn[2]};
    if constexpr (DRAG) {  // :264-277
      Fb[0] += -S[11] * vb[0];
      Fb[1] += -S[12] * vb[1];
      Fb[2] += -S[13] * vb[2] + S[14] * (vb[0] * vb[0] + vb[1] * vb[1]);
    }
    Fb[2] += p.T;
    const Quat t2 = qmul_qv(q, Fb[0], Fb[1], Fb[2]);
    const Quat r2 = qmul(t2, qconj(q));
    const float inv_m = fast_rcp(S[0]), g = d.cfloat[0];
    f[0] = x[3];
    f[1] = x[4];
    f[2] = x[5];
    f[3] = r2.x * inv_m;
    f[4] = r2.y * inv_m;
    f[5] = fmaf(r2.z, inv_m, -g);
    f[10] = (p.Ma[0] + Mn[0]) * fast_rcp(S[1]);
    f[11] = (p.Ma[1] + Mn[1]) * fast_rcp(S[2]);
    f[12] = (p.Ma[2] + Mn[2]) * fast_rcp(S[3]);
    f[6] = 0.5f * p.qd.w;
    f[7] = 0.5f * p.qd.x;
    f[8] = 0.5f * p.qd.y;
    f[9] = 0.5f * p.qd.z;
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
  // second half, in two parts.  vjp_post_a: what needs NO network result (rigid-body terms, quaternion kinematics, the
  // motor polynomial) -- the PHY role runs it while the residual networks' backward passes are in flight, so that only
  // vjp_post_b (everything fed by the networks' input cotangents giA / giB and the residual force Fn) is left on the
  // serial path after the hand-over barrier.
  struct PostPart {
    Quat gq;
    float gw[3];
  };
  SDE4_DEV static void vjp_post_a(const float* __restrict__ S, const float (&x)[NX], const float (&u)[NU],
                                  const Scaled<NX>& vf, const AdjTmp& a, PostPart& pp, float (&gx)[NX], float (&gu)[NU]) {
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
    pp.gq = gq;
    pp.gw[0] = gw[0];
    pp.gw[1] = gw[1];
    pp.gw[2] = gw[2];
  }
  SDE4_DEV static void vjp_post_b(const float* __restrict__ S, const sde4_model_desc& d, const float (&x)[NX],
                                  const float (&u)[NU], const Quat& t1, const float (&vb)[3], const float (&Fn)[3],
                                  const AdjTmp& a, const float (&giA)[NIA], const float (&giB)[NIB], const PostPart& pp,
                                  float (&gx)[NX]) {
    const Quat q{x[6], x[7], x[8], x[9]};
    const Quat qc = qconj(q);
    Quat gq = pp.gq;
    float gw[3] = {pp.gw[0], pp.gw[1], pp.gw[2]};
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
// Tensor-core evaluation of hk.nets.MLP (two hidden layers) for 16 particles per warp.
//
// mma.sync.m16n8k8 (TF32 inputs, FP32 accumulate) with the 3xTF32 split a = a_hi + a_lo, b = b_hi + b_lo,
// a*b ~ a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (relative error ~2^-21: inside the FP32 parity tolerance, unlike a single
// TF32 product at 2^-11).  M = 16 particles (rows), N = units of the layer in tiles of 8, K = units of the layer below.
// The accumulator tile of one layer IS the A operand of the next one: thread (g = lane/4, t = lane%4) holds columns
// 2t, 2t+1 of rows g, g+8 of every 8-wide tile, and the A operand wants K-indices t, t+4 of the same rows -- so the
// weight fragments are laid out with K relabelled (k = t -> unit 8j+2t, k = t+4 -> unit 8j+2t+1) and no value ever moves
// between lanes or through shared memory between layers.  Weight fragments (hi | lo) are pre-split once per CTA into
// a fragment image in shared memory, one LDS.128 per tile and thread: the FP32 lane-split evaluator pulled every
// weight through LDS.128 once per 4 FMAs, which left the LSU data pipe 68-79 % busy (profiles/r2_b_ws_fp32roles_ncu.txt).
// ----------------------------------------------------------------------------------------------------------------
// x = hi + lo with hi on the TF32 grid.  hi is x with the 13 low mantissa bits cleared (one LOP3; cvt.rna.tf32.f32 is
// emulated on sm_100a -- FSETP / SEL / IADD3 / LOP3 per value -- and rounding buys nothing here: x - hi is exact
// either way); lo goes to the tensor core as the FP32 word, which reads its upper 19 bits, leaving a residual of
// 2^-11 |lo| <= 2^-21 |x|, the same order as the dropped lo*lo product.
SDE4_DEV void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
SDE4_DEV void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// A operand (hi | lo) of one k-step
struct AFrag {
  uint32_t hi[4], lo[4];
  // from the inputs in the operand's own order: a0 (row g, k = t), a1 (row g+8, k = t), a2 (g, t+4), a3 (g+8, t+4)
  SDE4_DEV void from_values(const float (&a)[4]) {
#pragma unroll
    for (int r = 0; r < 4; ++r) tf32_split(a[r], hi[r], lo[r]);
  }
  // from an accumulator tile (relabelled K): c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
  SDE4_DEV void from_tile(const float (&c)[4]) {
    tf32_split(c[0], hi[0], lo[0]);
    tf32_split(c[2], hi[1], lo[1]);
    tf32_split(c[1], hi[2], lo[2]);
    tf32_split(c[3], hi[3], lo[3]);
  }
};
// Accumulator of one 16x8 output tile, one register tile per 3xTF32 term: the three products of a k-step go to
// different accumulators, so the dependent MMA chain of a layer is its number of k-steps, not three times that
// (the density value + gradient pass had 51 MMAs in a row on its critical path).
// Accumulator of one 16x8 output tile.  FAST: one register tile per 3xTF32 term and per k-step parity -- the dependent
// MMA chain of a layer is then ceil(k-steps / 2) instead of 3 * k-steps (HMMA latency is ~45 cycles here and the density
// value + gradient pass had 51 MMAs in a row); the compact form keeps one tile (smallest code).
template <bool FAST>
struct AccT {
  static constexpr int NTILE = FAST ? 6 : 1;
  float t[NTILE][4];
  SDE4_DEV void init(float c0, float c1) {
    t[0][0] = t[0][2] = c0;
    t[0][1] = t[0][3] = c1;
#pragma unroll
    for (int k = 1; k < NTILE; ++k) t[k][0] = t[k][1] = t[k][2] = t[k][3] = 0.0f;
  }
  // += A * B for one 8x8 weight tile (x, y = b0, b1 hi; z, w = b0, b1 lo) of k-step j
  SDE4_DEV void mma(const AFrag& a, const float4& b, int j) {
    if constexpr (FAST) {
      const int o = (j & 1) * 3;
      mma_tf32(t[o], a.hi, __float_as_uint(b.x), __float_as_uint(b.y));
      mma_tf32(t[o + 1], a.hi, __float_as_uint(b.z), __float_as_uint(b.w));
      mma_tf32(t[o + 2], a.lo, __float_as_uint(b.x), __float_as_uint(b.y));
    } else {  // small terms first
      mma_tf32(t[0], a.lo, __float_as_uint(b.x), __float_as_uint(b.y));
      mma_tf32(t[0], a.hi, __float_as_uint(b.z), __float_as_uint(b.w));
      mma_tf32(t[0], a.hi, __float_as_uint(b.x), __float_as_uint(b.y));
    }
  }
  SDE4_DEV void sum(float (&c)[4], int ksteps) const {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if constexpr (FAST) {
        float v = t[0][r] + (t[1][r] + t[2][r]);
        if (ksteps > 1) v += t[3][r] + (t[4][r] + t[5][r]);
        c[r] = v;
      } else {
        c[r] = t[0][r];
      }
    }
  }
};

// which passes use the FAST accumulators.  Measured on C3 (scripts/ws_time.py): none 0.358 ms, forward passes 0.358 ms,
// + residual-net backward 0.385 ms, all 0.476 ms -- the launch is bound by instruction fetch (four roles share the SM's
// instruction cache), not by the MMA chains, so the compact accumulators stay the default.
#ifndef SDE4_WS_FAST_FWD
#define SDE4_WS_FAST_FWD false
#endif
#ifndef SDE4_WS_FAST_RESB
#define SDE4_WS_FAST_RESB false
#endif
#ifndef SDE4_WS_FAST_DENSB
#define SDE4_WS_FAST_DENSB false
#endif

template <class N>
struct MmaNet {
  static constexpr int NT1 = N::H1 / 8, NT2 = N::H2 / 8;
  static_assert(N::H1 % 8 == 0 && N::H2 % 8 == 0 && N::IN <= 8 && N::OUT <= 4, "tile shapes of the tensor-core evaluator");
  // tiles of the fragment image, in this order
  static constexpr int T_L1 = 0, T_L2 = T_L1 + NT1, T_L3 = T_L2 + NT1 * NT2, T_W2T = T_L3 + NT2, T_W1T = T_W2T + NT2,
                       T_W0T = T_W1T + NT2 * NT1, NTILES = T_W0T + NT1;
  SDE4_DEV static float w0(const float* w, int k, int u) { return k < N::IN ? w[N::OFF_W0 + k * N::H1P + u] : 0.0f; }
  SDE4_DEV static float w1(const float* w, int u1, int u2) { return w[N::OFF_W1 + u1 * N::H2P + u2]; }
  SDE4_DEV static float w2(const float* w, int u2, int o) { return o < N::OUT ? w[N::OFF_W2 + o * N::H2P + u2] : 0.0f; }

  // fragment image [NTILES][32] from the packed network image w (nets.cuh layout)
  SDE4_DEV static void build(float4* tiles, const float* w, int tid, int nthreads) {
    for (int e = tid; e < NTILES * 32; e += nthreads) {
      const int T = e >> 5, ln = e & 31, g = ln >> 2, t = ln & 3;
      float v0, v1;
      if (T < T_L2) {  // layer 1: K = inputs (their own order), N = h1 units
        const int n = T - T_L1;
        v0 = w0(w, t, 8 * n + g);
        v1 = w0(w, t + 4, 8 * n + g);
      } else if (T < T_L3) {  // layer 2: K = h1 units (relabelled), N = h2 units
        const int q = T - T_L2, j = q / NT2, n = q - j * NT2;
        v0 = w1(w, 8 * j + 2 * t, 8 * n + g);
        v1 = w1(w, 8 * j + 2 * t + 1, 8 * n + g);
      } else if (T < T_W2T) {  // layer 3: K = h2 units (relabelled), N = outputs
        const int j = T - T_L3;
        v0 = w2(w, 8 * j + 2 * t, g);
        v1 = w2(w, 8 * j + 2 * t + 1, g);
      } else if (T < T_W1T) {  // backward through layer 3: K = outputs (their own order), N = h2 units
        const int n = T - T_W2T;
        v0 = w2(w, 8 * n + g, t);
        v1 = w2(w, 8 * n + g, t + 4);
      } else if (T < T_W0T) {  // backward through layer 2: K = h2 units (relabelled), N = h1 units
        const int q = T - T_W1T, j = q / NT1, n = q - j * NT1;
        v0 = w1(w, 8 * n + g, 8 * j + 2 * t);
        v1 = w1(w, 8 * n + g, 8 * j + 2 * t + 1);
      } else {  // backward through layer 1: K = h1 units (relabelled), N = inputs
        const int j = T - T_W0T;
        v0 = w0(w, g, 8 * j + 2 * t);
        v1 = w0(w, g, 8 * j + 2 * t + 1);
      }
      uint32_t h0, l0, h1, l1;
      tf32_split(v0, h0, l0);
      tf32_split(v1, h1, l1);
      tiles[e] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
    }
  }

  // forward for the 16 particles of this warp: out tile = columns 2t, 2t+1 (< OUT) of rows g, g+8.
  // KEEP: act' of both hidden layers stays in d1 / d2 (accumulator layout) for the backward pass.
  template <bool KEEP, bool FAST>
  SDE4_DEV static void fwd(const float4* tiles, const float* w, int lane, const float (&ain)[4], float (&out)[4],
                           float (&d1)[NT1][4], float (&d2)[NT2][4]) {
    const int t = lane & 3;
    AFrag a;
    a.from_values(ain);
    float c1[NT1][4];
#pragma unroll
    for (int n = 0; n < NT1; ++n) {
      AccT<FAST> acc;
      const float2 b = *reinterpret_cast<const float2*>(w + N::OFF_B0 + 8 * n + 2 * t);
      acc.init(b.x, b.y);
      acc.mma(a, tiles[(T_L1 + n) * 32 + lane], 0);
      acc.sum(c1[n], 1);
    }
#pragma unroll
    for (int n = 0; n < NT1; ++n) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (KEEP) act_fwd_d<N::ACT>(c1[n][r], c1[n][r], d1[n][r]);
        else c1[n][r] = act_fwd<N::ACT>(c1[n][r]);
      }
    }
    float c2[NT2][4];
    {
      AccT<FAST> acc[NT2];
#pragma unroll
      for (int n = 0; n < NT2; ++n) {
        const float2 b = *reinterpret_cast<const float2*>(w + N::OFF_B1 + 8 * n + 2 * t);
        acc[n].init(b.x, b.y);
      }
#pragma unroll
      for (int j = 0; j < NT1; ++j) {
        a.from_tile(c1[j]);
#pragma unroll
        for (int n = 0; n < NT2; ++n) acc[n].mma(a, tiles[(T_L2 + j * NT2 + n) * 32 + lane], j);
      }
#pragma unroll
      for (int n = 0; n < NT2; ++n) acc[n].sum(c2[n], NT1);
    }
#pragma unroll
    for (int n = 0; n < NT2; ++n) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (KEEP) act_fwd_d<N::ACT>(c2[n][r], c2[n][r], d2[n][r]);
        else c2[n][r] = act_fwd<N::ACT>(c2[n][r]);
      }
    }
    if constexpr (N::OUT == 1) {
      // scalar output (the density network): an FP32 dot product over the thread's own columns and two butterfly rounds
      // over the quad instead of NT2 k-steps of three dependent MMAs for one useful column in eight -- ~100 instead of
      // ~500 cycles on the chain of the forward sweep, where this network is the last arriver of every step.
      // Every thread of the quad ends up with the value of rows g (out[0]) and g + 8 (out[2]).
      float s0a = 0.0f, s0b = 0.0f, s1a = 0.0f, s1b = 0.0f;
#pragma unroll
      for (int n = 0; n < NT2; ++n) {
        const float2 v = *reinterpret_cast<const float2*>(w + N::OFF_W2 + 8 * n + 2 * t);
        s0a = fmaf(c2[n][0], v.x, s0a);
        s0b = fmaf(c2[n][1], v.y, s0b);
        s1a = fmaf(c2[n][2], v.x, s1a);
        s1b = fmaf(c2[n][3], v.y, s1b);
      }
      float s0 = s0a + s0b, s1 = s1a + s1b;
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
      s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      const float b2 = w[N::OFF_B2];
      out[0] = s0 + b2;
      out[2] = s1 + b2;
      out[1] = out[3] = 0.0f;
    } else {
      AccT<FAST> acc;
      acc.init(2 * t < N::OUT ? w[N::OFF_B2 + 2 * t] : 0.0f, 2 * t + 1 < N::OUT ? w[N::OFF_B2 + 2 * t + 1] : 0.0f);
#pragma unroll
      for (int j = 0; j < NT2; ++j) {
        a.from_tile(c2[j]);
        acc.mma(a, tiles[(T_L3 + j) * 32 + lane], j);
      }
      acc.sum(out, NT2);
    }
  }

  // g2 (cotangent of the pre-activation of layer 2, accumulator layout) -> input cotangent tile:
  // gin = columns 2t, 2t+1 (< IN) of rows g, g+8
  template <bool FAST>
  SDE4_DEV static void bwd_from_g2(const float4* tiles, int lane, const float (&g2)[NT2][4], const float (&d1)[NT1][4],
                                   float (&gin)[4]) {
    AFrag a;
    float g1[NT1][4];
    {
      AccT<FAST> acc[NT1];
#pragma unroll
      for (int n = 0; n < NT1; ++n) acc[n].init(0.0f, 0.0f);
#pragma unroll
      for (int j = 0; j < NT2; ++j) {
        a.from_tile(g2[j]);
#pragma unroll
        for (int n = 0; n < NT1; ++n) acc[n].mma(a, tiles[(T_W1T + j * NT1 + n) * 32 + lane], j);
      }
#pragma unroll
      for (int n = 0; n < NT1; ++n) acc[n].sum(g1[n], NT2);
    }
    AccT<FAST> acc;
    acc.init(0.0f, 0.0f);
#pragma unroll
    for (int j = 0; j < NT1; ++j) {
#pragma unroll
      for (int r = 0; r < 4; ++r) g1[j][r] *= d1[j][r];
      a.from_tile(g1[j]);
      acc.mma(a, tiles[(T_W0T + j) * 32 + lane], j);
    }
    acc.sum(gin, NT1);
  }
  // output cotangent (A operand in its own order: a0 = gout[g][t], a1 = gout[g+8][t], a2 = a3 = 0) -> input cotangent
  template <bool FAST>
  SDE4_DEV static void bwd(const float4* tiles, int lane, const float (&agout)[4], const float (&d1)[NT1][4],
                           const float (&d2)[NT2][4], float (&gin)[4]) {
    AFrag a;
    a.from_values(agout);
    float g2[NT2][4];
#pragma unroll
    for (int n = 0; n < NT2; ++n) {
      AccT<FAST> acc;
      acc.init(0.0f, 0.0f);
      acc.mma(a, tiles[(T_W2T + n) * 32 + lane], 0);
      acc.sum(g2[n], 1);
#pragma unroll
      for (int r = 0; r < 4; ++r) g2[n][r] *= d2[n][r];
    }
    bwd_from_g2<FAST>(tiles, lane, g2, d1, gin);
  }
  // scalar-output network: value (column 0 of `out`, threads with t == 0) and the input gradient d out / d in
  // `mid()` runs between the forward pass and the gradient pass (the two density producers of the backward sweep join a
  // CTA barrier there: an item is made over two steps)
  template <bool FAST, class Mid>
  SDE4_DEV static void value_grad(const float4* tiles, const float* w, int lane, const float (&ain)[4], float (&out)[4],
                                  float (&gin)[4], Mid&& mid) {
    static_assert(N::OUT == 1, "scalar-output network expected");
    const int t = lane & 3;
    float d1[NT1][4], d2[NT2][4];
    fwd<true, FAST>(tiles, w, lane, ain, out, d1, d2);
    mid();
#pragma unroll
    for (int n = 0; n < NT2; ++n) {
      const float2 v = *reinterpret_cast<const float2*>(w + N::OFF_W2 + 8 * n + 2 * t);
      d2[n][0] *= v.x;
      d2[n][1] *= v.y;
      d2[n][2] *= v.x;
      d2[n][3] *= v.y;
    }
    bwd_from_g2<FAST>(tiles, lane, d2, d1, gin);
  }
};

// ----------------------------------------------------------------------------------------------------------------
template <class A_>
struct WsCfg {
  using A = A_;
  using M8 = RotorModel<A, EvL<8>>;  // diffusion constants and projection of the lane-split model
  using RS = RotorSplit<A>;
  using DNet = typename A::DNet;
  using NetA = typename A::NetA;
  using NetB = typename A::NetB;
  using MD = MmaNet<DNet>;
  using MA = MmaNet<NetA>;
  using MB = MmaNet<NetB>;
  static_assert(A::MODEL == SDE4_MODEL_ROTOR && A::HAS_DENSITY && !Diffusion<M8>::CDN && NetA::PRESENT && NetB::PRESENT,
                "multirotor SDE with a density net and both residual nets");
  static constexpr int NPART = 16;                 // particles per CTA = rows of one MMA tile
  static constexpr int THREADS = 128;              // warp 0 PHY, warp 1 DENS, warp 2 RES, warp 3 AUX (one role per scheduler)
  static constexpr int NX = 13, NU = A::NU, NZ = DNet::IN, NO = A::NO;
  static constexpr int NIA = RS::NIA, NIB = RS::NIB;
  // fragment image
  static constexpr int TILE_D = 0, TILE_A = MD::NTILES, TILE_B = TILE_A + MA::NTILES, NTILES = TILE_B + MB::NTILES;
  // exchange buffer, [value][NPART]
  static constexpr int X_IN6 = 0, X_Z = 6, X_FN = X_Z + NZ, X_MN = X_FN + 3, X_ETA = X_MN + 3, X_DW = X_ETA + NO,
                       FWD_VALS = X_DW + 3 * NX;  // three dw slots in rotation
  // (density value + input gradient: two slots in alternation, B_DENS + (t & 1) * B_DSLOT -- the DENS role writes the item
  //  of step t any time before B4(t) while PHY still reads the item of step t + 1, so it stays out of B3)
  static constexpr int B_GF = 0, B_GM = 3, B_GIA = 6, B_GIB = B_GIA + NIA, B_FN = B_GIB + NIB, B_DENS = B_FN + 3,
                       B_J = B_DENS + 1, B_DSLOT = 1 + NZ, B_GU = B_DENS + 2 * B_DSLOT, BWD_VALS = B_GU + NU;
  static constexpr int VALS = FWD_VALS > BWD_VALS ? FWD_VALS : BWD_VALS;
  static constexpr int N_MISC = 32;  // noise prior [16] | list of the states that carry noise [16]
  static constexpr int OFF_TILES = A::SMEM_FLOATS;                 // float4-aligned (SMEM_FLOATS is a multiple of 4)
  static constexpr int OFF_XCH = OFF_TILES + NTILES * 32 * 4;
  static constexpr int OFF_MISC = OFF_XCH + VALS * NPART;
  static constexpr size_t smem_bytes() { return sizeof(float) * ((size_t)OFF_MISC + N_MISC); }
};

#ifdef SDE4_WS_TRACE
// tuning aid: arrival / departure clocks of every barrier of CTA 64, per role (scripts/ws_trace.py)
__device__ long long g_ws_trace[4][2 * 512];
__device__ int g_ws_trace_n[4];
#endif
// Named barriers (the roles run different code, so no __syncthreads()).
template <int ID, int COUNT>
SDE4_DEV void ws_bar() {
#ifdef SDE4_WS_TRACE
  const bool rec = blockIdx.x == 64 && (threadIdx.x & 31) == 0;
  const int role = threadIdx.x >> 5;
  int slot = 0;
  if (rec) {
    slot = g_ws_trace_n[role];
    if (slot < 512) g_ws_trace[role][2 * slot] = clock64();
  }
#endif
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
#ifdef SDE4_WS_TRACE
  if (rec) {
    if (slot < 512) g_ws_trace[role][2 * slot + 1] = clock64();
    g_ws_trace_n[role] = slot + 1;
  }
#endif
}
// B2 / B4 (networks -> PHY hand-over): the whole CTA
template <int THREADS>
SDE4_DEV void ws_sync() {
  ws_bar<1, THREADS>();
}
// B1 / B3 (PHY -> networks hand-over) involve the first three warps only: the AUX role never reads what PHY publishes
// there, and keeping it out lets it generate noise during BOTH halves of a step.
SDE4_DEV void ws_sync3() { ws_bar<2, 96>(); }
// B3 of the backward sweep: PHY and RES only -- the DENS role works one step ahead through two output slots and meets the
// others at B4 alone, which takes its value + gradient pass (the longest item of a backward step) off the PHY -> RES path.
SDE4_DEV void ws_sync_pr() { ws_bar<3, 64>(); }

// A operand of a network input read from the exchange buffer: value (row p, k) at xch[(base + idx(k)) * NPART + p]
template <int NPART, class IdxFn>
SDE4_DEV void a_from_xch(const float* xch, int base, int nk, int lane, IdxFn idx, float (&a)[4]) {
  const int g = lane >> 2, t = lane & 3;
  a[0] = t < nk ? xch[(base + idx(t)) * NPART + g] : 0.0f;
  a[1] = t < nk ? xch[(base + idx(t)) * NPART + g + 8] : 0.0f;
  a[2] = t + 4 < nk ? xch[(base + idx(t + 4)) * NPART + g] : 0.0f;
  a[3] = t + 4 < nk ? xch[(base + idx(t + 4)) * NPART + g + 8] : 0.0f;
}
// k-th set bit of a compile-time mask for a run-time k (k < popc(MASK) <= 8)
template <uint32_t MASK>
SDE4_DEV int nth_bit_rt(int k) {
  int r = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (q < popc(MASK) && k == q) r = nth_bit(MASK, q < popc(MASK) ? q : 0);
  return r;
}
// publish columns 2t, 2t+1 (< n) of rows g, g+8 of an output tile
template <int NPART>
SDE4_DEV void tile_to_xch(float* xch, int base, int n, int lane, const float (&c)[4]) {
  const int g = lane >> 2, t = lane & 3;
  if (2 * t < n) {
    xch[(base + 2 * t) * NPART + g] = c[0];
    xch[(base + 2 * t) * NPART + g + 8] = c[2];
  }
  if (2 * t + 1 < n) {
    xch[(base + 2 * t + 1) * NPART + g] = c[1];
    xch[(base + 2 * t + 1) * NPART + g + 8] = c[3];
  }
}

// ---- PHY role: one lane per particle ---------------------------------------------------------------------------------
template <class Cfg, bool PLAINC, bool PHYMB>
SDE4_DEV void ws_role_phy(const float* W, const float4* tiles, float* xch, const sde4_model_desc& d, const sde4_cost_desc& cd,
                          const sde4_cost_desc& td, const CostK& a, float* ftape, int pl, int g, int p, int pd, bool active) {
  using A = typename Cfg::A;
  using RS = typename Cfg::RS;
  using M8 = typename Cfg::M8;
  using MB = typename Cfg::MB;
  using NetB = typename A::NetB;
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
  ws_sync<Cfg::THREADS>();  // B0: dw_0 is there
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
    ws_sync3();  // B1: the networks may start
    if (active) {             // network inputs of the residual nets, kept for the backward sweep
      float* ft = ftape + (size_t)t * 6 * G + g;
#pragma unroll
      for (int k = 0; k < 6; ++k) ft[(size_t)k * G] = in6[k];
    }
    typename RS::DriftPart dp;
    RS::drift_post_a(S, x, u, dp);  // the network-independent part of the drift, while the networks run
    if constexpr (NetB::PRESENT && PHYMB) {
      // forward sweep only, latency regime only (PHYMB, see k_cost_ws): this warp would idle ~1 500 cycles here while the RES role evaluates its two networks back to
      // back (the longest item of a forward step) -- it evaluates the residual-moment network itself, all 32 lanes
      const int lane = threadIdx.x & 31;
      float aB[4], Mn4[4];
      float dB1[MB::NT1][4], dB2[MB::NT2][4];
      a_from_xch<NPART>(xch, Cfg::X_IN6, NetB::IN, lane, [](int k) { return nth_bit_rt<A::B_MASK>(k); }, aB);
      MB::template fwd<false, SDE4_WS_FAST_FWD>(tiles + Cfg::TILE_B * 32, W + A::OFF_NET_B, lane, aB, Mn4, dB1, dB2);
      tile_to_xch<NPART>(xch, Cfg::X_MN, 3, lane, Mn4);
    }
    // Positions and quaternion: their drift (v, 0.5 q (x) [0, w]) needs no network output and, unless the noise-output
    // mask says otherwise, no eta; dw_t was generated a step ahead.  Their Euler-Maruyama update, the projection and
    // their checkpoint rows are done here, in front of the hand-over barrier (same operations, same order: bitwise equal).
    constexpr uint32_t KIN_MASK = 0x3C7u;  // states 0-2, 6-9
    constexpr bool EARLY = (A::NOUT_MASK & KIN_MASK) == 0;
    const float* dwb = xch + (Cfg::X_DW + (t % 3) * NX) * NPART + pl;
    float xe[NX], aux = 0.0f;
    float* ckp = a.xbuf + (size_t)(t + 1) * a.x_rows * G + g;
    if constexpr (EARLY) {
#pragma unroll
      for (int k = 0; k < 3; ++k) xe[k] = fmaf(x[3 + k], dt, x[k]) + d.noise_prior[k] * dwb[k * NPART];
      xe[6] = fmaf(0.5f * dp.qd.w, dt, x[6]) + d.noise_prior[6] * dwb[6 * NPART];
      xe[7] = fmaf(0.5f * dp.qd.x, dt, x[7]) + d.noise_prior[7] * dwb[7 * NPART];
      xe[8] = fmaf(0.5f * dp.qd.y, dt, x[8]) + d.noise_prior[8] * dwb[8 * NPART];
      xe[9] = fmaf(0.5f * dp.qd.z, dt, x[9]) + d.noise_prior[9] * dwb[9 * NPART];
      M8::project(xe, aux);  // touches the quaternion only
      if (active) {
        a.auxbuf[(size_t)t * G + g] = aux;
#pragma unroll
        for (int i = 0; i < NX; ++i)
          if ((KIN_MASK >> i) & 1u) ckp[(size_t)i * G] = xe[i];
      }
    }
    ws_sync<Cfg::THREADS>();  // B2: network outputs, eta and dw_t are there
    float Fn[3], Mn[3], f[NX];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      Fn[k] = xch[(Cfg::X_FN + k) * NPART + pl];
      Mn[k] = xch[(Cfg::X_MN + k) * NPART + pl];
    }
    RS::drift_post_b(S, d, x, vb, Fn, Mn, dp, f);
    // sigma o dw: sigma_i = prior_i * eta_k(i), eta = 1 outside the noise-output mask (nsde.py:363,449)
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      if (EARLY && ((KIN_MASK >> i) & 1u)) {
        x[i] = xe[i];
      } else {
        float s = d.noise_prior[i];
        if ((A::NOUT_MASK >> i) & 1u) s *= xch[(Cfg::X_ETA + popc(A::NOUT_MASK & ((1u << i) - 1u))) * NPART + pl];
        x[i] = fmaf(f[i], dt, x[i]) + s * dwb[i * NPART];
      }
    }
    if constexpr (!EARLY) M8::project(x, aux);
    if (active) {
      if (!EARLY) a.auxbuf[(size_t)t * G + g] = aux;
#pragma unroll
      for (int i = 0; i < NX; ++i)
        if (!(EARLY && ((KIN_MASK >> i) & 1u))) ckp[(size_t)i * G] = x[i];
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
  const int pw = a.warp_reduce ? (g / NPART) : g;
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
    ws_sync_pr();  // B3: the network backward passes may start
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
    typename RS::PostPart pp;
    RS::vjp_post_a(S, x, u, vf, at, pp, gx, gu);  // the network-independent half of the drift adjoint
    ws_sync<Cfg::THREADS>();  // B4: input cotangents, residual force, density value + gradient are there
    float giA[Cfg::NIA], giB[Cfg::NIB], Fn[3];
#pragma unroll
    for (int k = 0; k < Cfg::NIA; ++k) giA[k] = xch[(Cfg::B_GIA + k) * NPART + pl];
#pragma unroll
    for (int k = 0; k < Cfg::NIB; ++k) giB[k] = xch[(Cfg::B_GIB + k) * NPART + pl];
#pragma unroll
    for (int k = 0; k < 3; ++k) Fn[k] = xch[(Cfg::B_FN + k) * NPART + pl];
    typename Diffusion<M8>::Tape dtp;
    const int so = (t & 1) * Cfg::B_DSLOT;
    dtp.dens = xch[(Cfg::B_DENS + so) * NPART + pl];
#pragma unroll
    for (int k = 0; k < NZ; ++k) dtp.J[k] = xch[(Cfg::B_J + so + k) * NPART + pl];
    RS::vjp_post_b(S, d, x, u, t1, vb, Fn, at, giA, giB, pp, gx);
    Diffusion<M8>::apply(W, d, x, lam, [&](int i) { return nzv[i]; }, dtp, gx, gu);
    // row t of the gradient.  When the CTA's particles share a problem the row is summed over them by the RES role
    // (30 dependent shuffle / add pairs here were half of this role's backward sweep); otherwise one partial per
    // particle, stored directly.
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

// In-kernel threefry noise of the CTA's 16 particles, shared by the two network warps (PART 0: RES, PART 1: DENS).
// A thread draws, for ITS particle (lane % 16), the noise-carrying states of list index (lane / 16) + 2 q; the q range
// is cut in two between the warps.  dw_{t+1} is generated while PHY finishes step t and stored straight into the exchange
// buffer (three slots in rotation, so no write can meet a read of the slot's previous use) and, for the backward sweep,
// the global noise buffer.  The loop over q is ROLLED, two independent threefry chains per iteration: fully unrolled the
// generator was 770 instructions of each network role's loop body, and the three roles' bodies together have to stay
// inside the 32 KB instruction cache (profiles/r2_c_ws_mma_ncu.txt: no_instruction was the top stall otherwise).
template <class Cfg, int PART>
struct WsNoise {
  static constexpr int NX = Cfg::NX, NPART = Cfg::NPART;
  float* xch;
  const CostK& a;
  const int* nzlist;
  int nnz, lane, pi, g0, H, qlo, qhi;
  bool lane_active;
  uint32_t b0, b1;
  SDE4_DEV WsNoise(float* xch_, const sde4_model_desc& d, const CostK& a_, int g0_, bool lane_active_, int n, int pd)
      : xch(xch_), a(a_), g0(g0_), lane_active(lane_active_) {
    nzlist = reinterpret_cast<const int*>(xch + (Cfg::OFF_MISC - Cfg::OFF_XCH) + 16);
    nnz = __popc(noise_mask<NX>(d));
    lane = threadIdx.x & 31;
    pi = lane & 15;
    H = a.H;
    const int nq = (nnz + 1) >> 1, mid = (nq + 1) >> 1;
    qlo = PART == 1 ? mid : 0;  // PART 2: the whole range
    qhi = PART == 0 ? mid : nq;
    const uint32_t* k = a.key + (size_t)pd * a.key_sp;
    brownian_key(__ldg(k), __ldg(k + 1), (uint32_t)(n + a.n_off), (uint32_t)a.n_tot, a.rng_mode, b0, b1, a.rng_chain);
  }
  SDE4_DEV void one(int ts, int q, bool on) const {
    const int idx = (lane >> 4) + 2 * q;
    const bool ok = on && idx < nnz;
    const int i = nzlist[ok ? idx : 0];
    const float z = bits_to_normal(jax_bits(b0, b1, (uint64_t)ts * NX + i, (uint64_t)H * NX, a.rng_mode));
    if (ok) {
      xch[(Cfg::X_DW + (ts % 3) * NX + i) * NPART + pi] = z;
      if (lane_active) a.dwbuf[((size_t)ts * NX + i) * (unsigned)a.G + g0 + pi] = z;
    }
  }
  template <int NQ>
  SDE4_DEV void gen_unrolled(int ts) const {
    // NQ independent threefry / erf_inv chains in ONE basic block (ptxas interleaves dependency chains only inside a
    // block: with the predicated stores between them the five chains ran one after the other, 3.3 k cycles per step)
    int i[NQ];
    bool ok[NQ];
    float z[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int idx = (lane >> 4) + 2 * q;
      ok[q] = idx < nnz;
      i[q] = nzlist[ok[q] ? idx : 0];
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q)
      z[q] = bits_to_normal(jax_bits(b0, b1, (uint64_t)ts * NX + i[q], (uint64_t)H * NX, a.rng_mode));
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      if (ok[q]) {
        xch[(Cfg::X_DW + (ts % 3) * NX + i[q]) * NPART + pi] = z[q];
        if (lane_active) a.dwbuf[((size_t)ts * NX + i[q]) * (unsigned)a.G + g0 + pi] = z[q];
      }
    }
  }
  SDE4_DEV void gen(int ts) const {
    if (PART == 2 && qhi == 5) {  // 9 or 10 noise-carrying states (the shipped multirotor models): five chains side by side
      gen_unrolled<5>(ts);
      return;
    }
#pragma unroll 1
    for (int q = qlo; q < qhi; q += 2) {
      one(ts, q, true);
      one(ts, q + 1, q + 1 < qhi);
    }
  }
};

// ---- DENS role: one warp, 16 particles; density network on the tensor cores, eta, half of the noise ----------------------------------------
// OWN_BWD: the role's own backward sweep follows (every item in one piece, one step ahead; many CTAs per SM) -- otherwise
// the caller continues with ws_densaux_bwd.
template <class Cfg, bool OWN_BWD>
SDE4_DEV void ws_role_dens(const float* W, const float4* tiles, float* xch, const sde4_model_desc& d, const CostK& a,
                           int g0, int gmax) {
  using A = typename Cfg::A;
  using MD = typename Cfg::MD;
  using DNet = typename A::DNet;
  constexpr int NPART = Cfg::NPART, NZ = Cfg::NZ, NO = Cfg::NO;
  const int H = a.H;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const float* wD = W + A::OFF_DENSITY;
  const float4* tD = tiles + Cfg::TILE_D * 32;
  auto ident = [](int k) { return k; };
  // ---------------- forward sweep ----------------
  ws_sync<Cfg::THREADS>();  // B0
  for (int ts = 0; ts < H; ++ts) {
    ws_sync3();  // B1
    float az[4], out[4];
    a_from_xch<NPART>(xch, Cfg::X_Z, NZ, lane, ident, az);
    float d1[MD::NT1][4], d2[MD::NT2][4];
    MD::template fwd<false, SDE4_WS_FAST_FWD>(tD, wD, lane, az, out, d1, d2);
    // density of rows g, g+8: thread t of the quad forms eta_t, eta_{t+4}
    const float dg = out[0], dg8 = out[2];  // (the scalar-output layer leaves the value with every thread of the quad)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int k = t + 4 * q;
      if (k < NO) {
        float coeff = 0.0f, offs = 0.0f;
        if constexpr (A::HAS_SCALER) {
          coeff = W[A::OFF_DERIVED + k];
          offs = W[A::OFF_DERIVED + NO + k];
        }
        xch[(Cfg::X_ETA + k) * NPART + g] = sigmoidf_((d.aggr + coeff) * dg - 7.0f + offs);  // nsde.py:426
        xch[(Cfg::X_ETA + k) * NPART + g + 8] = sigmoidf_((d.aggr + coeff) * dg8 - 7.0f + offs);
      }
    }
    ws_sync<Cfg::THREADS>();  // B2
  }
  if constexpr (!OWN_BWD) return;
  const unsigned G = (unsigned)a.G;
  // ---------------- backward sweep: density value + input gradient of step t, one step ahead of its use ----------------
  // rows of the A operand: particles g0 + g, g0 + g + 8 (clamped to the last valid sample)
  const unsigned ga = min(g0 + g, gmax), gb = min(g0 + g + 8, gmax);
  float az[4];
  auto load_inputs = [&](int ts) {
    const float* xin = a.xbuf + (size_t)ts * a.x_rows * G;
    const int k0 = nth_bit_rt<A::NIN_MASK>(t), k1 = nth_bit_rt<A::NIN_MASK>(t + 4 < NZ ? t + 4 : 0);
    az[0] = t < NZ ? xin[(size_t)k0 * G + ga] * d.inv_red_scale : 0.0f;
    az[1] = t < NZ ? xin[(size_t)k0 * G + gb] * d.inv_red_scale : 0.0f;
    az[2] = t + 4 < NZ ? xin[(size_t)k1 * G + ga] * d.inv_red_scale : 0.0f;
    az[3] = t + 4 < NZ ? xin[(size_t)k1 * G + gb] * d.inv_red_scale : 0.0f;
  };
  load_inputs(H - 1);  // the forward sweep's last B2 ordered PHY's checkpoint stores before these loads
  for (int ts = H - 1; ts >= 0; --ts) {
    float out[4], J[4];
    MD::template value_grad<SDE4_WS_FAST_DENSB>(tD, wD, lane, az, out, J, []() {});
    load_inputs(ts > 0 ? ts - 1 : 0);
    const int so = (ts & 1) * Cfg::B_DSLOT;  // PHY read this slot's previous item (step ts + 2) before B4(ts + 1)
    if (t == 0) {
      xch[(Cfg::B_DENS + so) * NPART + g] = out[0];
      xch[(Cfg::B_DENS + so) * NPART + g + 8] = out[2];
    }
    tile_to_xch<NPART>(xch, Cfg::B_J + so, NZ, lane, J);
    ws_sync<Cfg::THREADS>();  // B4
  }
  ws_sync<Cfg::THREADS>();  // (PHY's last gradient row for the RES role)
}

// ---- backward sweep of the DENS and AUX warps: the density value + input gradient of every step ---------------------------
// The item of step t depends on the checkpoint x_t only, not on the cotangent, but it is the longest single piece of a
// backward step (a value + gradient pass of the 32 x 32 network: the DENS warp alone was the last arriver at B4 in 72 % of
// the steps, PHY and RES waiting ~480 cycles for it).  The AUX warp has next to nothing to do in this sweep, so the two
// warps share the items by step parity -- par = 0 (DENS): steps H-1, H-3, ...; par = 1 (AUX): steps H-2, H-4, ... -- and
// each makes its item over TWO steps: forward pass, B4 of the step in between (the other warp's item), gradient pass,
// publish, B4 of its own step.  Both warps run this ONE copy of the code (the instruction cache is what binds the kernel
// otherwise); item t goes to slot t & 1, which PHY finished reading (item t + 2) before B4(t + 1).  The AUX warp keeps
// its own duty, the fixed-order sum of the control-gradient row PHY left in front of each B4.
// Latency regime only (PHYMB).  With many CTAs per SM the busiest scheduler binds and the kernel sits on an instruction-cache
// cliff: two producers cost 65 k samples 1.02 -> 1.33 ms, and so did this shared loop run with ONE producer (stride 1) --
// the extra 3 KB of code alone -- so those instantiations keep the roles' own backward sweeps (OWN_BWD above).
template <class Cfg>
SDE4_DEV void ws_densaux_bwd(const float* W, const float4* tiles, float* xch, const sde4_model_desc& d, const CostK& a,
                             int g0, int gmax, int par) {
  using A = typename Cfg::A;
  using MD = typename Cfg::MD;
  constexpr int NPART = Cfg::NPART, NZ = Cfg::NZ, NU = Cfg::NU;
  const unsigned G = (unsigned)a.G;
  const int H = a.H;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const float* wD = W + A::OFF_DENSITY;
  const float4* tD = tiles + Cfg::TILE_D * 32;
  // Gradient row of step ts: fixed-order sum over the CTA's particles of what PHY left in the exchange buffer:
  // thread = (control j, quarter of the particles); serial over the quarter, two butterfly rounds.
  constexpr int QP = NPART / 4;
  const int rj = lane >> 2, rq = lane & 3;
  auto sum_row = [&](int ts) {
    if (!a.warp_reduce) return;
    float v = 0.0f;
    if (rj < NU) {
#pragma unroll
      for (int k = 0; k < QP; ++k) v += xch[(Cfg::B_GU + rj) * NPART + rq * QP + k];
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (rj < NU && rq == 0) a.gpart[((size_t)ts * NU + rj) * a.PW + g0 / NPART] = v;
  };
  int nb = 0;  // B4 barriers joined so far; barrier number k is B4 of step H - 1 - k
  auto join = [&]() {
    ws_sync<Cfg::THREADS>();  // B4: PHY stored gradient row H - nb in front of it and rewrites the slot only after it
    if (par == 1 && nb >= 1) sum_row(H - nb);
    ++nb;
  };
  // rows of the A operand: particles g0 + g, g0 + g + 8 (clamped to the last valid sample)
  const unsigned ga = min(g0 + g, gmax), gb = min(g0 + g + 8, gmax);
  float az[4];
  auto load_inputs = [&](int ts) {
    const float* xin = a.xbuf + (size_t)ts * a.x_rows * G;
    const int k0 = nth_bit_rt<A::NIN_MASK>(t), k1 = nth_bit_rt<A::NIN_MASK>(t + 4 < NZ ? t + 4 : 0);
    az[0] = t < NZ ? xin[(size_t)k0 * G + ga] * d.inv_red_scale : 0.0f;
    az[1] = t < NZ ? xin[(size_t)k0 * G + gb] * d.inv_red_scale : 0.0f;
    az[2] = t + 4 < NZ ? xin[(size_t)k1 * G + ga] * d.inv_red_scale : 0.0f;
    az[3] = t + 4 < NZ ? xin[(size_t)k1 * G + gb] * d.inv_red_scale : 0.0f;
  };
  int ts = H - 1 - par;
  bool whole = par == 0;  // the item of step H - 1 is due at the first barrier: made in one piece
  if (ts >= 0) load_inputs(ts);  // the forward sweep's last B2 ordered PHY's checkpoint stores before these loads
  for (; ts >= 0; ts -= 2) {
    float out[4], J[4];
    MD::template value_grad<SDE4_WS_FAST_DENSB>(tD, wD, lane, az, out, J, [&]() {
      if (!whole) join();
      whole = false;
    });
    load_inputs(ts > 1 ? ts - 2 : 0);
    const int so = (ts & 1) * Cfg::B_DSLOT;
    if (t == 0) {
      xch[(Cfg::B_DENS + so) * NPART + g] = out[0];
      xch[(Cfg::B_DENS + so) * NPART + g + 8] = out[2];
    }
    tile_to_xch<NPART>(xch, Cfg::B_J + so, NZ, lane, J);
    join();
  }
  while (nb < H) join();       // (the last item belongs to the other warp)
  ws_sync<Cfg::THREADS>();     // PHY's last gradient row is there
  if (par == 1) sum_row(0);
}

// ---- RES role: one warp, 16 particles; residual force / moment networks on the tensor cores, in-kernel noise ----------
template <class Cfg, bool PHYMB>
SDE4_DEV void ws_role_res(const float* W, const float4* tiles, float* xch, const sde4_model_desc& d, const CostK& a,
                          const float* ftape, int g0, int gmax) {
  using A = typename Cfg::A;
  using MA = typename Cfg::MA;
  using MB = typename Cfg::MB;
  using NetA = typename A::NetA;
  using NetB = typename A::NetB;
  constexpr int NPART = Cfg::NPART;
  const unsigned G = (unsigned)a.G;
  const int H = a.H;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const float* wA = W + A::OFF_NET_A;
  const float* wB = W + A::OFF_NET_B;
  const float4* tA = tiles + Cfg::TILE_A * 32;
  const float4* tB = tiles + Cfg::TILE_B * 32;
  auto idxA = [](int k) { return nth_bit_rt<A::A_MASK>(k); };
  auto idxB = [](int k) { return nth_bit_rt<A::B_MASK>(k); };
  // ---------------- forward sweep ----------------
  ws_sync<Cfg::THREADS>();  // B0
  for (int ts = 0; ts < H; ++ts) {
    ws_sync3();  // B1
    float aA[4], Fn[4];
    a_from_xch<NPART>(xch, Cfg::X_IN6, NetA::IN, lane, idxA, aA);
    float dA1f[MA::NT1][4], dA2f[MA::NT2][4];
    MA::template fwd<false, SDE4_WS_FAST_FWD>(tA, wA, lane, aA, Fn, dA1f, dA2f);
    tile_to_xch<NPART>(xch, Cfg::X_FN, 3, lane, Fn);
    if constexpr (!PHYMB) {  // (PHYMB: the residual-moment network of the forward sweep runs on the PHY warp)
      float aB[4], Mn[4];
      float dB1f[MB::NT1][4], dB2f[MB::NT2][4];
      a_from_xch<NPART>(xch, Cfg::X_IN6, NetB::IN, lane, idxB, aB);
      MB::template fwd<false, SDE4_WS_FAST_FWD>(tB, wB, lane, aB, Mn, dB1f, dB2f);
      tile_to_xch<NPART>(xch, Cfg::X_MN, 3, lane, Mn);
    }
    ws_sync<Cfg::THREADS>();  // B2
  }
  // ---------------- backward sweep ----------------
  const unsigned ga = min(g0 + g, gmax), gb = min(g0 + g + 8, gmax);
  float aA[4], aB[4], Fn[4];
  float dA1[MA::NT1][4], dA2[MA::NT2][4], dB1[MB::NT1][4], dB2[MB::NT2][4];
  auto load_inputs = [&](int ts) {
    const float* ft = ftape + (size_t)ts * 6 * G;
    const int a0 = idxA(t), a1 = idxA(t + 4 < NetA::IN ? t + 4 : 0), c0 = idxB(t), c1 = idxB(t + 4 < NetB::IN ? t + 4 : 0);
    aA[0] = t < NetA::IN ? ft[(size_t)a0 * G + ga] : 0.0f;
    aA[1] = t < NetA::IN ? ft[(size_t)a0 * G + gb] : 0.0f;
    aA[2] = t + 4 < NetA::IN ? ft[(size_t)a1 * G + ga] : 0.0f;
    aA[3] = t + 4 < NetA::IN ? ft[(size_t)a1 * G + gb] : 0.0f;
    aB[0] = t < NetB::IN ? ft[(size_t)c0 * G + ga] : 0.0f;
    aB[1] = t < NetB::IN ? ft[(size_t)c0 * G + gb] : 0.0f;
    aB[2] = t + 4 < NetB::IN ? ft[(size_t)c1 * G + ga] : 0.0f;
    aB[3] = t + 4 < NetB::IN ? ft[(size_t)c1 * G + gb] : 0.0f;
  };
  load_inputs(H - 1);  // the forward sweep's last B2 ordered PHY's feature stores before these loads
  for (int ts = H - 1; ts >= 0; --ts) {
    float Mn_unused[4];
    MA::template fwd<true, SDE4_WS_FAST_RESB>(tA, wA, lane, aA, Fn, dA1, dA2);  // cotangent independent: ahead of the barrier
    MB::template fwd<true, SDE4_WS_FAST_RESB>(tB, wB, lane, aB, Mn_unused, dB1, dB2);
    load_inputs(ts > 0 ? ts - 1 : 0);  // requested one step ahead
    ws_sync_pr();  // B3: cotangents of the network outputs are there
    float agF[4], agM[4], giA[4], giB[4];
    a_from_xch<NPART>(xch, Cfg::B_GF, 3, lane, [](int k) { return k; }, agF);
    a_from_xch<NPART>(xch, Cfg::B_GM, 3, lane, [](int k) { return k; }, agM);
    MA::template bwd<SDE4_WS_FAST_RESB>(tA, lane, agF, dA1, dA2, giA);
    MB::template bwd<SDE4_WS_FAST_RESB>(tB, lane, agM, dB1, dB2, giB);
    tile_to_xch<NPART>(xch, Cfg::B_GIA, NetA::IN, lane, giA);
    tile_to_xch<NPART>(xch, Cfg::B_GIB, NetB::IN, lane, giB);
    tile_to_xch<NPART>(xch, Cfg::B_FN, 3, lane, Fn);
    ws_sync<Cfg::THREADS>();  // B4
  }
  ws_sync<Cfg::THREADS>();
}

// ---- AUX role: one warp; in-kernel threefry noise one step ahead (forward sweep), fixed-order sum of the control-gradient
//      rows over the CTA's particles (backward sweep) ------------------------------------------------------------------
template <class Cfg, bool OWN_BWD>
SDE4_DEV void ws_role_aux(float* xch, const sde4_model_desc& d, const CostK& a, int g0, bool lane_active, int n, int pd) {
  const int H = a.H;
  WsNoise<Cfg, 2> noise(xch, d, a, g0, lane_active, n, pd);
  noise.gen(0);
  ws_sync<Cfg::THREADS>();  // B0: dw_0 is there (PHY reads dw_t in front of B2 of step t)
  for (int ts = 0; ts < H; ++ts) {
    if (ts + 1 < H) noise.gen(ts + 1);  // slot (ts + 1) % 3: PHY reads it after B2 of the next step
    ws_sync<Cfg::THREADS>();  // B2
  }
  if constexpr (!OWN_BWD) return;
  constexpr int NPART = Cfg::NPART, NU = Cfg::NU;
  const int lane = threadIdx.x & 31;
  // Gradient row of step ts: fixed-order sum over the CTA's particles of what PHY left in the exchange buffer:
  // thread = (control j, quarter of the particles); serial over the quarter, two butterfly rounds.
  constexpr int QP = NPART / 4;
  const int rj = lane >> 2, rq = lane & 3;
  auto sum_row = [&](int ts) {
    if (!a.warp_reduce) return;
    float v = 0.0f;
    if (rj < NU) {
#pragma unroll
      for (int k = 0; k < QP; ++k) v += xch[(Cfg::B_GU + rj) * NPART + rq * QP + k];
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (rj < NU && rq == 0) a.gpart[((size_t)ts * NU + rj) * a.PW + g0 / NPART] = v;
  };
  for (int ts = H - 1; ts >= 0; --ts) {
    ws_sync<Cfg::THREADS>();  // B4 of step ts: PHY stored row ts + 1 before it and rewrites the slot only after it
    if (ts + 1 < H) sum_row(ts + 1);
  }
  ws_sync<Cfg::THREADS>();
  sum_row(0);
}

// (__launch_bounds__(128, 3): 153 registers.  Capping at 128 for four CTAs per SM helps at 16 k samples (282 -> 225 us, H = 20)
// and costs everywhere else: C3 0.358 -> 0.377 ms, 131 k samples 2.03 -> 2.24 ms.)
// One CTA = 16 particles, four warps: PHY, DENS, RES, AUX -- warp k of every CTA lands on scheduler k of its SM, so each
// scheduler (and its L0 instruction cache) runs the code of ONE role.
// (That locality is worth a factor of two: rotating the roles per CTA, so that every scheduler holds a mix of roles and no
// scheduler is left with the busiest role of every resident CTA, ran 1.2x (16 k samples) to 1.75x (131 k samples) SLOWER --
// 299 -> 354 us and 2.18 -> 3.80 ms, H = 20, scripts/ws_sizes.py -- three roles' loop bodies do not fit one L0 cache.)
// Requirements checked by the launcher: value + gradient, in-kernel threefry noise, no state constraints.
// PHYMB: where the residual-moment network of the FORWARD sweep runs.  With one wave of at most 3 CTAs per SM the launch is bound by
// the dependent chain of a step and the PHY warp, idle while the networks run, takes it (true); with many CTAs per SM
// every scheduler holds the same role of several CTAs, the launch is bound by the busiest ROLE, which is PHY, and the
// network stays with the RES role (false).
// MINB = 4 (128 registers, a few spills): only where a fourth resident CTA per SM saves a whole wave of CTAs -- 7 168
// samples (448 CTAs: one wave instead of 444 + 4) and 16 384 samples (two waves instead of three); the launcher decides.
template <class A, bool PLAINC, bool PHYMB, int MINB = 3>
__global__ void __launch_bounds__(128, MINB) k_cost_ws(const __grid_constant__ sde4_model_desc d,
                                                    const __grid_constant__ sde4_cost_desc cd,
                                                    const __grid_constant__ sde4_cost_desc td,
                                                    const __grid_constant__ CostK a, float* ftape) {
  using Cfg = WsCfg<A>;
  constexpr int NPART = Cfg::NPART;
  extern __shared__ __align__(16) float W[];
  __shared__ __align__(8) uint64_t mbar;
  asm volatile("griddepcontrol.launch_dependents;");  // the reduction kernel behind this launch may take its place now
  stage_params_bulk(W, a.params, A::PARAM_TOTAL * sizeof(float), &mbar);
  Diffusion<typename Cfg::M8>::derive(W);
  float4* tiles = reinterpret_cast<float4*>(W + Cfg::OFF_TILES);
  float* xch = W + Cfg::OFF_XCH;
  {  // fragment image, zeroed exchange buffer (dw of noise-free states stays 0), noise prior and noise-state list
    Cfg::MD::build(tiles + Cfg::TILE_D * 32, W + A::OFF_DENSITY, threadIdx.x, Cfg::THREADS);
    Cfg::MA::build(tiles + Cfg::TILE_A * 32, W + A::OFF_NET_A, threadIdx.x, Cfg::THREADS);
    Cfg::MB::build(tiles + Cfg::TILE_B * 32, W + A::OFF_NET_B, threadIdx.x, Cfg::THREADS);
    for (int i = threadIdx.x; i < Cfg::VALS * NPART; i += Cfg::THREADS) xch[i] = 0.0f;
    if (threadIdx.x == 0) {
      int* nzlist = reinterpret_cast<int*>(W + Cfg::OFF_MISC + 16);
      int c = 0;
      for (int i = 0; i < 13; ++i)
        if (d.noise_prior[i] != 0.0f) nzlist[c++] = i;
      for (; c < 16; ++c) nzlist[c] = 0;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g0 = blockIdx.x * NPART;                            // first sample of this CTA
  const int pl = lane & (NPART - 1);                             // PHY / RES-noise: particle of this lane
  const bool in_range = g0 + pl < a.G;
  const bool active = in_range && lane < NPART;                  // one lane per particle stores
  const int g = in_range ? g0 + pl : a.G - 1;                    // lanes beyond G shadow the last sample
  const int p = g / a.N, n = g - p * a.N;
  const int pd = p % a.data_mod;
  if (a.mask) {  // CTA-uniform exit (the roles meet at CTA-wide barriers)
    const int on = __ldg(a.mask + p) != 0;
    if (!__syncthreads_or(on)) return;
  }
  if (warp == 0) ws_role_phy<Cfg, PLAINC, PHYMB>(W, tiles, xch, d, cd, td, a, ftape, pl, g, p, pd, active);
  else if (warp == 2) ws_role_res<Cfg, PHYMB>(W, tiles, xch, d, a, ftape, g0, a.G - 1);
  else {  // forward sweep: their own roles; backward sweep (latency regime): the two density producers, one copy of the code
    if (warp == 1) ws_role_dens<Cfg, !PHYMB>(W, tiles, xch, d, a, g0, a.G - 1);
    else ws_role_aux<Cfg, !PHYMB>(xch, d, a, g0, in_range, n, pd);
    if constexpr (PHYMB) ws_densaux_bwd<Cfg>(W, tiles, xch, d, a, g0, a.G - 1, warp == 3 ? 1 : 0);
  }
}

}  // namespace sde4

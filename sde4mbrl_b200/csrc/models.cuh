// models.cuh -- the three ControlledSDE subclasses of the reference as per-thread device
// code: drift, distance-aware diffusion, projection, and their hand-derived
// vector-Jacobian products (the reverse-mode pass APG differentiates through).
//
//   MassSpringDamper  sde4mbrlExamples/mass_spring_damper/mass_spring_model.py:21-123
//   CartPoleSDE       sde4mbrlExamples/cartpole/cartpole_sde.py:20-106
//   SDERotorModel     sde4mbrlExamples/rotor_uav/sde_rotor_model.py:121-458,
//                     motor_models.py:21-95, rotor_uav/utils.py:13-63
//   diffusion         sde4mbrl/nsde.py:298-449
#pragma once
#include "nets.cuh"

namespace sde4 {

// ----------------------------------------------------------------------------
// Static architecture: everything that changes the *code* (sizes, activations,
// index sets, structural flags).  Numeric constants come from sde4_model_desc.
// ----------------------------------------------------------------------------
template <int MODEL_, int NX_, int NU_, uint32_t FLAGS_, class DNet_, uint32_t NIN_MASK_, uint32_t NOUT_MASK_,
          bool HAS_SCALER_, class NetA_, uint32_t A_MASK_, class NetB_, uint32_t B_MASK_>
struct Arch {
  static constexpr int MODEL = MODEL_, NX = NX_, NU = NU_;
  static constexpr uint32_t FLAGS = FLAGS_, NIN_MASK = NIN_MASK_, NOUT_MASK = NOUT_MASK_, A_MASK = A_MASK_,
                            B_MASK = B_MASK_;
  using DNet = DNet_;
  using NetA = NetA_;
  using NetB = NetB_;
  static constexpr bool HAS_DENSITY = DNet::PRESENT;
  static constexpr bool HAS_SCALER = HAS_SCALER_ && HAS_DENSITY;
  static constexpr int NO = popc(NOUT_MASK_);  // number of eta-scaled states
  static constexpr int N_SCALER = HAS_SCALER ? 2 * NO : 0;
  static constexpr int N_SCALARS = MODEL_ == SDE4_MODEL_ROTOR ? 16 : 0;
  // packed parameter block
  static constexpr int OFF_SCALARS = 0;
  static constexpr int OFF_SCALER = OFF_SCALARS + N_SCALARS;
  static constexpr int OFF_DENSITY = OFF_SCALER + up4(N_SCALER);
  static constexpr int OFF_NET_A = OFF_DENSITY + DNet::SIZE;
  static constexpr int OFF_NET_B = OFF_NET_A + NetA::SIZE;
  static constexpr int PARAM_TOTAL = up4(OFF_NET_B + NetB::SIZE);
  // derived per-block constants appended after the parameter image in shared memory
  static constexpr int OFF_DERIVED = PARAM_TOTAL;
  static constexpr int N_DERIVED = up4(2 * NO);
  static constexpr int SMEM_FLOATS = PARAM_TOTAL + N_DERIVED;
  static constexpr int MACS = DNet::MACS + NetA::MACS + NetB::MACS;
  static constexpr int smem_floats(int /*L*/) { return SMEM_FLOATS; }
  template <int SLOT, int L>
  SDE4_DEV static NetRef ref(const float* W) {
    constexpr int off = SLOT == 0 ? OFF_DENSITY : (SLOT == 1 ? OFF_NET_A : OFF_NET_B);
    return NetRef{SLOT, W + off};
  }

  static bool matches(const sde4_model_desc& d) {
    return d.model_id == MODEL && d.n_x == NX && d.n_u == NU && (uint32_t)d.flags == FLAGS &&
           DNet::matches(d.density) && NetA::matches(d.net_a) && NetB::matches(d.net_b) &&
           (!HAS_DENSITY || (d.noise_in_mask == NIN_MASK && d.noise_out_mask == NOUT_MASK)) &&
           d.n_scaler == N_SCALER && (!NetA::PRESENT || d.net_a_in_mask == A_MASK) &&
           (!NetB::PRESENT || d.net_b_in_mask == B_MASK);
  }
};

// Lazily scaled cotangent: vf[i] = s * v[i] (avoids materialising dt*lambda in registers).
template <int N>
struct Scaled {
  const float (&v)[N];
  float s;
  SDE4_DEV float operator[](int i) const { return s * v[i]; }
};

// ----------------------------------------------------------------------------
// quaternion algebra, [w, x, y, z]  (rotor_uav/utils.py:13-63)
// ----------------------------------------------------------------------------
struct Quat {
  float w, x, y, z;
};
SDE4_DEV Quat qmul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}
SDE4_DEV Quat qconj(const Quat& a) { return Quat{a.w, -a.x, -a.y, -a.z}; }
SDE4_DEV Quat qvec(float x, float y, float z) { return Quat{0.0f, x, y, z}; }
SDE4_DEV void qacc(Quat& a, const Quat& b) {
  a.w += b.w;
  a.x += b.x;
  a.y += b.y;
  a.z += b.z;
}
// For r = a (x) b (bilinear):  abar = rbar (x) conj(b),  bbar = conj(a) (x) rbar.
// Products with a pure-vector operand [0, v] (no multiplications by the literal zero, which the
// compiler may not fold under IEEE semantics).
SDE4_DEV Quat qmul_vq(float vx, float vy, float vz, const Quat& b) {  // [0,v] (x) b
  Quat r;
  r.w = -vx * b.x - vy * b.y - vz * b.z;
  r.x = vx * b.w + vy * b.z - vz * b.y;
  r.y = -vx * b.z + vy * b.w + vz * b.x;
  r.z = vx * b.y - vy * b.x + vz * b.w;
  return r;
}
SDE4_DEV Quat qmul_qv(const Quat& a, float vx, float vy, float vz) {  // a (x) [0,v]
  Quat r;
  r.w = -a.x * vx - a.y * vy - a.z * vz;
  r.x = a.w * vx + a.y * vz - a.z * vy;
  r.y = a.w * vy - a.x * vz + a.z * vx;
  r.z = a.w * vz + a.x * vy - a.y * vx;
  return r;
}

// ----------------------------------------------------------------------------
// Diffusion (nsde.py:298-449), shared by all models through Model::reduced_state.
// Derived constants D[0..NO) = exp(scaler[k])*1e-7 (coeff), D[NO..2NO) = offset.
// ----------------------------------------------------------------------------
template <class M>
struct Diffusion {
  using A = typename M::A;
  using DNet = typename A::DNet;
  using Ev = typename M::Ev;
  using Ctx = typename Ev::Ctx;
  static constexpr int NX = A::NX, NO = A::NO, NU = A::NU;
  // control_dependent_noise (nsde.py:362): the density network reads concat(reduced_state[indx_noise_in], u), i.e. its
  // input is wider than the noise-input mask by the number of controls
  static constexpr int NZ = DNet::IN > 0 ? popc(A::NIN_MASK) : 0;
  static constexpr bool CDN = DNet::IN > NZ;
  static_assert(DNet::IN == NZ || DNet::IN == NZ + NU, "density input = noise inputs (+ all controls)");

  SDE4_DEV static void gather_in(const float (&x)[NX], const float (&u)[NU], const sde4_model_desc& d,
                                 float (&z)[DNet::IN > 0 ? DNet::IN : 1]) {
    float red[M::NR];
    M::reduced_state(x, d, red);
#pragma unroll
    for (int k = 0; k < NZ; ++k) z[k] = red[nth_bit(A::NIN_MASK, k)];
    if constexpr (CDN) {
#pragma unroll
      for (int j = 0; j < NU; ++j) z[NZ + j] = u[j];
    }
  }

  SDE4_DEV static float eta_of(const float* W, const sde4_model_desc& d, float dens, int k) {
    float coeff = 0.0f, offs = 0.0f;
    if constexpr (A::HAS_SCALER) {
      coeff = W[A::OFF_DERIVED + k];
      offs = W[A::OFF_DERIVED + NO + k];
    }
    return sigmoidf_((d.aggr + coeff) * dens - 7.0f + offs);  // nsde.py:426
  }

  // sig[i] = eta_full[i] * prior[i]
  SDE4_DEV static void fwd(const float* W, Ctx& sc, const sde4_model_desc& d, const float (&x)[NX],
                           const float (&u)[NU], float (&sig)[NX]) {
#pragma unroll
    for (int i = 0; i < NX; ++i) sig[i] = d.noise_prior[i];  // non-noise outputs: eta = 1 (nsde.py:363)
    if constexpr (A::HAS_DENSITY) {
      float z[DNet::IN > 0 ? DNet::IN : 1];
      gather_in(x, u, d, z);
      float dens[1];
      Ev::template fwd<DNet, false>(A::template ref<0, Ev::L>(W), sc, z, dens);
#pragma unroll
      for (int k = 0; k < NO; ++k) sig[nth_bit(A::NOUT_MASK, k)] *= eta_of(W, d, dens[0], k);
    }
  }

  // Reverse mode of sig wrt x (and u, control_dependent_noise):  gx += (d sig / d x)^T (dw o lam), gu likewise, with
  // the noise read on the fly.
  // The density gradient J = d dens / d x does not depend on the cotangent, so it is formed
  // first (fused value+gradient pass) and scaled afterwards.
  template <class NoiseAt>
  SDE4_DEV static void vjp(const float* W, Ctx& sc, const sde4_model_desc& d, const float (&x)[NX],
                           const float (&u)[NU], const float (&lam)[NX], NoiseAt noise_at, float (&gx)[NX],
                           float (&gu)[NU]) {
    if constexpr (!A::HAS_DENSITY) {
      return;
    } else {
      float z[DNet::IN], J[DNet::IN], dens;
      gather_in(x, u, d, z);
      // cotangent of the density output: gd = sum_k (dw_i lam_i prior_i) eta'(.) (aggr + coeff_k)
      auto gd_of = [&](float dv) {
        float gd = 0.0f;
#pragma unroll
        for (int k = 0; k < NO; ++k) {
          const int i = nth_bit(A::NOUT_MASK, k);
          float coeff = 0.0f;
          if constexpr (A::HAS_SCALER) coeff = W[A::OFF_DERIVED + k];
          const float e = eta_of(W, d, dv, k);
          // states with zero prior have no noise sample (never generated / stored): guard the read
          const float dwi = d.noise_prior[i] != 0.0f ? noise_at(i) : 0.0f;
          const float t1 = dwi * lam[i] * d.noise_prior[i] * (e * (1.0f - e));  // d loss / d (sigmoid arg)
          gd = fmaf(t1, d.aggr + coeff, gd);
          if constexpr (Ev::PGRAD && A::HAS_SCALER) {  // scaler = [coeff | offset] = exp(.)*1e-7 (nsde.py:413-414)
            sc.dnet[0][(size_t)(DNet::DUMP_ROWS + k) * sc.dstride + sc.dcol] = t1 * dv * coeff;
            sc.dnet[0][(size_t)(DNet::DUMP_ROWS + NO + k) * sc.dstride + sc.dcol] = t1 * W[A::OFF_DERIVED + NO + k];
          }
        }
        return gd;
      };
      Ev::template scalar_value_grad<DNet>(A::template ref<0, Ev::L>(W), sc, z, dens, J, gd_of);
      float gred[M::NR];
#pragma unroll
      for (int r = 0; r < M::NR; ++r) gred[r] = 0.0f;
#pragma unroll
      for (int k = 0; k < NZ; ++k) gred[nth_bit(A::NIN_MASK, k)] = J[k];
      M::reduced_state_vjp(x, d, gred, gx);
      if constexpr (CDN) {
#pragma unroll
        for (int j = 0; j < NU; ++j) gu[j] += J[NZ + j];
      }
    }
  }

  // The same VJP in two halves for the software-pipelined backward sweep of the straight-line lane-split kernels:
  // `prep` is everything that does not depend on the cotangent -- the density value and its UNSCALED input gradient
  // J = d dens / d z (the longest dependency chain of a backward step) -- and is evaluated for step t-1 while step t
  // is applied; `apply` scales J by the density cotangent gd(lam, dw, dens) and pushes it through reduced_state.
  struct Tape {
    float dens, J[DNet::IN > 0 ? DNet::IN : 1];
  };
  SDE4_DEV static void prep(const float* W, Ctx& sc, const sde4_model_desc& d, const float (&x)[NX],
                            const float (&u)[NU], Tape& tp) {
    if constexpr (A::HAS_DENSITY) {
      float z[DNet::IN];
      gather_in(x, u, d, z);
      Ev::template scalar_value_grad<DNet>(A::template ref<0, Ev::L>(W), sc, z, tp.dens, tp.J, [](float) { return 1.0f; });
    }
  }
  template <class NoiseAt>
  SDE4_DEV static void apply(const float* W, const sde4_model_desc& d, const float (&x)[NX], const float (&lam)[NX],
                             NoiseAt noise_at, const Tape& tp, float (&gx)[NX], float (&gu)[NU]) {
    if constexpr (A::HAS_DENSITY) {
      float gd = 0.0f;
#pragma unroll
      for (int k = 0; k < NO; ++k) {
        const int i = nth_bit(A::NOUT_MASK, k);
        float coeff = 0.0f;
        if constexpr (A::HAS_SCALER) coeff = W[A::OFF_DERIVED + k];
        const float e = eta_of(W, d, tp.dens, k);
        const float dwi = d.noise_prior[i] != 0.0f ? noise_at(i) : 0.0f;
        gd = fmaf(dwi * lam[i] * d.noise_prior[i] * (e * (1.0f - e)), d.aggr + coeff, gd);
      }
      float gred[M::NR];
#pragma unroll
      for (int r = 0; r < M::NR; ++r) gred[r] = 0.0f;
#pragma unroll
      for (int k = 0; k < NZ; ++k) gred[nth_bit(A::NIN_MASK, k)] = tp.J[k] * gd;
      M::reduced_state_vjp(x, d, gred, gx);
      if constexpr (CDN) {
#pragma unroll
        for (int j = 0; j < NU; ++j) gu[j] = fmaf(tp.J[NZ + j], gd, gu[j]);
      }
    }
  }

  // prologue: threads compute exp(scaler)*1e-7 once per block (nsde.py:413)
  SDE4_DEV static void derive(float* __restrict__ W) {
    if constexpr (A::HAS_SCALER) {
      for (int k = threadIdx.x; k < 2 * NO; k += blockDim.x) W[A::OFF_DERIVED + k] = expf(W[A::OFF_SCALER + k]) * 1e-7f;
    }
  }
};

// ----------------------------------------------------------------------------
// Mass-spring-damper  (mass_spring_model.py:36-86)
// ----------------------------------------------------------------------------
template <class A_, class Ev_ = Ev1>
struct MsdModel {
  using A = A_;
  using Ev = Ev_;
  using Ctx = typename Ev_::Ctx;
  static constexpr int NX = 2, NU = 1, NR = 2, NAUX = 0, NPHYS = 0;
  static constexpr bool GT = (A::FLAGS & SDE4_MSD_GROUND_TRUTH) != 0;
  static constexpr bool SI = (A::FLAGS & SDE4_MSD_SIDE_INFO) != 0;
  static constexpr bool CTRL = (A::FLAGS & SDE4_MSD_CONTROL) != 0;
  using NetA = typename A::NetA;

  SDE4_DEV static void reduced_state(const float (&x)[NX], const sde4_model_desc&, float (&r)[NR]) {
    r[0] = x[0];
    r[1] = x[1];
  }
  SDE4_DEV static void reduced_state_vjp(const float (&)[NX], const sde4_model_desc&, const float (&g)[NR],
                                         float (&gx)[NX]) {
    gx[0] += g[0];
    gx[1] += g[1];
  }

  SDE4_DEV static void drift(const float* W, Ctx& sc, const sde4_model_desc& d, const float (&x)[NX],
                             const float (&u)[NU], float (&f)[NX]) {
    const float m = d.cfloat[0], k = d.cfloat[1], b = d.cfloat[2];
    const float u0 = CTRL ? u[0] : 0.0f;
    f[0] = x[1];
    if constexpr (GT) {
      f[1] = (u0 - k * x[0] - b * x[1]) / m;
    } else if constexpr (SI) {
      float in[1] = {x[1]}, out[1];
      Ev::template fwd<NetA, false>(A::template ref<1, Ev::L>(W), sc, in, out);
      f[1] = (u0 - k * x[0] + out[0]) / m;
    } else {
      float in[2] = {x[0], x[1]}, out[1];
      Ev::template fwd<NetA, false>(A::template ref<1, Ev::L>(W), sc, in, out);
      f[1] = out[0];
    }
  }

  SDE4_DEV static void drift_vjp(const float* W, Ctx& sc, const sde4_model_desc& d, const float (&x)[NX],
                                 const float (&u)[NU], const Scaled<NX>& vf, float (&gx)[NX], float (&gu)[NU]) {
    const float m = d.cfloat[0], k = d.cfloat[1], b = d.cfloat[2];
    gx[1] += vf[0];
    if constexpr (GT) {
      const float s = vf[1] / m;
      gx[0] -= k * s;
      gx[1] -= b * s;
      if constexpr (CTRL) gu[0] += s;
    } else if constexpr (SI) {
      const float s = vf[1] / m;
      gx[0] -= k * s;
      if constexpr (CTRL) gu[0] += s;
      float in[1] = {x[1]}, go[1] = {s}, gi[1];
      Ev::template fwd_vjp<NetA>(A::template ref<1, Ev::L>(W), sc, in, go, gi);
      gx[1] += gi[0];
    } else {
      float in[2] = {x[0], x[1]}, go[1] = {vf[1]}, gi[2];
      Ev::template fwd_vjp<NetA>(A::template ref<1, Ev::L>(W), sc, in, go, gi);
      gx[0] += gi[0];
      gx[1] += gi[1];
    }
  }

  SDE4_DEV static void project(float (&)[NX], float&) {}
  SDE4_DEV static void project_vjp(const float (&)[NX], float, float (&)[NX]) {}
};

// ----------------------------------------------------------------------------
// Cartpole  (cartpole_sde.py:56-83); state [x, xdot, theta, thetadot]
// ----------------------------------------------------------------------------
template <class A_, class Ev_ = Ev1>
struct CartpoleModel {
  using A = A_;
  using Ev = Ev_;
  using Ctx = typename Ev_::Ctx;
  static constexpr int NX = 4, NU = 1, NR = 4, NAUX = 0, NPHYS = 0;
  static constexpr bool SI = (A::FLAGS & SDE4_CART_SIDE_INFO) != 0;
  using NetA = typename A::NetA;
  using NetB = typename A::NetB;

  // [xdot, sin th, cos th, thdot] / state_scaling[-1]   (cartpole_sde.py:41)
  SDE4_DEV static void reduced_state(const float (&x)[NX], const sde4_model_desc& d, float (&r)[NR]) {
    float s, c;
    sincosf(x[2], &s, &c);
    r[0] = x[1] * d.inv_red_scale;
    r[1] = s * d.inv_red_scale;
    r[2] = c * d.inv_red_scale;
    r[3] = x[3] * d.inv_red_scale;
  }
  SDE4_DEV static void reduced_state_vjp(const float (&x)[NX], const sde4_model_desc& d, const float (&g)[NR],
                                         float (&gx)[NX]) {
    float s, c;
    sincosf(x[2], &s, &c);
    gx[1] += g[0] * d.inv_red_scale;
    gx[2] += (g[1] * c - g[2] * s) * d.inv_red_scale;
    gx[3] += g[3] * d.inv_red_scale;
  }

  SDE4_DEV static void drift(const float* W, Ctx& sc, const sde4_model_desc& d, const float (&x)[NX],
                             const float (&u)[NU], float (&f)[NX]) {
    float s, c;
    sincosf(x[2], &s, &c);
    const float thd_sc = x[3] * d.inv_state_scaling[3];
    const float sc0 = d.cfloat[0], sc1 = d.cfloat[1];
    f[0] = x[1];
    f[2] = x[3];
    if constexpr (!SI) {
      float in[4] = {s, c, thd_sc, u[0]}, out[2];
      Ev::template fwd<NetA, false>(A::template ref<1, Ev::L>(W), sc, in, out);
      f[1] = sc0 * out[0];
      f[3] = sc1 * out[1];
    } else {
      float in[3] = {s, c, thd_sc}, out[2];
      Ev::template fwd<NetA, false>(A::template ref<1, Ev::L>(W), sc, in, out);
      float cin[2] = {c, c * c}, ce[2];
      Ev::template fwd<NetB, false>(A::template ref<2, Ev::L>(W), sc, cin, ce);
      f[1] = sc0 * (out[0] + ce[0] * u[0]);
      f[3] = sc1 * (out[1] + ce[1] * u[0]);
    }
  }

  SDE4_DEV static void drift_vjp(const float* W, Ctx& sc, const sde4_model_desc& d, const float (&x)[NX],
                                 const float (&u)[NU], const Scaled<NX>& vf, float (&gx)[NX], float (&gu)[NU]) {
    float s, c;
    sincosf(x[2], &s, &c);
    const float thd_sc = x[3] * d.inv_state_scaling[3];
    const float sc0 = d.cfloat[0], sc1 = d.cfloat[1];
    gx[1] += vf[0];
    gx[3] += vf[2];
    float gs = 0.0f, gc = 0.0f;  // cotangents of sin(theta), cos(theta)
    if constexpr (!SI) {
      float in[4] = {s, c, thd_sc, u[0]}, go[2] = {sc0 * vf[1], sc1 * vf[3]}, gi[4];
      Ev::template fwd_vjp<NetA>(A::template ref<1, Ev::L>(W), sc, in, go, gi);
      gs = gi[0];
      gc = gi[1];
      gx[3] += gi[2] * d.inv_state_scaling[3];
      gu[0] += gi[3];
    } else {
      const float go0 = sc0 * vf[1], go1 = sc1 * vf[3];
      {
        float in[3] = {s, c, thd_sc}, go[2] = {go0, go1}, gi[3];
        Ev::template fwd_vjp<NetA>(A::template ref<1, Ev::L>(W), sc, in, go, gi);
        gs = gi[0];
        gc = gi[1];
        gx[3] += gi[2] * d.inv_state_scaling[3];
      }
      {
        float cin[2] = {c, c * c}, ce[2];
        Ev::template fwd<NetB, true>(A::template ref<2, Ev::L>(W), sc, cin, ce);
        gu[0] += go0 * ce[0] + go1 * ce[1];
        float gce[2] = {go0 * u[0], go1 * u[0]}, gcin[2];
        Ev::template bwd<NetB>(A::template ref<2, Ev::L>(W), sc, gce, gcin);
        gc += gcin[0] + 2.0f * c * gcin[1];
      }
    }
    gx[2] += gs * c - gc * s;
  }

  SDE4_DEV static void project(float (&)[NX], float&) {}
  SDE4_DEV static void project_vjp(const float (&)[NX], float, float (&)[NX]) {}
};

// ----------------------------------------------------------------------------
// Multirotor  (sde_rotor_model.py:121-458); state [p(3), v(3), q(wxyz), w(3)]
// Scalars in the packed block (A::OFF_SCALARS + k):
//   0 mass 1 Ixx 2 Iyy 3 Izz 4 Lmx 5 Lmy 6 CoeffDrag 7..10 motor poly c3,c2,c1,c0
//   11 aero_kdx 12 aero_kdy 13 aero_kdz 14 aero_kh
// ----------------------------------------------------------------------------
template <int NU>
struct Mixer;
template <>
struct Mixer<4> {  // motor_models.py:9-16
  SDE4_DEV static float m(int r, int i) {
    constexpr float v[4][4] = {{-0.707f, 0.707f, 0.707f, -0.707f},
                               {-0.707f, 0.707f, -0.707f, 0.707f},
                               {-1.0f, -1.0f, 1.0f, 1.0f},
                               {1.0f, 1.0f, 1.0f, 1.0f}};
    return v[r][i];
  }
};
template <>
struct Mixer<6> {  // motor_models.py:21-28
  SDE4_DEV static float m(int r, int i) {
    constexpr float v[4][6] = {{-1.0f, 1.0f, 0.5f, -0.5f, -0.5f, 0.5f},
                               {0.0f, 0.0f, -0.866f, 0.866f, -0.866f, 0.866f},
                               {1.0f, -1.0f, 1.0f, -1.0f, -1.0f, 1.0f},
                               {1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f}};
    return v[r][i];
  }
};

template <class A_, class Ev_ = Ev1>
struct RotorModel {
  using A = A_;
  using Ev = Ev_;
  using Ctx = typename Ev_::Ctx;
  static constexpr int NX = 13, NU = A::NU, NR = 13, NAUX = 1;
  static constexpr int NPHYS = 16;  // physical scalars of the packed block whose gradients the training kernel forms
  static constexpr bool DRAG = (A::FLAGS & SDE4_ROTOR_AERO_DRAG) != 0;
  using NetA = typename A::NetA;  // res_forces
  using NetB = typename A::NetB;  // res_moments
  static constexpr bool HAS_F = NetA::PRESENT, HAS_M = NetB::PRESENT;

  SDE4_DEV static void reduced_state(const float (&x)[NX], const sde4_model_desc& d, float (&r)[NR]) {
#pragma unroll
    for (int i = 0; i < NX; ++i) r[i] = x[i] * d.inv_red_scale;  // sde_rotor_model.py:139-144
  }
  SDE4_DEV static void reduced_state_vjp(const float (&)[NX], const sde4_model_desc& d, const float (&g)[NR],
                                         float (&gx)[NX]) {
#pragma unroll
    for (int i = 0; i < NX; ++i) gx[i] = fmaf(g[i], d.inv_red_scale, gx[i]);
  }

  // (v_b, w) / state_scaling[[3,4,5,10,11,12]]   (sde_rotor_model.py:311-315)
  SDE4_DEV static void nn_features(const float (&x)[NX], const sde4_model_desc& d, const float (&vb)[3],
                                   float (&in6)[6]) {
    in6[0] = vb[0] * d.inv_state_scaling[3];
    in6[1] = vb[1] * d.inv_state_scaling[4];
    in6[2] = vb[2] * d.inv_state_scaling[5];
    in6[3] = x[10] * d.inv_state_scaling[10];
    in6[4] = x[11] * d.inv_state_scaling[11];
    in6[5] = x[12] * d.inv_state_scaling[12];
  }

  SDE4_DEV static float motor_thrust(const float* __restrict__ S, float u) {
    return fmaf(fmaf(fmaf(S[7], u, S[8]), u, S[9]), u, S[10]);  // motor_models.py:47-60
  }
  SDE4_DEV static float motor_thrust_du(const float* __restrict__ S, float u) {
    return fmaf(fmaf(3.0f * S[7], u, 2.0f * S[8]), u, S[9]);
  }

  SDE4_DEV static void drift(const float* W, Ctx& sc, const sde4_model_desc& d, const float (&x)[NX],
                             const float (&u)[NU], float (&f)[NX]) {
    const float* __restrict__ S = W + A::OFF_SCALARS;
    const Quat q{x[6], x[7], x[8], x[9]};
    // v_b = q^-1 (x) (v (x) q)   (rotor_uav/utils.py:53-63, :439)
    const Quat t1 = qmul_vq(x[3], x[4], x[5], q);
    const Quat r1 = qmul(qconj(q), t1);
    const float vb[3] = {r1.x, r1.y, r1.z};
    float in6[6];
    nn_features(x, d, vb, in6);
    float Fb[3] = {0.0f, 0.0f, 0.0f}, Mt[3] = {0.0f, 0.0f, 0.0f};
    if constexpr (HAS_F) {
      float in[NetA::IN > 0 ? NetA::IN : 1], out[3];
#pragma unroll
      for (int k = 0; k < NetA::IN; ++k) in[k] = in6[nth_bit(A::A_MASK, k)];
      Ev::template fwd<NetA, false>(A::template ref<1, Ev::L>(W), sc, in, out);
      Fb[0] = out[0];
      Fb[1] = out[1];
      Fb[2] = out[2];
    }
    if constexpr (DRAG) {  // :264-277
      Fb[0] += -S[11] * vb[0];
      Fb[1] += -S[12] * vb[1];
      Fb[2] += -S[13] * vb[2] + S[14] * (vb[0] * vb[0] + vb[1] * vb[1]);
    }
    if constexpr (HAS_M) {
      float in[NetB::IN > 0 ? NetB::IN : 1], out[3];
#pragma unroll
      for (int k = 0; k < NetB::IN; ++k) in[k] = in6[nth_bit(A::B_MASK, k)];
      Ev::template fwd<NetB, false>(A::template ref<2, Ev::L>(W), sc, in, out);
      Mt[0] = out[0];
      Mt[1] = out[1];
      Mt[2] = out[2];
    }
    // motors -> thrust and moments (motor_models.py:74-95)
    float T = 0.0f, Mm[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const float th = motor_thrust(S, u[i]);
      Mm[0] = fmaf(Mixer<NU>::m(0, i) * S[4], th, Mm[0]);
      Mm[1] = fmaf(Mixer<NU>::m(1, i) * S[5], th, Mm[1]);
      Mm[2] = fmaf(Mixer<NU>::m(2, i) * S[6], th, Mm[2]);
      T = fmaf(Mixer<NU>::m(3, i), th, T);
    }
    // translational dynamics (:173-210): a = rot(q, [0,0,T] + Fres)/m + [0,0,-g]
    Fb[2] += T;
    const Quat t2 = qmul_qv(q, Fb[0], Fb[1], Fb[2]);
    const Quat r2 = qmul(t2, qconj(q));
    const float inv_m = fast_rcp(S[0]), g = d.cfloat[0];  // F/m as F*(1/m): <= 2 ulp from the division
    f[0] = x[3];
    f[1] = x[4];
    f[2] = x[5];
    f[3] = r2.x * inv_m;
    f[4] = r2.y * inv_m;
    f[5] = fmaf(r2.z, inv_m, -g);
    // rotational dynamics (:212-250)
    const float I0 = S[1], I1 = S[2], I2 = S[3];
    const float w0 = x[10], w1 = x[11], w2 = x[12];
    const float Iw0 = I0 * w0, Iw1 = I1 * w1, Iw2 = I2 * w2;
    f[10] = (Mm[0] + Mt[0] - (w1 * Iw2 - w2 * Iw1)) * fast_rcp(I0);
    f[11] = (Mm[1] + Mt[1] - (w2 * Iw0 - w0 * Iw2)) * fast_rcp(I1);
    f[12] = (Mm[2] + Mt[2] - (w0 * Iw1 - w1 * Iw0)) * fast_rcp(I2);
    const Quat qd = qmul_qv(q, w0, w1, w2);
    f[6] = 0.5f * qd.w;
    f[7] = 0.5f * qd.x;
    f[8] = 0.5f * qd.y;
    f[9] = 0.5f * qd.z;
  }

  // Cotangent-independent half of drift_vjp for the warp-specialised kernels (kernels_ws.cuh): the forward passes of
  // the two residual networks at x_t.  What the cotangent-dependent half needs of them is the residual force itself
  // and act' of the hidden units this lane owns.
  static constexpr int TB1A = HAS_F ? Ev_::L > 1 ? (NetA::H1 + Ev_::L - 1) / Ev_::L : NetA::H1 : 1;
  static constexpr int TB2A = HAS_F ? Ev_::L > 1 ? (NetA::H2 + Ev_::L - 1) / Ev_::L : NetA::H2 : 1;
  static constexpr int TB1B = HAS_M ? Ev_::L > 1 ? (NetB::H1 + Ev_::L - 1) / Ev_::L : NetB::H1 : 1;
  static constexpr int TB2B = HAS_M ? Ev_::L > 1 ? (NetB::H2 + Ev_::L - 1) / Ev_::L : NetB::H2 : 1;
  struct DriftTape {
    float Fb[3];                  // residual force (network output only)
    float dA1[TB1A], dA2[TB2A];   // act' of the own hidden units, res_forces
    float dB1[TB1B], dB2[TB2B];   // res_moments
  };
  SDE4_DEV static void drift_prep(const float* W, Ctx& sc, const sde4_model_desc& d, const float (&x)[NX],
                                  DriftTape& tp) {
    static_assert(Ev_::L > 1, "tape split is a lane-split evaluator feature");
    const Quat q{x[6], x[7], x[8], x[9]};
    const Quat t1 = qmul_vq(x[3], x[4], x[5], q);
    const Quat r1 = qmul(qconj(q), t1);
    const float vb[3] = {r1.x, r1.y, r1.z};
    float in6[6];
    nn_features(x, d, vb, in6);
    tp.Fb[0] = tp.Fb[1] = tp.Fb[2] = 0.0f;
    if constexpr (HAS_F) {
      float in[NetA::IN > 0 ? NetA::IN : 1];
#pragma unroll
      for (int k = 0; k < NetA::IN; ++k) in[k] = in6[nth_bit(A::A_MASK, k)];
      Ev::template fwd<NetA, true>(A::template ref<1, Ev::L>(W), sc, in, tp.Fb);
#pragma unroll
      for (int m = 0; m < TB1A; ++m) tp.dA1[m] = sc.d1[m];
#pragma unroll
      for (int m = 0; m < TB2A; ++m) tp.dA2[m] = sc.d2[m];
    }
    if constexpr (HAS_M) {
      float in[NetB::IN > 0 ? NetB::IN : 1], out[3];
#pragma unroll
      for (int k = 0; k < NetB::IN; ++k) in[k] = in6[nth_bit(A::B_MASK, k)];
      Ev::template fwd<NetB, true>(A::template ref<2, Ev::L>(W), sc, in, out);
#pragma unroll
      for (int m = 0; m < TB1B; ++m) tp.dB1[m] = sc.d1[m];
#pragma unroll
      for (int m = 0; m < TB2B; ++m) tp.dB2[m] = sc.d2[m];
    }
  }
  // cotangent-dependent half: drift_vjp with the network forward passes replaced by the tape
  SDE4_DEV static void drift_apply(const float* W, Ctx& sc, const sde4_model_desc& d, const float (&x)[NX],
                                   const float (&u)[NU], const Scaled<NX>& vf, const DriftTape& tp, float (&gx)[NX],
                                   float (&gu)[NU]) {
    drift_vjp_impl<true>(W, sc, d, x, u, vf, &tp, gx, gu);
  }
  SDE4_DEV static void drift_vjp(const float* W, Ctx& sc, const sde4_model_desc& d, const float (&x)[NX],
                                 const float (&u)[NU], const Scaled<NX>& vf, float (&gx)[NX], float (&gu)[NU]) {
    drift_vjp_impl<false>(W, sc, d, x, u, vf, nullptr, gx, gu);
  }

  template <bool FROM_TAPE>
  SDE4_DEV static void drift_vjp_impl(const float* W, Ctx& sc, const sde4_model_desc& d, const float (&x)[NX],
                                      const float (&u)[NU], const Scaled<NX>& vf, const DriftTape* tp, float (&gx)[NX],
                                      float (&gu)[NU]) {
    static_assert(!(FROM_TAPE && Ev::PGRAD), "the training kernels evaluate the networks in place");
    const float* __restrict__ S = W + A::OFF_SCALARS;
    const Quat q{x[6], x[7], x[8], x[9]};
    const Quat qc = qconj(q);
    const Quat V = qvec(x[3], x[4], x[5]);
    const Quat t1 = qmul_vq(x[3], x[4], x[5], q);
    const Quat r1 = qmul(qc, t1);
    const float vb[3] = {r1.x, r1.y, r1.z};
    float in6[6];
    nn_features(x, d, vb, in6);
    const float m = S[0];
    const float I0 = S[1], I1 = S[2], I2 = S[3];
    const float w0 = x[10], w1 = x[11], w2 = x[12];
    Quat gq{0.0f, 0.0f, 0.0f, 0.0f};
    float gw[3] = {0.0f, 0.0f, 0.0f};
    // pdot = v
    gx[3] += vf[0];
    gx[4] += vf[1];
    gx[5] += vf[2];
    // ---- rotational part: alpha = (Mm + Mres - w x (I w)) / I
    const float gM[3] = {vf[10] * fast_rcp(I0), vf[11] * fast_rcp(I1), vf[12] * fast_rcp(I2)};
    {
      // c = w x b, b = I*w ; Mtot -= c  => cbar = -gM ; wbar += b x cbar ; bbar = cbar x w ; wbar += I*bbar
      const float b0 = I0 * w0, b1 = I1 * w1, b2 = I2 * w2;
      const float c0 = -gM[0], c1 = -gM[1], c2 = -gM[2];
      gw[0] += b1 * c2 - b2 * c1;
      gw[1] += b2 * c0 - b0 * c2;
      gw[2] += b0 * c1 - b1 * c0;
      const float bb0 = c1 * w2 - c2 * w1, bb1 = c2 * w0 - c0 * w2, bb2 = c0 * w1 - c1 * w0;
      gw[0] = fmaf(I0, bb0, gw[0]);
      gw[1] = fmaf(I1, bb1, gw[1]);
      gw[2] = fmaf(I2, bb2, gw[2]);
    }
    // ---- qdot = 0.5 q (x) [0,w]
    {
      const Quat gqd{0.5f * vf[6], 0.5f * vf[7], 0.5f * vf[8], 0.5f * vf[9]};
      qacc(gq, qmul_qv(gqd, -w0, -w1, -w2));
      const Quat gOm = qmul(qc, gqd);
      gw[0] += gOm.x;
      gw[1] += gOm.y;
      gw[2] += gOm.z;
    }
    // ---- translational part: a = vec(q (x) Fq (x) conj(q))/m + [0,0,-g], Fq = [0, Fb].
    // The cotangent of Fb does not depend on Fb itself, so it is formed first and the
    // residual-force net is evaluated once (forward keeping act', then backward).
    const float inv_m = fast_rcp(m);
    const Quat R2 = qvec(vf[3] * inv_m, vf[4] * inv_m, vf[5] * inv_m);
    const Quat t2bar = qmul_vq(R2.x, R2.y, R2.z, q);  // r2 = t2 (x) qc  =>  t2bar = R2 (x) conj(qc)
    float gFb[3];
    {
      const Quat gFq = qmul(qc, t2bar);  // t2 = q (x) Fq  =>  Fqbar = conj(q) (x) t2bar
      gFb[0] = gFq.x;
      gFb[1] = gFq.y;
      gFb[2] = gFq.z;
    }
    float gvb[3] = {0.0f, 0.0f, 0.0f};
    float gin6[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    float Fb[3] = {0.0f, 0.0f, 0.0f};
    if constexpr (HAS_F) {
      float in[NetA::IN > 0 ? NetA::IN : 1], gi[NetA::IN > 0 ? NetA::IN : 1], out[3];
#pragma unroll
      for (int k = 0; k < NetA::IN; ++k) in[k] = in6[nth_bit(A::A_MASK, k)];
      if constexpr (FROM_TAPE) {
        out[0] = tp->Fb[0];
        out[1] = tp->Fb[1];
        out[2] = tp->Fb[2];
#pragma unroll
        for (int m = 0; m < TB1A; ++m) sc.d1[m] = tp->dA1[m];
#pragma unroll
        for (int m = 0; m < TB2A; ++m) sc.d2[m] = tp->dA2[m];
      } else {
        Ev::template fwd<NetA, true>(A::template ref<1, Ev::L>(W), sc, in, out);
      }
      Fb[0] = out[0];
      Fb[1] = out[1];
      Fb[2] = out[2];
      Ev::template bwd<NetA>(A::template ref<1, Ev::L>(W), sc, gFb, gi);
#pragma unroll
      for (int k = 0; k < NetA::IN; ++k) gin6[nth_bit(A::A_MASK, k)] += gi[k];
    }
    if constexpr (DRAG) {
      Fb[0] += -S[11] * vb[0];
      Fb[1] += -S[12] * vb[1];
      Fb[2] += -S[13] * vb[2] + S[14] * (vb[0] * vb[0] + vb[1] * vb[1]);
      gvb[0] += -S[11] * gFb[0] + 2.0f * S[14] * vb[0] * gFb[2];
      gvb[1] += -S[12] * gFb[1] + 2.0f * S[14] * vb[1] * gFb[2];
      gvb[2] += -S[13] * gFb[2];
    }
    {
      float T = 0.0f;
#pragma unroll
      for (int i = 0; i < NU; ++i) T = fmaf(Mixer<NU>::m(3, i), motor_thrust(S, u[i]), T);
      Fb[2] += T;
      const Quat Fq = qvec(Fb[0], Fb[1], Fb[2]);
      const Quat t2 = qmul_qv(q, Fb[0], Fb[1], Fb[2]);
      qacc(gq, qconj(qmul_qv(qconj(t2), R2.x, R2.y, R2.z)));  // qcbar = conj(t2) (x) R2
      qacc(gq, qmul_qv(t2bar, -Fb[0], -Fb[1], -Fb[2]));        // qbar += t2bar (x) conj(Fq)
    }
    // ---- motors
    {
      const float gT = gFb[2];
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        float gth = Mixer<NU>::m(3, i) * gT;
        gth = fmaf(Mixer<NU>::m(0, i) * S[4], gM[0], gth);
        gth = fmaf(Mixer<NU>::m(1, i) * S[5], gM[1], gth);
        gth = fmaf(Mixer<NU>::m(2, i) * S[6], gM[2], gth);
        gu[i] = fmaf(gth, motor_thrust_du(S, u[i]), gu[i]);
      }
    }
    float Mt[3] = {0.0f, 0.0f, 0.0f};  // residual moment (only needed for the scalar gradients)
    if constexpr (HAS_M) {
      float in[NetB::IN > 0 ? NetB::IN : 1], gi[NetB::IN > 0 ? NetB::IN : 1];
#pragma unroll
      for (int k = 0; k < NetB::IN; ++k) in[k] = in6[nth_bit(A::B_MASK, k)];
      if constexpr (FROM_TAPE) {
#pragma unroll
        for (int m = 0; m < TB1B; ++m) sc.d1[m] = tp->dB1[m];
#pragma unroll
        for (int m = 0; m < TB2B; ++m) sc.d2[m] = tp->dB2[m];
        Ev::template bwd<NetB>(A::template ref<2, Ev::L>(W), sc, gM, gi);
      } else if constexpr (Ev::PGRAD) {
        Ev::template fwd<NetB, true>(A::template ref<2, Ev::L>(W), sc, in, Mt);
        Ev::template bwd<NetB>(A::template ref<2, Ev::L>(W), sc, gM, gi);
      } else {
        Ev::template fwd_vjp<NetB>(A::template ref<2, Ev::L>(W), sc, in, gM, gi);
      }
#pragma unroll
      for (int k = 0; k < NetB::IN; ++k) gin6[nth_bit(A::B_MASK, k)] += gi[k];
    }
    if constexpr (Ev::PGRAD) {
      // ---- cotangents of the physical scalars S (packed order: mass Ixx Iyy Izz Lmx Lmy CoeffDrag c3 c2 c1 c0
      //      kdx kdy kdz kh), accumulated over the sweep; vf = lambda * dt is the cotangent of the drift
      float th[NU], sm0 = 0.0f, sm1 = 0.0f, sm2 = 0.0f, T = 0.0f;
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        th[i] = motor_thrust(S, u[i]);
        sm0 = fmaf(Mixer<NU>::m(0, i), th[i], sm0);
        sm1 = fmaf(Mixer<NU>::m(1, i), th[i], sm1);
        sm2 = fmaf(Mixer<NU>::m(2, i), th[i], sm2);
        T = fmaf(Mixer<NU>::m(3, i), th[i], T);
      }
      // a = rot(q, Fb)/m + g  =>  d/dm = -(a - g)/m ; Fb here already holds residual + drag + thrust
      {
        const Quat t2 = qmul_qv(q, Fb[0], Fb[1], Fb[2]);
        const Quat r2 = qmul(t2, qc);
        sc.gS[0] -= (vf[3] * r2.x + vf[4] * r2.y + vf[5] * r2.z) * inv_m * inv_m;
      }
      // alpha_k = (Mm_k + Mt_k - c_k)/I_k, c = w x (I w)
      {
        const float Mm0 = S[4] * sm0, Mm1 = S[5] * sm1, Mm2 = S[6] * sm2;
        const float c0 = w1 * I2 * w2 - w2 * I1 * w1, c1 = w2 * I0 * w0 - w0 * I2 * w2, c2 = w0 * I1 * w1 - w1 * I0 * w0;
        const float a0 = (Mm0 + Mt[0] - c0) * fast_rcp(I0), a1 = (Mm1 + Mt[1] - c1) * fast_rcp(I1),
                    a2 = (Mm2 + Mt[2] - c2) * fast_rcp(I2);
        sc.gS[1] += -gM[0] * a0 - gM[1] * (w2 * w0) + gM[2] * (w1 * w0);
        sc.gS[2] += -gM[1] * a1 + gM[0] * (w2 * w1) - gM[2] * (w0 * w1);
        sc.gS[3] += -gM[2] * a2 - gM[0] * (w1 * w2) + gM[1] * (w0 * w2);
        sc.gS[4] = fmaf(gM[0], sm0, sc.gS[4]);
        sc.gS[5] = fmaf(gM[1], sm1, sc.gS[5]);
        sc.gS[6] = fmaf(gM[2], sm2, sc.gS[6]);
      }
      // motor polynomial th(u) = ((c3 u + c2) u + c1) u + c0, cotangent gth_i as in the motor block above
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        float gth = Mixer<NU>::m(3, i) * gFb[2];
        gth = fmaf(Mixer<NU>::m(0, i) * S[4], gM[0], gth);
        gth = fmaf(Mixer<NU>::m(1, i) * S[5], gM[1], gth);
        gth = fmaf(Mixer<NU>::m(2, i) * S[6], gM[2], gth);
        const float ui = u[i];
        sc.gS[7] = fmaf(gth, ui * ui * ui, sc.gS[7]);
        sc.gS[8] = fmaf(gth, ui * ui, sc.gS[8]);
        sc.gS[9] = fmaf(gth, ui, sc.gS[9]);
        sc.gS[10] += gth;
      }
      if constexpr (DRAG) {
        sc.gS[11] -= gFb[0] * vb[0];
        sc.gS[12] -= gFb[1] * vb[1];
        sc.gS[13] -= gFb[2] * vb[2];
        sc.gS[14] = fmaf(gFb[2], vb[0] * vb[0] + vb[1] * vb[1], sc.gS[14]);
      }
      (void)T;
    }
    gvb[0] = fmaf(gin6[0], d.inv_state_scaling[3], gvb[0]);
    gvb[1] = fmaf(gin6[1], d.inv_state_scaling[4], gvb[1]);
    gvb[2] = fmaf(gin6[2], d.inv_state_scaling[5], gvb[2]);
    gw[0] = fmaf(gin6[3], d.inv_state_scaling[10], gw[0]);
    gw[1] = fmaf(gin6[4], d.inv_state_scaling[11], gw[1]);
    gw[2] = fmaf(gin6[5], d.inv_state_scaling[12], gw[2]);
    // ---- v_b = vec(qc (x) t1), t1 = V (x) q
    {
      const Quat R1 = qvec(gvb[0], gvb[1], gvb[2]);
      // r1 = qc (x) t1 : qcbar = R1 (x) conj(t1) ; t1bar = conj(qc) (x) R1 = q (x) R1
      qacc(gq, qconj(qmul_vq(gvb[0], gvb[1], gvb[2], qconj(t1))));
      const Quat t1bar = qmul_qv(q, gvb[0], gvb[1], gvb[2]);
      // t1 = V (x) q : Vbar = t1bar (x) conj(q) ; qbar += conj(V) (x) t1bar
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

  // q <- q/|q| (sde_rotor_model.py:160-171); aux = 1/|q_z| for the VJP
  SDE4_DEV static void project(float (&z)[NX], float& aux) {
    const float s = z[6] * z[6] + z[7] * z[7] + z[8] * z[8] + z[9] * z[9];
    float r = rsqrtf(s);
    r = r * fmaf(-0.5f * s * r, r, 1.5f);  // one Newton step -> ~0.5 ulp
    z[6] *= r;
    z[7] *= r;
    z[8] *= r;
    z[9] *= r;
    aux = r;
  }
  // lam (wrt projected state) -> lam (wrt z); xn = the projected x_{t+1} (still in registers from the
  // previous backward iteration) and aux = 1/|q_z|
  SDE4_DEV static void project_vjp(const float (&xn)[NX], float aux, float (&lam)[NX]) {
    const float q0 = xn[6], q1 = xn[7], q2 = xn[8], q3 = xn[9];
    const float dot = q0 * lam[6] + q1 * lam[7] + q2 * lam[8] + q3 * lam[9];
    lam[6] = (lam[6] - q0 * dot) * aux;
    lam[7] = (lam[7] - q1 * dot) * aux;
    lam[8] = (lam[8] - q2 * dot) * aux;
    lam[9] = (lam[9] - q3 * dot) * aux;
  }
};

template <class A>
struct ModelOf;

}  // namespace sde4

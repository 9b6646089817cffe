// nets.cuh -- hk.nets.MLP (two hidden layers) evaluated per thread in registers.
//
// Reference: hk.nets.MLP call sites sde4mbrl/nsde.py:376,387,674,
// sde_rotor_model.py:400,409, cartpole_sde.py:91,103, mass_spring_model.py:113.
// y = act(x @ w + b) for the hidden layers, the output layer is linear.
//
// Weights are read from shared memory with warp-uniform addresses (one LDS.128
// broadcast feeds 4 FFMA of every lane); all loops are fully unrolled so every
// register array index and every weight offset is a compile-time constant.
#pragma once
#include "common.cuh"

namespace sde4 {

// Packed layout of one layer: w[IN][up4(OUT)] row-major, then b[up4(OUT)].
template <int IN_, int H1_, int H2_, int OUT_, int ACT_>
struct Net {
  static constexpr int IN = IN_, H1 = H1_, H2 = H2_, OUT = OUT_, ACT = ACT_;
  static constexpr bool PRESENT = IN_ > 0;
  static constexpr int H1P = up4(H1_), H2P = up4(H2_), OUTP = up4(OUT_);
  static constexpr int OFF_W0 = 0;
  static constexpr int OFF_B0 = OFF_W0 + IN_ * H1P;
  static constexpr int OFF_W1 = OFF_B0 + H1P;
  static constexpr int OFF_B1 = OFF_W1 + H1_ * H2P;
  static constexpr int OFF_W2 = OFF_B1 + H2P;
  static constexpr int OFF_B2 = OFF_W2 + H2_ * OUTP;
  static constexpr int SIZE = PRESENT ? OFF_B2 + OUTP : 0;
  static constexpr int MACS = IN_ * H1_ + H1_ * H2_ + H2_ * OUT_;

  static bool matches(const sde4_net_desc& d) {
    if (!PRESENT) return d.n_in == 0;
    return d.n_in == IN && d.h1 == H1 && d.h2 == H2 && d.n_out == OUT && d.act == ACT;
  }
};

using NoNet = Net<0, 0, 0, 0, 0>;

// out[j] = b[j] + sum_i in[i] * w[i][j]
template <int IN, int OUT, int OUTP>
SDE4_DEV void dense_fwd(const float* __restrict__ w, const float* __restrict__ b, const float (&in)[IN],
                        float (&out)[OUT]) {
#pragma unroll
  for (int j = 0; j < OUT; ++j) out[j] = b[j];
#pragma unroll
  for (int i = 0; i < IN; ++i) {
#pragma unroll
    for (int j = 0; j < OUT; ++j) out[j] = fmaf(in[i], w[i * OUTP + j], out[j]);
  }
}

// gin[i] = sum_j w[i][j] * gout[j]
template <int IN, int OUT, int OUTP>
SDE4_DEV void dense_bwd(const float* __restrict__ w, const float (&gout)[OUT], float (&gin)[IN]) {
#pragma unroll
  for (int i = 0; i < IN; ++i) {
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < OUT; ++j) acc = fmaf(w[i * OUTP + j], gout[j], acc);
    gin[i] = acc;
  }
}

// Forward only.
template <class N>
SDE4_DEV void mlp_fwd(const float* __restrict__ W, const float (&in)[N::IN], float (&out)[N::OUT]) {
  float h1[N::H1];
  dense_fwd<N::IN, N::H1, N::H1P>(W + N::OFF_W0, W + N::OFF_B0, in, h1);
#pragma unroll
  for (int j = 0; j < N::H1; ++j) h1[j] = act_fwd<N::ACT>(h1[j]);
  float h2[N::H2];
  dense_fwd<N::H1, N::H2, N::H2P>(W + N::OFF_W1, W + N::OFF_B1, h1, h2);
#pragma unroll
  for (int j = 0; j < N::H2; ++j) h2[j] = act_fwd<N::ACT>(h2[j]);
  dense_fwd<N::H2, N::OUT, N::OUTP>(W + N::OFF_W2, W + N::OFF_B2, h2, out);
}

// Forward keeping act'(a) of the two hidden layers (H1 + H2 registers) for mlp_bwd.
template <class N>
SDE4_DEV void mlp_fwd_keep(const float* __restrict__ W, const float (&in)[N::IN], float (&out)[N::OUT],
                           float (&d1)[N::H1], float (&d2)[N::H2]) {
  float h1[N::H1];
  dense_fwd<N::IN, N::H1, N::H1P>(W + N::OFF_W0, W + N::OFF_B0, in, h1);
#pragma unroll
  for (int j = 0; j < N::H1; ++j) act_fwd_d<N::ACT>(h1[j], h1[j], d1[j]);
  float h2[N::H2];
  dense_fwd<N::H1, N::H2, N::H2P>(W + N::OFF_W1, W + N::OFF_B1, h1, h2);
#pragma unroll
  for (int j = 0; j < N::H2; ++j) act_fwd_d<N::ACT>(h2[j], h2[j], d2[j]);
  dense_fwd<N::H2, N::OUT, N::OUTP>(W + N::OFF_W2, W + N::OFF_B2, h2, out);
}

// Input-VJP: gin = (d out / d in)^T gout, from the kept activation derivatives.
template <class N>
SDE4_DEV void mlp_bwd(const float* __restrict__ W, const float (&d1)[N::H1], const float (&d2)[N::H2],
                      const float (&gout)[N::OUT], float (&gin)[N::IN]) {
  float g2[N::H2];
  dense_bwd<N::H2, N::OUT, N::OUTP>(W + N::OFF_W2, gout, g2);
#pragma unroll
  for (int j = 0; j < N::H2; ++j) g2[j] *= d2[j];
  float g1[N::H1];
  dense_bwd<N::H1, N::H2, N::H2P>(W + N::OFF_W1, g2, g1);
#pragma unroll
  for (int j = 0; j < N::H1; ++j) g1[j] *= d1[j];
  dense_bwd<N::IN, N::H1, N::H1P>(W + N::OFF_W0, g1, gin);
}

// Recompute-forward + input-VJP when the output cotangent is known up front.
template <class N>
SDE4_DEV void mlp_fwd_vjp(const float* __restrict__ W, const float (&in)[N::IN], const float (&gout)[N::OUT],
                          float (&gin)[N::IN]) {
  float d1[N::H1], d2[N::H2], out[N::OUT];
  mlp_fwd_keep<N>(W, in, out, d1, d2);
  mlp_bwd<N>(W, d1, d2, gout, gin);
}

}  // namespace sde4

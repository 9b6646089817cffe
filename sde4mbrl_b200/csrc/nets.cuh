// nets.cuh -- hk.nets.MLP (two hidden layers) evaluated per thread.
//
// Reference: hk.nets.MLP call sites sde4mbrl/nsde.py:376,387,674,
// sde_rotor_model.py:400,409, cartpole_sde.py:91,103, mass_spring_model.py:113.
// y = act(x @ w + b) for the hidden layers, the output layer is linear.
//
// Design notes (measured, profiles/r1_notes.md): a fully unrolled per-thread MLP makes the kernel
// instruction-fetch bound (320 KB of straight-line SASS, 41 % "no_instructions" stalls) and lets
// ptxas hoist hundreds of weight loads (2 KB of spills).  So only the small dimensions are
// unrolled: the hidden->hidden product is a ROLLED loop over input rows whose activations live in
// a per-thread column of shared memory (conflict free: element i of thread t at [i*BS + t]);
// the accumulators of one row sweep stay in registers; weights are read from shared memory with
// warp-uniform addresses (one LDS.128 broadcast feeds 4 FFMA of every lane).
#pragma once
#include "common.cuh"

namespace sde4 {

constexpr int MAXW = 32;  // widest hidden layer supported by the scratch layout

// Packed layout of one network (floats, every tensor 4-float aligned):
//   w0[IN][H1P] b0[H1P] w1[H1][H2P] b1[H2P] w2T[OUT][H2P] b2[OUTP]
// (w0, w1 row-major [in][out] exactly as haiku stores them, out padded to a multiple of 4;
//  the last layer is stored TRANSPOSED so that each output is a contiguous dot product).
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
  static constexpr int OFF_B2 = OFF_W2 + OUT_ * H2P;
  static constexpr int SIZE = PRESENT ? OFF_B2 + OUTP : 0;
  static constexpr int MACS = IN_ * H1_ + H1_ * H2_ + H2_ * OUT_;
  static_assert(H1_ <= MAXW && H2_ <= MAXW, "hidden width above MAXW");

  static bool matches(const sde4_net_desc& d) {
    if (!PRESENT) return d.n_in == 0;
    return d.n_in == IN && d.h1 == H1 && d.h2 == H2 && d.n_out == OUT && d.act == ACT;
  }
};

using NoNet = Net<0, 0, 0, 0, 0>;

// Per-thread scratch columns in shared memory: slot s, element i of this thread at
// base[(s*MAXW + i) * bs]  (base already points at this thread's lane).
struct Scratch {
  float* base;
  int bs;
  static constexpr int SLOT_H = 0;   // activations feeding the rolled row loop
  static constexpr int SLOT_D1 = 1;  // act'(a1) kept for the backward sweep
  static constexpr int SLOT_D2 = 2;  // act'(a2)
  static constexpr int SLOT_X = 3;   // noise dw[] / misc (NX <= 16 entries)
  static constexpr int NSLOTS = 4;
  SDE4_DEV float* slot(int s) const { return base + (size_t)s * MAXW * bs; }
};

// ---- layer 1: few inputs (registers), fully unrolled ----------------------------------------
template <int IN, int OUT, int OUTP>
SDE4_DEV void dense_in_regs(const float* w, const float* b, const float (&in)[IN], float (&out)[OUT]) {
#pragma unroll
  for (int j = 0; j < OUT; ++j) out[j] = b[j];
#pragma unroll
  for (int i = 0; i < IN; ++i) {
#pragma unroll
    for (int j = 0; j < OUT; ++j) out[j] = fmaf(in[i], w[i * OUTP + j], out[j]);
  }
}

// ---- layer 2: inputs in a scratch column, ROLLED over the input rows --------------------------
template <int IN, int OUT, int OUTP>
SDE4_DEV void dense_in_smem(const float* w, const float* b, const float* hin, int bs, float (&out)[OUT]) {
#pragma unroll
  for (int j = 0; j < OUT; ++j) out[j] = b[j];
#pragma unroll 4
  for (int i = 0; i < IN; ++i) {
    const float a = hin[(size_t)i * bs];
    const float* row = w + i * OUTP;
#pragma unroll
    for (int j = 0; j < OUT; ++j) out[j] = fmaf(a, row[j], out[j]);
  }
}

// ---- layer 3: few outputs, transposed weights, inputs in registers -----------------------------
template <int IN, int INP, int OUT>
SDE4_DEV void dense_out_t(const float* wT, const float* b, const float (&in)[IN], float (&out)[OUT]) {
#pragma unroll
  for (int o = 0; o < OUT; ++o) {
    float a0 = b[o], a1 = 0.0f;
#pragma unroll
    for (int j = 0; j < IN; j += 2) {
      a0 = fmaf(in[j], wT[o * INP + j], a0);
      if (j + 1 < IN) a1 = fmaf(in[j + 1], wT[o * INP + j + 1], a1);
    }
    out[o] = a0 + a1;
  }
}

// Forward.  KEEP: act' of both hidden layers is left in scratch slots D1 / D2 for mlp_bwd.
template <class N, bool KEEP>
SDE4_DEV void mlp_fwd(const float* W, const Scratch& sc, const float (&in)[N::IN], float (&out)[N::OUT]) {
  float* sh = sc.slot(Scratch::SLOT_H);
  {
    float a1[N::H1];
    dense_in_regs<N::IN, N::H1, N::H1P>(W + N::OFF_W0, W + N::OFF_B0, in, a1);
#pragma unroll
    for (int j = 0; j < N::H1; ++j) {
      float h, dd;
      if (KEEP) {
        act_fwd_d<N::ACT>(a1[j], h, dd);
        sc.slot(Scratch::SLOT_D1)[(size_t)j * sc.bs] = dd;
      } else {
        h = act_fwd<N::ACT>(a1[j]);
      }
      sh[(size_t)j * sc.bs] = h;
    }
  }
  float a2[N::H2];
  dense_in_smem<N::H1, N::H2, N::H2P>(W + N::OFF_W1, W + N::OFF_B1, sh, sc.bs, a2);
#pragma unroll
  for (int j = 0; j < N::H2; ++j) {
    if (KEEP) {
      float dd;
      act_fwd_d<N::ACT>(a2[j], a2[j], dd);
      sc.slot(Scratch::SLOT_D2)[(size_t)j * sc.bs] = dd;
    } else {
      a2[j] = act_fwd<N::ACT>(a2[j]);
    }
  }
  dense_out_t<N::H2, N::H2P, N::OUT>(W + N::OFF_W2, W + N::OFF_B2, a2, out);
}

// Shared tail of the backward sweeps: g2[] (cotangent of the pre-activation a2, registers) ->
// cotangent of a1 (rolled over the H1 rows, through scratch) -> gin.
template <class N>
SDE4_DEV void mlp_bwd_tail(const float* W, const Scratch& sc, const float (&g2)[N::H2], float (&gin)[N::IN]) {
  float* sh = sc.slot(Scratch::SLOT_H);
  const float* d1 = sc.slot(Scratch::SLOT_D1);
#pragma unroll 4
  for (int i = 0; i < N::H1; ++i) {
    const float* row = W + N::OFF_W1 + i * N::H2P;
    float r[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // 4 independent chains per row (FMA latency)
#pragma unroll
    for (int j = 0; j < N::H2; ++j) r[j & 3] = fmaf(row[j], g2[j], r[j & 3]);
    sh[(size_t)i * sc.bs] = ((r[0] + r[1]) + (r[2] + r[3])) * d1[(size_t)i * sc.bs];
  }
  float g1[N::H1];
#pragma unroll
  for (int i = 0; i < N::H1; ++i) g1[i] = sh[(size_t)i * sc.bs];
#pragma unroll
  for (int k = 0; k < N::IN; ++k) {
    const float* row = W + N::OFF_W0 + k * N::H1P;
    float r0 = 0.0f, r1 = 0.0f;
#pragma unroll
    for (int i = 0; i < N::H1; i += 2) {
      r0 = fmaf(row[i], g1[i], r0);
      if (i + 1 < N::H1) r1 = fmaf(row[i + 1], g1[i + 1], r1);
    }
    gin[k] = r0 + r1;
  }
}

// Input-VJP after mlp_fwd<N, true>:  gin = (d out / d in)^T gout.
template <class N>
SDE4_DEV void mlp_bwd(const float* W, const Scratch& sc, const float (&gout)[N::OUT], float (&gin)[N::IN]) {
  float g2[N::H2];
  const float* d2 = sc.slot(Scratch::SLOT_D2);
#pragma unroll
  for (int j = 0; j < N::H2; ++j) {
    float s = 0.0f;
#pragma unroll
    for (int o = 0; o < N::OUT; ++o) s = fmaf(W[N::OFF_W2 + o * N::H2P + j], gout[o], s);
    g2[j] = s * d2[(size_t)j * sc.bs];
  }
  mlp_bwd_tail<N>(W, sc, g2, gin);
}

// Recompute-forward + input-VJP when the output cotangent is known up front (nothing but act'(a1)
// goes through scratch: the cotangent of a2 is formed in registers right after layer 2).
template <class N>
SDE4_DEV void mlp_fwd_vjp(const float* W, const Scratch& sc, const float (&in)[N::IN], const float (&gout)[N::OUT],
                          float (&gin)[N::IN]) {
  float* sh = sc.slot(Scratch::SLOT_H);
  {
    float a1[N::H1];
    dense_in_regs<N::IN, N::H1, N::H1P>(W + N::OFF_W0, W + N::OFF_B0, in, a1);
#pragma unroll
    for (int j = 0; j < N::H1; ++j) {
      float h, dd;
      act_fwd_d<N::ACT>(a1[j], h, dd);
      sc.slot(Scratch::SLOT_D1)[(size_t)j * sc.bs] = dd;
      sh[(size_t)j * sc.bs] = h;
    }
  }
  float g2[N::H2];
  dense_in_smem<N::H1, N::H2, N::H2P>(W + N::OFF_W1, W + N::OFF_B1, sh, sc.bs, g2);
#pragma unroll
  for (int j = 0; j < N::H2; ++j) {
    float h, dd;
    act_fwd_d<N::ACT>(g2[j], h, dd);
    float s = 0.0f;
#pragma unroll
    for (int o = 0; o < N::OUT; ++o) s = fmaf(W[N::OFF_W2 + o * N::H2P + j], gout[o], s);
    g2[j] = s * dd;
  }
  mlp_bwd_tail<N>(W, sc, g2, gin);
}

// Scalar-output MLP (the diffusion density net, nsde.py:387-407): value and input-gradient in one
// pass:  out = mlp(in),  J[k] = d out / d in[k].
template <class N>
SDE4_DEV void mlp_scalar_value_grad(const float* W, const Scratch& sc, const float (&in)[N::IN], float& out,
                                    float (&J)[N::IN]) {
  static_assert(N::OUT == 1, "scalar-output network expected");
  float* sh = sc.slot(Scratch::SLOT_H);
  {
    float a1[N::H1];
    dense_in_regs<N::IN, N::H1, N::H1P>(W + N::OFF_W0, W + N::OFF_B0, in, a1);
#pragma unroll
    for (int j = 0; j < N::H1; ++j) {
      float h, dd;
      act_fwd_d<N::ACT>(a1[j], h, dd);
      sc.slot(Scratch::SLOT_D1)[(size_t)j * sc.bs] = dd;
      sh[(size_t)j * sc.bs] = h;
    }
  }
  float g2[N::H2];
  dense_in_smem<N::H1, N::H2, N::H2P>(W + N::OFF_W1, W + N::OFF_B1, sh, sc.bs, g2);
  float o = W[N::OFF_B2];
#pragma unroll
  for (int j = 0; j < N::H2; ++j) {
    float h, dd;
    act_fwd_d<N::ACT>(g2[j], h, dd);
    const float w2 = W[N::OFF_W2 + j];
    o = fmaf(w2, h, o);
    g2[j] = w2 * dd;
  }
  out = o;
  mlp_bwd_tail<N>(W, sc, g2, J);
}

}  // namespace sde4

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
  // rows of the parameter-gradient stream of this network
  static constexpr int R_IN = 0, R_H1 = IN_, R_H2 = IN_ + H1_, R_C1 = IN_ + H1_ + H2_, R_G2 = R_C1 + H1_,
                       R_GOUT = R_G2 + H2_, DUMP_ROWS = PRESENT ? R_GOUT + OUT_ : 0;
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
  static constexpr int SLOT_X = 1;   // noise dw[] / misc (NX <= 16 entries)
  static constexpr int SLOT_D1 = 2;  // act'(a1) kept for the backward sweep
  static constexpr int SLOT_D2 = 3;  // act'(a2)
  static constexpr int NSLOTS = 4;      // kernels with a backward sweep
  static constexpr int NSLOTS_FWD = 2;  // forward-only kernels (rollout, value-only cost): 2x the resident CTAs
  SDE4_DEV float* slot(int s) const { return base + (size_t)s * MAXW * bs; }
  // parameter-gradient streams (training-loss kernel only, DUMP evaluators): for network `slot` the
  // rows [in | h1 | h2 | c1 | g2 | gout] of this thread's (step, sample) column, see wgrad in misc_kernels.cuh
  float* dnet[3];
  size_t dstride, dcol;
  SDE4_DEV void dump(int slot_, int row, float v) const { dnet[slot_][(size_t)row * dstride + dcol] = v; }
  // training loss: cotangents of the model's physical scalars, accumulated over the backward sweep (registers;
  // unused members vanish in the other instantiations)
  float gS[16];
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
template <class N, bool KEEP, bool DUMP = false>
SDE4_DEV void mlp_fwd(const float* W, const Scratch& sc, const float (&in)[N::IN], float (&out)[N::OUT], int slot = 0) {
  float* sh = sc.slot(Scratch::SLOT_H);
  if (DUMP && KEEP) {
#pragma unroll
    for (int k = 0; k < N::IN; ++k) sc.dump(slot, N::R_IN + k, in[k]);
  }
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
      if (DUMP && KEEP) sc.dump(slot, N::R_H1 + j, h);
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
      if (DUMP) sc.dump(slot, N::R_H2 + j, a2[j]);
    } else {
      a2[j] = act_fwd<N::ACT>(a2[j]);
    }
  }
  dense_out_t<N::H2, N::H2P, N::OUT>(W + N::OFF_W2, W + N::OFF_B2, a2, out);
}

// Shared tail of the backward sweeps: g2[] (cotangent of the pre-activation a2, registers) ->
// cotangent of a1 (rolled over the H1 rows, through scratch) -> gin.
template <class N, bool DUMP = false>
SDE4_DEV void mlp_bwd_tail(const float* W, const Scratch& sc, const float (&g2)[N::H2], float (&gin)[N::IN], int slot = 0) {
  float* sh = sc.slot(Scratch::SLOT_H);
  const float* d1 = sc.slot(Scratch::SLOT_D1);
#pragma unroll 4
  for (int i = 0; i < N::H1; ++i) {
    const float* row = W + N::OFF_W1 + i * N::H2P;
    float r[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // 4 independent chains per row (FMA latency)
#pragma unroll
    for (int j = 0; j < N::H2; ++j) r[j & 3] = fmaf(row[j], g2[j], r[j & 3]);
    const float c1 = ((r[0] + r[1]) + (r[2] + r[3])) * d1[(size_t)i * sc.bs];
    sh[(size_t)i * sc.bs] = c1;
    if (DUMP) sc.dump(slot, N::R_C1 + i, c1);
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
template <class N, bool DUMP = false>
SDE4_DEV void mlp_bwd(const float* W, const Scratch& sc, const float (&gout)[N::OUT], float (&gin)[N::IN], int slot = 0) {
  float g2[N::H2];
  const float* d2 = sc.slot(Scratch::SLOT_D2);
  if (DUMP) {
#pragma unroll
    for (int o = 0; o < N::OUT; ++o) sc.dump(slot, N::R_GOUT + o, gout[o]);
  }
#pragma unroll
  for (int j = 0; j < N::H2; ++j) {
    float s = 0.0f;
#pragma unroll
    for (int o = 0; o < N::OUT; ++o) s = fmaf(W[N::OFF_W2 + o * N::H2P + j], gout[o], s);
    g2[j] = s * d2[(size_t)j * sc.bs];
    if (DUMP) sc.dump(slot, N::R_G2 + j, g2[j]);
  }
  mlp_bwd_tail<N, DUMP>(W, sc, g2, gin, slot);
}

// Recompute-forward + input-VJP when the output cotangent is known up front (nothing but act'(a1)
// goes through scratch: the cotangent of a2 is formed in registers right after layer 2).
template <class N, bool DUMP = false>
SDE4_DEV void mlp_fwd_vjp(const float* W, const Scratch& sc, const float (&in)[N::IN], const float (&gout)[N::OUT],
                          float (&gin)[N::IN], int slot = 0) {
  float* sh = sc.slot(Scratch::SLOT_H);
  if (DUMP) {
#pragma unroll
    for (int k = 0; k < N::IN; ++k) sc.dump(slot, N::R_IN + k, in[k]);
#pragma unroll
    for (int o = 0; o < N::OUT; ++o) sc.dump(slot, N::R_GOUT + o, gout[o]);
  }
  {
    float a1[N::H1];
    dense_in_regs<N::IN, N::H1, N::H1P>(W + N::OFF_W0, W + N::OFF_B0, in, a1);
#pragma unroll
    for (int j = 0; j < N::H1; ++j) {
      float h, dd;
      act_fwd_d<N::ACT>(a1[j], h, dd);
      sc.slot(Scratch::SLOT_D1)[(size_t)j * sc.bs] = dd;
      sh[(size_t)j * sc.bs] = h;
      if (DUMP) sc.dump(slot, N::R_H1 + j, h);
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
    if (DUMP) {
      sc.dump(slot, N::R_H2 + j, h);
      sc.dump(slot, N::R_G2 + j, g2[j]);
    }
  }
  mlp_bwd_tail<N, DUMP>(W, sc, g2, gin, slot);
}

// Scalar-output MLP (the diffusion density net, nsde.py:387-407): value and input-VJP in one pass:
//   out = mlp(in),  gd = gd_of(out) (the output cotangent, which depends on the value),
//   J[k] = gd * d out / d in[k].
template <class N, bool DUMP = false, class GdFn>
SDE4_DEV void mlp_scalar_value_grad(const float* W, const Scratch& sc, const float (&in)[N::IN], float& out,
                                    float (&J)[N::IN], GdFn gd_of, int slot = 0) {
  static_assert(N::OUT == 1, "scalar-output network expected");
  float* sh = sc.slot(Scratch::SLOT_H);
  if (DUMP) {
#pragma unroll
    for (int k = 0; k < N::IN; ++k) sc.dump(slot, N::R_IN + k, in[k]);
  }
  {
    float a1[N::H1];
    dense_in_regs<N::IN, N::H1, N::H1P>(W + N::OFF_W0, W + N::OFF_B0, in, a1);
#pragma unroll
    for (int j = 0; j < N::H1; ++j) {
      float h, dd;
      act_fwd_d<N::ACT>(a1[j], h, dd);
      sc.slot(Scratch::SLOT_D1)[(size_t)j * sc.bs] = dd;
      sh[(size_t)j * sc.bs] = h;
      if (DUMP) sc.dump(slot, N::R_H1 + j, h);
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
    if (DUMP) sc.dump(slot, N::R_H2 + j, h);
  }
  out = o;
  const float gd = gd_of(o);
#pragma unroll
  for (int j = 0; j < N::H2; ++j) {
    g2[j] *= gd;
    if (DUMP) sc.dump(slot, N::R_G2 + j, g2[j]);
  }
  if (DUMP) sc.dump(slot, N::R_GOUT, gd);
  mlp_bwd_tail<N, DUMP>(W, sc, g2, J, slot);
}


// ================================================================================================
// Evaluator policies.  The model code (models.cuh) is written once against
//   Ev::fwd<N,KEEP>, Ev::bwd<N>, Ev::fwd_vjp<N>, Ev::scalar_value_grad<N>
// taking a NetRef (pointers into the shared-memory parameter image) and an evaluator context.
//   Ev1      one thread per sample  (throughput configurations: many samples)
//   EvL<L>   L adjacent lanes per sample (latency configurations: few samples, e.g. one MPC
//            problem with 4096 particles = only 128 warps for 592 warp schedulers)
// ================================================================================================
struct NetRef {
  int slot;          // 0 density, 1 net A, 2 net B
  const float* w;    // packed network image (layout above)
};

template <bool DUMP>
struct Ev1T {
  static constexpr int L = 1;
  static constexpr bool PGRAD = DUMP;  // also streams what the parameter gradients need (training loss)
  using Ctx = Scratch;
  template <class N, bool KEEP>
  SDE4_DEV static void fwd(const NetRef& r, Ctx& c, const float (&in)[N::IN], float (&out)[N::OUT]) {
    mlp_fwd<N, KEEP, DUMP>(r.w, c, in, out, r.slot);
  }
  template <class N>
  SDE4_DEV static void bwd(const NetRef& r, Ctx& c, const float (&gout)[N::OUT], float (&gin)[N::IN]) {
    mlp_bwd<N, DUMP>(r.w, c, gout, gin, r.slot);
  }
  template <class N>
  SDE4_DEV static void fwd_vjp(const NetRef& r, Ctx& c, const float (&in)[N::IN], const float (&gout)[N::OUT],
                               float (&gin)[N::IN]) {
    mlp_fwd_vjp<N, DUMP>(r.w, c, in, gout, gin, r.slot);
  }
  template <class N, class GdFn>
  SDE4_DEV static void scalar_value_grad(const NetRef& r, Ctx& c, const float (&in)[N::IN], float& out,
                                         float (&J)[N::IN], GdFn gd_of) {
    mlp_scalar_value_grad<N, DUMP>(r.w, c, in, out, J, gd_of, r.slot);
  }
};
using Ev1 = Ev1T<false>;
using Ev1P = Ev1T<true>;

// ---- lane-split evaluator ------------------------------------------------------------------------
// Lane l of a sample owns the hidden units [l*B, (l+1)*B) of each hidden layer (B = H/L).
//   layer 1: every lane has the (replicated) inputs and computes its own block of outputs;
//   layer 2: the L lanes all-gather h1 with log2(L) rounds of SHFL.BFLY (XOR order, no selects) and
//            every lane forms its own block of outputs from its column chunk of w1;
//   layer 3: partial dot products over the own block + butterfly all-reduce.
// The backward sweep mirrors this with the transposed products.
// DUMP (training loss): every lane also streams the rows of its own hidden units -- and its share of the
// replicated input / output rows -- of [in | h1 | h2 | c1 | g2 | gout] for the weight-gradient kernels (K5).
template <int L_, bool DUMP_ = false>
struct EvLT {
  static constexpr int L = L_;
  static constexpr bool PGRAD = DUMP_, DUMP = DUMP_;
  static_assert(L_ == 2 || L_ == 4 || L_ == 8 || L_ == 16, "lanes per sample");
  struct Ctx {
    int l;             // lane within the sample's group
    float d1[32 / L_ > 0 ? 32 / L_ : 1], d2[32 / L_ > 0 ? 32 / L_ : 1];  // act' of the own blocks (KEEP)
    float* dnet[3];    // DUMP: see Scratch
    size_t dstride, dcol;
    float gS[16];      // DUMP: cotangents of the model's physical scalars (see Scratch)
    SDE4_DEV void dump(int slot_, int row, float v) const { dnet[slot_][(size_t)row * dstride + dcol] = v; }
    // a replicated row (inputs, output cotangents): written by lane (k mod L)
    SDE4_DEV void dump_rep(int slot_, int row0, int k, float v) const {
      if ((k & (L_ - 1)) == l) dump(slot_, row0 + k, v);
    }
  };
  // own block size of a hidden layer of width H: ceil(H / L).  Widths that are not multiples of L are PADDED: the
  // packed image stores every layer with its width rounded up to a multiple of 4 (zero weights / biases), so ghost
  // units l*B+m >= H are legal to address as long as L*B <= up4(H); their activations and cotangents are forced to 0.
  static constexpr int blk(int H) { return (H + L_ - 1) / L_; }
  template <class N>
  static constexpr bool ok() {
    return blk(N::H1) * L <= N::H1P && blk(N::H2) * L <= N::H2P;
  }
  template <int H, int B>
  SDE4_DEV static float live(int l, int m, float v) {  // v for a real unit, 0 for a ghost one
    if constexpr (B * L_ == H) return v;
    else return (l * B + m < H) ? v : 0.0f;
  }
  template <int H, int B>
  SDE4_DEV static bool is_live(int l, int m) {
    if constexpr (B * L_ == H) return true;
    else return l * B + m < H;
  }

  SDE4_DEV static float all_reduce(float v) {
#pragma unroll
    for (int mask = L / 2; mask >= 1; mask >>= 1) v += __shfl_xor_sync(0xffffffffu, v, mask);
    return v;
  }

  // layer 1 own block: a1[m] = b0[l*B1+m] + sum_k in[k]*w0[k][l*B1+m]
  template <class N>
  SDE4_DEV static void layer1(const NetRef& r, int l, const float (&in)[N::IN], float (&a1)[EvLT::blk(N::H1)]) {
    constexpr int B1 = EvLT::blk(N::H1);
    const float* w0 = r.w + N::OFF_W0 + l * B1;
    const float* b0 = r.w + N::OFF_B0 + l * B1;
#pragma unroll
    for (int m = 0; m < B1; ++m) a1[m] = b0[m];
#pragma unroll
    for (int k = 0; k < N::IN; ++k) {
#pragma unroll
      for (int m = 0; m < B1; ++m) a1[m] = fmaf(in[k], w0[k * N::H1P + m], a1[m]);
    }
  }
  // all-gather over the sample's L lanes: local block q of `all` = own block of lane (l ^ q).
  // XOR order: no selects are needed, the consumer indexes the weights with (l ^ q) instead.
  template <int B>
  SDE4_DEV static void all_gather(const float (&own)[B], float (&all)[B * L]) {
#pragma unroll
    for (int m = 0; m < B; ++m) all[m] = own[m];
#ifndef SDE4_GATHER_DIRECT
    // doubling butterfly: log2(L) dependent rounds
#pragma unroll
    for (int mask = 1, n = B; mask < L; mask <<= 1, n <<= 1) {
#pragma unroll
      for (int q = 0; q < n; ++q) all[n + q] = __shfl_xor_sync(0xffffffffu, all[q], mask);
    }
#else
    // every block straight from its owner: the same (L-1)*B shuffles, none depending on another.  Measured in round 2:
    // the 8-lane forward rollout of C3 went from 0.178 to 0.235 ms with it (the independent shuffles are hoisted and
    // lengthen the live ranges), so the butterfly stays the default.
#pragma unroll
    for (int q = 1; q < L; ++q) {
#pragma unroll
      for (int m = 0; m < B; ++m) all[q * B + m] = __shfl_xor_sync(0xffffffffu, own[m], q);
    }
#endif
  }
  // layer 2, OUTPUT split: gather all of h1 (SHFL.BFLY), every lane forms its own block of outputs
  //   a2own[m] = b1[l*B2+m] + sum_i h1[i] * w1[i][l*B2+m]
  // (lanes read distinct 4-bank column chunks of distinct rows: conflict free without padding)
  template <class N>
  SDE4_DEV static void layer2(const NetRef& r, int l, const float (&h1)[EvLT::blk(N::H1)], float (&a2)[EvLT::blk(N::H2)]) {
    constexpr int B1 = EvLT::blk(N::H1), B2 = EvLT::blk(N::H2);
    float hall[B1 * L];
    all_gather<B1>(h1, hall);
    const float* w1 = r.w + N::OFF_W1 + l * B2;
    const float* b1 = r.w + N::OFF_B1 + l * B2;
#pragma unroll
    for (int m = 0; m < B2; ++m) a2[m] = b1[m];
#pragma unroll
    for (int q = 0; q < L; ++q) {
#pragma unroll
      for (int mm = 0; mm < B1; ++mm) {
        // row of w1 = gathered unit; ghost units (h = 0) are redirected to the last real row (no out-of-image read)
        int row = (l ^ q) * B1 + mm;
        if constexpr (B1 * L != N::H1) row = min(row, N::H1 - 1);
        const float* wq = w1 + row * N::H2P;
#pragma unroll
        for (int m = 0; m < B2; ++m) a2[m] = fmaf(hall[q * B1 + mm], wq[m], a2[m]);
      }
    }
  }
  // layer 3: out[o] = b2[o] + sum_j h2[j]*w2T[o][j]
  template <class N>
  SDE4_DEV static void layer3(const NetRef& r, int l, const float (&h2)[EvLT::blk(N::H2)], float (&out)[N::OUT]) {
    constexpr int B2 = EvLT::blk(N::H2);
    const float* w2 = r.w + N::OFF_W2 + l * B2;
#pragma unroll
    for (int o = 0; o < N::OUT; ++o) {
      float a = 0.0f;
#pragma unroll
      for (int m = 0; m < B2; ++m) a = fmaf(h2[m], w2[o * N::H2P + m], a);
      out[o] = all_reduce(a) + r.w[N::OFF_B2 + o];
    }
  }
  // backward tail: g2own (cotangent of own a2 block) -> gin (replicated)
  //   c1own[m] = d1[m] * sum_j w1[l*B1+m][j] * g2[j]   (g2 gathered over the lanes)
  template <class N>
  SDE4_DEV static void bwd_tail(const NetRef& r, const Ctx& c, const float (&g2)[EvLT::blk(N::H2)], const float (&d1)[EvLT::blk(N::H1)],
                                float (&gin)[N::IN]) {
    constexpr int B1 = EvLT::blk(N::H1), B2 = EvLT::blk(N::H2);
    const int l = c.l;
    float gall[B2 * L];
    all_gather<B2>(g2, gall);
    float c1[B1];
#pragma unroll
    for (int m = 0; m < B1; ++m) {
      float a0 = 0.0f, a1 = 0.0f;
      int row = l * B1 + m;
      if constexpr (B1 * L != N::H1) row = min(row, N::H1 - 1);  // ghost unit: any real row, the result is masked
      const float* w1 = r.w + N::OFF_W1 + row * N::H2P;
#pragma unroll
      for (int q = 0; q < L; ++q) {
        const float* wq = w1 + (l ^ q) * B2;
#pragma unroll
        for (int mm = 0; mm < B2; ++mm) {
          if ((q & 1) == 0) a0 = fmaf(wq[mm], gall[q * B2 + mm], a0);
          else a1 = fmaf(wq[mm], gall[q * B2 + mm], a1);
        }
      }
      c1[m] = live<N::H1, B1>(l, m, (a0 + a1) * d1[m]);
      if constexpr (DUMP) {
        if (is_live<N::H1, B1>(l, m)) c.dump(r.slot, N::R_C1 + l * B1 + m, c1[m]);
      }
    }
    const float* w0 = r.w + N::OFF_W0 + l * B1;
#pragma unroll
    for (int k = 0; k < N::IN; ++k) {
      float a = 0.0f;
#pragma unroll
      for (int m = 0; m < B1; ++m) a = fmaf(w0[k * N::H1P + m], c1[m], a);
      gin[k] = all_reduce(a);
    }
  }

  template <class N, bool KEEP>
  SDE4_DEV static void fwd(const NetRef& r, Ctx& c, const float (&in)[N::IN], float (&out)[N::OUT]) {
    static_assert(ok<N>(), "hidden widths (rounded up to a multiple of 4) must cover ceil(H/L)*L units");
    constexpr int B1 = EvLT::blk(N::H1), B2 = EvLT::blk(N::H2);
    float a1[B1], a2[B2];
    if constexpr (DUMP && KEEP) {
#pragma unroll
      for (int k = 0; k < N::IN; ++k) c.dump_rep(r.slot, N::R_IN, k, in[k]);
    }
    layer1<N>(r, c.l, in, a1);
#pragma unroll
    for (int m = 0; m < B1; ++m) {
      if (KEEP) act_fwd_d<N::ACT>(a1[m], a1[m], c.d1[m]);
      else a1[m] = act_fwd<N::ACT>(a1[m]);
      a1[m] = live<N::H1, B1>(c.l, m, a1[m]);
      if constexpr (DUMP && KEEP) {
        if (is_live<N::H1, B1>(c.l, m)) c.dump(r.slot, N::R_H1 + c.l * B1 + m, a1[m]);
      }
    }
    layer2<N>(r, c.l, a1, a2);
#pragma unroll
    for (int m = 0; m < B2; ++m) {
      if (KEEP) act_fwd_d<N::ACT>(a2[m], a2[m], c.d2[m]);
      else a2[m] = act_fwd<N::ACT>(a2[m]);
      a2[m] = live<N::H2, B2>(c.l, m, a2[m]);
      if constexpr (DUMP && KEEP) {
        if (is_live<N::H2, B2>(c.l, m)) c.dump(r.slot, N::R_H2 + c.l * B2 + m, a2[m]);
      }
    }
    layer3<N>(r, c.l, a2, out);
  }
  template <class N>
  SDE4_DEV static void bwd(const NetRef& r, Ctx& c, const float (&gout)[N::OUT], float (&gin)[N::IN]) {
    constexpr int B1 = EvLT::blk(N::H1), B2 = EvLT::blk(N::H2);
    float g2[B2], d1[B1];
    const float* w2 = r.w + N::OFF_W2 + c.l * B2;
    if constexpr (DUMP) {
#pragma unroll
      for (int o = 0; o < N::OUT; ++o) c.dump_rep(r.slot, N::R_GOUT, o, gout[o]);
    }
#pragma unroll
    for (int m = 0; m < B2; ++m) {
      float s = 0.0f;
#pragma unroll
      for (int o = 0; o < N::OUT; ++o) s = fmaf(w2[o * N::H2P + m], gout[o], s);
      g2[m] = live<N::H2, B2>(c.l, m, s * c.d2[m]);
      if constexpr (DUMP) {
        if (is_live<N::H2, B2>(c.l, m)) c.dump(r.slot, N::R_G2 + c.l * B2 + m, g2[m]);
      }
    }
#pragma unroll
    for (int m = 0; m < B1; ++m) d1[m] = c.d1[m];
    bwd_tail<N>(r, c, g2, d1, gin);
  }
  template <class N>
  SDE4_DEV static void fwd_vjp(const NetRef& r, Ctx& c, const float (&in)[N::IN], const float (&gout)[N::OUT],
                               float (&gin)[N::IN]) {
    static_assert(ok<N>(), "hidden widths (rounded up to a multiple of 4) must cover ceil(H/L)*L units");
    constexpr int B1 = EvLT::blk(N::H1), B2 = EvLT::blk(N::H2);
    float a1[B1], d1[B1], a2[B2];
    if constexpr (DUMP) {
#pragma unroll
      for (int k = 0; k < N::IN; ++k) c.dump_rep(r.slot, N::R_IN, k, in[k]);
#pragma unroll
      for (int o = 0; o < N::OUT; ++o) c.dump_rep(r.slot, N::R_GOUT, o, gout[o]);
    }
    layer1<N>(r, c.l, in, a1);
#pragma unroll
    for (int m = 0; m < B1; ++m) {
      act_fwd_d<N::ACT>(a1[m], a1[m], d1[m]);
      a1[m] = live<N::H1, B1>(c.l, m, a1[m]);
      if constexpr (DUMP) {
        if (is_live<N::H1, B1>(c.l, m)) c.dump(r.slot, N::R_H1 + c.l * B1 + m, a1[m]);
      }
    }
    layer2<N>(r, c.l, a1, a2);
    const float* w2 = r.w + N::OFF_W2 + c.l * B2;
#pragma unroll
    for (int m = 0; m < B2; ++m) {
      float h, dd;
      act_fwd_d<N::ACT>(a2[m], h, dd);
      float s = 0.0f;
#pragma unroll
      for (int o = 0; o < N::OUT; ++o) s = fmaf(w2[o * N::H2P + m], gout[o], s);
      a2[m] = live<N::H2, B2>(c.l, m, s * dd);
      if constexpr (DUMP) {
        if (is_live<N::H2, B2>(c.l, m)) {
          c.dump(r.slot, N::R_H2 + c.l * B2 + m, h);
          c.dump(r.slot, N::R_G2 + c.l * B2 + m, a2[m]);
        }
      }
    }
    bwd_tail<N>(r, c, a2, d1, gin);
  }
  template <class N, class GdFn>
  SDE4_DEV static void scalar_value_grad(const NetRef& r, Ctx& c, const float (&in)[N::IN], float& out,
                                         float (&J)[N::IN], GdFn gd_of) {
    static_assert(N::OUT == 1 && ok<N>(), "scalar-output network with lane-coverable widths expected");
    constexpr int B1 = EvLT::blk(N::H1), B2 = EvLT::blk(N::H2);
    float a1[B1], d1[B1], a2[B2];
    if constexpr (DUMP) {
#pragma unroll
      for (int k = 0; k < N::IN; ++k) c.dump_rep(r.slot, N::R_IN, k, in[k]);
    }
    layer1<N>(r, c.l, in, a1);
#pragma unroll
    for (int m = 0; m < B1; ++m) {
      act_fwd_d<N::ACT>(a1[m], a1[m], d1[m]);
      a1[m] = live<N::H1, B1>(c.l, m, a1[m]);
      if constexpr (DUMP) {
        if (is_live<N::H1, B1>(c.l, m)) c.dump(r.slot, N::R_H1 + c.l * B1 + m, a1[m]);
      }
    }
    layer2<N>(r, c.l, a1, a2);
    const float* w2 = r.w + N::OFF_W2 + c.l * B2;
    float o = 0.0f;
#pragma unroll
    for (int m = 0; m < B2; ++m) {
      float h, dd;
      act_fwd_d<N::ACT>(a2[m], h, dd);
      h = live<N::H2, B2>(c.l, m, h);
      o = fmaf(w2[m], h, o);
      a2[m] = live<N::H2, B2>(c.l, m, w2[m] * dd);
      if constexpr (DUMP) {
        if (is_live<N::H2, B2>(c.l, m)) c.dump(r.slot, N::R_H2 + c.l * B2 + m, h);
      }
    }
    out = all_reduce(o) + r.w[N::OFF_B2];
    const float gd = gd_of(out);
#pragma unroll
    for (int m = 0; m < B2; ++m) {
      a2[m] *= gd;
      if constexpr (DUMP) {
        if (is_live<N::H2, B2>(c.l, m)) c.dump(r.slot, N::R_G2 + c.l * B2 + m, a2[m]);
      }
    }
    if constexpr (DUMP) c.dump_rep(r.slot, N::R_GOUT, 0, gd);
    bwd_tail<N>(r, c, a2, d1, J);
  }
};
template <int L_>
using EvL = EvLT<L_, false>;

}  // namespace sde4

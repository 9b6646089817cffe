// kernels_tc.cuh -- forward rollout for the throughput configurations with the 32 x 32 hidden->hidden products of the
// networks on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32 split, accumulator in tensor memory).
//
// Reference path: sde_solve / euler_maruyama (sde4mbrl/sde_solver.py:245-334) vmapped over particles and start states
// (nsde.py:1024-1028, 1322-1323) -- the same K1 as kernels.cuh, one thread per sample, CTA = one tile of 128 samples.
// Measured basis: profiles/r2_tcgen05_experiment.md (one layer costs 1.9x less than the FP32 row loop with two resident
// tiles per SM, 2.8x less at full occupancy, error 1e-6).
//
// Per network with hidden_layers [32, 32] (hk.nets.MLP, nsde.py:387, sde_rotor_model.py:400-409, cartpole_sde.py:91):
//   layer 1 in registers (few inputs) -> activation -> every thread splits its 32 values into TF32 high / low parts
//   (hi = x & 0xffffe000, lo = x - hi) and stores its row of the A tiles (K-major, no swizzle, 8 x 16-byte core
//   matrices) -> fence.proxy.async + CTA barrier -> thread 0 issues 4 K steps x {hi*hi, lo*hi, hi*lo} MMAs against the
//   pre-split weight tiles and commits to an mbarrier -> every thread reads its accumulator row (tcgen05.ld 32x32b.x32),
//   adds the bias, applies the activation and forms the few outputs of layer 3 in registers.
// Networks of other shapes keep the FP32 evaluator of nets.cuh.
#pragma once
#include "kernels.cuh"

namespace sde4 {

namespace tc {
constexpr int TM = 128, TK = 32, TN = 32;
constexpr uint32_t A_LBO = (TM / 8) * 128, A_SBO = 128, A_TILE = TM * TK * 4;  // 2048, 128, 16 KB
constexpr uint32_t B_LBO = (TN / 8) * 128, B_SBO = 128, B_TILE = TN * TK * 4;  // 512, 128, 4 KB
// instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

SDE4_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory matrix descriptor: [0,14) address >> 4, [16,30) LBO >> 4 (distance of the two 16-byte K chunks of one
// instruction), [32,46) SBO >> 4 (distance of consecutive 8-row groups), [46,48) version = 1, no swizzle
SDE4_DEV uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
SDE4_DEV void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
      : "memory");
}
SDE4_DEV float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

}  // namespace tc

// evaluator context: the FP32 scratch columns plus the tile state of this CTA
struct ScratchTC : Scratch {
  unsigned char* a_row;  // this thread's 16-byte slot in chunk 0 of A_hi (A_lo: + A_TILE; chunk c: + c * A_LBO)
  uint64_t da_hi, da_lo;  // descriptors of the A tiles
  uint64_t db[3], dbl[3];  // descriptors of B_hi / B_lo of network slot s; slots without a [32, 32] network: unused
  uint32_t mbar, tmem, phase;
};

template <class N>
constexpr bool tc_net() {
  return N::PRESENT && N::H1 == tc::TK && N::H2 == tc::TN;
}

template <class N>
SDE4_DEV void mlp_fwd_tc(const float* W, ScratchTC& sc, const float (&in)[N::IN], float (&out)[N::OUT], int slot) {
  using namespace tc;
  {
    float a1[N::H1];
    dense_in_regs<N::IN, N::H1, N::H1P>(W + N::OFF_W0, W + N::OFF_B0, in, a1);
#pragma unroll
    for (int c = 0; c < TK / 4; ++c) {
      float4 hi, lo;
      const float h0 = act_fwd<N::ACT>(a1[4 * c + 0]), h1 = act_fwd<N::ACT>(a1[4 * c + 1]);
      const float h2 = act_fwd<N::ACT>(a1[4 * c + 2]), h3 = act_fwd<N::ACT>(a1[4 * c + 3]);
      hi.x = tf32_hi(h0);
      hi.y = tf32_hi(h1);
      hi.z = tf32_hi(h2);
      hi.w = tf32_hi(h3);
      lo.x = h0 - hi.x;
      lo.y = h1 - hi.y;
      lo.z = h2 - hi.z;
      lo.w = h3 - hi.w;
      if (c == 0) __syncthreads();  // the A tiles double as the FP32 scratch columns of the other networks: every
                                    // thread is past its last column access before the first row is stored
      *reinterpret_cast<float4*>(sc.a_row + c * A_LBO) = hi;
      *reinterpret_cast<float4*>(sc.a_row + A_TILE + c * A_LBO) = lo;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;");           // (orders the previous tcgen05.ld of this thread too)
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint64_t dbh = sc.db[slot], dbl = sc.dbl[slot];
#pragma unroll
    for (int ks = 0; ks < TK / 8; ++ks) {  // the start address advances by two K chunks per step
      const uint64_t sa = (uint64_t)((2 * ks * A_LBO) >> 4), sb = (uint64_t)((2 * ks * B_LBO) >> 4);
      mma_tf32(sc.tmem, sc.da_hi + sa, dbh + sb, ks > 0 ? 1u : 0u);
      mma_tf32(sc.tmem, sc.da_lo + sa, dbh + sb, 1u);
      mma_tf32(sc.tmem, sc.da_hi + sa, dbl + sb, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sc.mbar) : "memory");
  }
  {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 24) && !done; ++spin) {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}\n"
          : "=r"(done)
          : "r"(sc.mbar), "r"(sc.phase)
          : "memory");
    }
    if (!done) __trap();  // (an error code for the caller instead of a hung device)
    sc.phase ^= 1u;
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t r[TN];
  const uint32_t taddr = sc.tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);  // lane = row of the tile
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  float a2[N::H2];
#pragma unroll
  for (int j = 0; j < N::H2; ++j) a2[j] = act_fwd<N::ACT>(__uint_as_float(r[j]) + W[N::OFF_B1 + j]);
  dense_out_t<N::H2, N::H2P, N::OUT>(W + N::OFF_W2, W + N::OFF_B2, a2, out);
}

// One thread per sample; forward evaluations of [32, 32] networks on the tensor cores, everything else as Ev1.
struct Ev1TC {
  static constexpr int L = 1;
  static constexpr bool PGRAD = false;
  using Ctx = ScratchTC;
  template <class N, bool KEEP>
  SDE4_DEV static void fwd(const NetRef& r, Ctx& c, const float (&in)[N::IN], float (&out)[N::OUT]) {
    if constexpr (tc_net<N>() && !KEEP) mlp_fwd_tc<N>(r.w, c, in, out, r.slot);
    else mlp_fwd<N, KEEP, false>(r.w, c, in, out, r.slot);
  }
  template <class N>
  SDE4_DEV static void bwd(const NetRef& r, Ctx& c, const float (&gout)[N::OUT], float (&gin)[N::IN]) {
    mlp_bwd<N, false>(r.w, c, gout, gin, r.slot);
  }
  template <class N>
  SDE4_DEV static void fwd_vjp(const NetRef& r, Ctx& c, const float (&in)[N::IN], const float (&gout)[N::OUT],
                               float (&gin)[N::IN]) {
    mlp_fwd_vjp<N, false>(r.w, c, in, gout, gin, r.slot);
  }
  template <class N, class GdFn>
  SDE4_DEV static void scalar_value_grad(const NetRef& r, Ctx& c, const float (&in)[N::IN], float& out,
                                         float (&J)[N::IN], GdFn gd_of) {
    mlp_scalar_value_grad<N, false>(r.w, c, in, out, J, gd_of, r.slot);
  }
};

template <class A>
constexpr int tc_nets() {
  return (tc_net<typename A::DNet>() ? 1 : 0) + (tc_net<typename A::NetA>() ? 1 : 0) + (tc_net<typename A::NetB>() ? 1 : 0);
}
template <class A>
constexpr int tc_smem_bytes() {
  // parameter image (the w1 blocks of the [32, 32] networks hold B_lo) | A_hi, A_lo (= the two FP32 scratch slots of the
  // forward evaluator, MAXW x 128 floats each; the uses never overlap in time, see mlp_fwd_tc) | B_hi of every [32, 32] network
  static_assert(Scratch::NSLOTS_FWD * MAXW * tc::TM * sizeof(float) == 2 * tc::A_TILE, "scratch columns = the two A tiles");
  return (int)(sizeof(float) * (size_t)A::smem_floats(1)) + 128 + 2 * (int)tc::A_TILE + tc_nets<A>() * (int)tc::B_TILE;
}

// weight tiles of one network slot: B[n][k] = w1[k][n] (haiku stores [in][out]), split into TF32 high / low parts.  B_hi
// goes to its own tile; B_lo REPLACES w1 in the parameter image (same 4 KB, the FP32 copy is dead in this kernel) -- the
// 4 KB saved per network are what lets a fifth tile be resident per SM.
template <class N>
SDE4_DEV void tc_build_b(float* Wnet, unsigned char* b_hi) {
  using namespace tc;
  if constexpr (tc_net<N>()) {
    static_assert(N::H2P == TN && (N::OFF_W1 % 4) == 0, "w1 is a dense, 16-byte aligned 32 x 32 block");
    constexpr int PER = TK * TN / TM;
    float v[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) v[q] = Wnet[N::OFF_W1 + threadIdx.x + q * TM];
    __syncthreads();  // every w1 value is in a register before the block is overwritten
    unsigned char* b_lo = reinterpret_cast<unsigned char*>(Wnet + N::OFF_W1);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int i = threadIdx.x + q * TM, k = i / TN, n = i - k * TN;
      const float hi = tf32_hi(v[q]);
      const uint32_t off = (uint32_t)(k >> 2) * B_LBO + (uint32_t)(n >> 3) * B_SBO + (uint32_t)(n & 7) * 16 + (uint32_t)(k & 3) * 4;
      *reinterpret_cast<float*>(b_hi + off) = hi;
      *reinterpret_cast<float*>(b_lo + off) = v[q] - hi;
    }
  }
}

// K1 on the tensor cores.  SIM: the simulator-summary instantiation (see k_rollout).
template <class M, bool SIM>
__global__ void __launch_bounds__(tc::TM) k_rollout_tc(const __grid_constant__ sde4_model_desc d,
                                                        const __grid_constant__ RolloutK a) {
  using A = typename M::A;
  constexpr int NX = M::NX, NU = M::NU;
  static_assert(M::Ev::L == 1, "one thread per sample");
  extern __shared__ __align__(16) float W[];
  __shared__ __align__(8) uint64_t mbar_params;
  __shared__ __align__(8) uint64_t mbar_mma;
  __shared__ uint32_t tmem_slot;
  stage_params_bulk(W, a.params, A::PARAM_TOTAL * sizeof(float), &mbar_params);
  Diffusion<M>::derive(W);
  unsigned char* tiles = reinterpret_cast<unsigned char*>(((uintptr_t)(W + A::smem_floats(1)) + 127) & ~(uintptr_t)127);
  unsigned char* a_hi = tiles;
  unsigned char* b0 = tiles + 2 * tc::A_TILE;
  __syncthreads();  // the parameter image is complete
  constexpr int BS1 = tc_net<typename A::DNet>() ? 1 : 0, BS2 = BS1 + (tc_net<typename A::NetA>() ? 1 : 0);  // compact slots
  tc_build_b<typename A::DNet>(W + A::OFF_DENSITY, b0);
  tc_build_b<typename A::NetA>(W + A::OFF_NET_A, b0 + BS1 * tc::B_TILE);
  tc_build_b<typename A::NetB>(W + A::OFF_NET_B, b0 + BS2 * tc::B_TILE);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(&mbar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {  // tensor memory: 32 columns = the 128 x 32 accumulator tile
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_slot)), "r"(32u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the weight tiles are read by the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  Tctx<M> tcx;
  tcx.l = 0;
  tcx.ev.base = reinterpret_cast<float*>(a_hi) + threadIdx.x;  // scratch columns alias the A tiles
  tcx.ev.bs = (int)blockDim.x;
  tcx.ev.a_row = a_hi + (threadIdx.x >> 3) * tc::A_SBO + (threadIdx.x & 7) * 16;
  tcx.ev.da_hi = tc::umma_desc(tc::smem_u32(a_hi), tc::A_LBO, tc::A_SBO);
  tcx.ev.da_lo = tc::umma_desc(tc::smem_u32(a_hi + tc::A_TILE), tc::A_LBO, tc::A_SBO);
  tcx.ev.db[0] = tc::umma_desc(tc::smem_u32(b0), tc::B_LBO, tc::B_SBO);
  tcx.ev.db[1] = tc::umma_desc(tc::smem_u32(b0 + BS1 * tc::B_TILE), tc::B_LBO, tc::B_SBO);
  tcx.ev.db[2] = tc::umma_desc(tc::smem_u32(b0 + BS2 * tc::B_TILE), tc::B_LBO, tc::B_SBO);
  tcx.ev.dbl[0] = tc::umma_desc(tc::smem_u32(W + A::OFF_DENSITY + A::DNet::OFF_W1), tc::B_LBO, tc::B_SBO);
  tcx.ev.dbl[1] = tc::umma_desc(tc::smem_u32(W + A::OFF_NET_A + A::NetA::OFF_W1), tc::B_LBO, tc::B_SBO);
  tcx.ev.dbl[2] = tc::umma_desc(tc::smem_u32(W + A::OFF_NET_B + A::NetB::OFF_W1), tc::B_LBO, tc::B_SBO);
  tcx.ev.mbar = tc::smem_u32(&mbar_mma);
  tcx.ev.tmem = tmem_slot;
  tcx.ev.phase = 0;

  const int g0 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = g0 < a.G;
  const int g = active ? g0 : a.G - 1;  // inactive threads shadow a valid sample: the CTA stays in lock step
  const int p = g / a.N, n = g - p * a.N;
  float x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = __ldg(a.y0 + (size_t)p * NX + i);
  const bool traj = !SIM || a.xevol != nullptr;
  if (traj) store_state<M>(a.xevol + g, a.G, x, 0, active);
  [[maybe_unused]] float racc = 0.0f;
  [[maybe_unused]] bool alive = true;
  const bool noisy = a.noise_mode != SDE4_NOISE_NONE;
  const uint32_t nzmask = noise_mask<NX>(d);
  uint32_t k0 = 0, k1 = 0;
  if (a.noise_mode == SDE4_NOISE_THREEFRY) {
    const uint32_t* k = a.key + (size_t)p * a.key_sp;
    brownian_key(__ldg(k), __ldg(k + 1), (uint32_t)(n + a.n_off), (uint32_t)a.n_tot, a.rng_mode, k0, k1);
  }
  const float* up = a.u + (size_t)p * a.u_sp;
  for (int t = 0; t < a.H; ++t) {
    float u[NU], dw[NX];
#pragma unroll
    for (int j = 0; j < NU; ++j) u[j] = __ldg(up + (size_t)t * a.u_st + (size_t)j * a.u_sj);
    if (noisy) draw_noise<M>(d, a.noise_mode, a.rng_mode, a.noise, k0, k1, t, a.H, (unsigned)a.G, g, nzmask, tcx, dw, nullptr);
    float aux;
    em_step<M>(W, tcx, d, x, u, dw, __ldg(a.dt + t), noisy, aux);
    if (traj) store_state<M>(a.xevol + (size_t)(t + 1) * a.x_rows * a.G + g, a.G, x, 0, active);
    if constexpr (SIM && NX == 4) {
      if (a.reward_kind == SDE4_REWARD_CARTPOLE_SWINGUP) {
        const float r = sim_reward_cartpole(x);
        racc += alive ? r : 0.0f;
        alive = alive && !sim_done_cartpole(x);
      }
    }
  }
  if constexpr (SIM) {
    if (active) {
#pragma unroll
      for (int i = 0; i < NX; ++i) a.last[(size_t)i * a.G + g] = x[i];
      if (a.racc) a.racc[g] = racc;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tcx.ev.tmem), "r"(32u));
}

}  // namespace sde4

// registry.cuh -- table of compiled architecture specialisations.  Each inst_*.cu translation
// unit instantiates the kernels of a few architectures and registers them at load time.
#pragma once
#include <type_traits>
#include <vector>
#include "kernels.cuh"
#include "reg_kernels.cuh"

namespace sde4 {

struct ArchEntry {
  const char* name;
  bool (*matches)(const sde4_model_desc&);
  int nx, nu, naux;
  int param_total;
  int offsets[6];
  int smem_bytes;
  int macs;
  // index = log2(lanes per sample): [0] one thread per sample, [2] 4 lanes, [3] 8 lanes; null = not compiled
  cudaError_t (*launch_rollout[5])(const sde4_model_desc&, const RolloutK&, int blocks, int threads, cudaStream_t);
  cudaError_t (*launch_cost[5])(const sde4_model_desc&, const sde4_cost_desc&, const sde4_cost_desc&, const CostK&,
                                bool grad, bool constr, int blocks, int threads, cudaStream_t);
  // training loss (value + input gradient + parameter-gradient streams); null = not compiled
  int n_phys;      // physical scalars whose gradients the training kernel forms (rotor: 16; packed offset offsets[0])
  int loss_lanes;  // lanes per sample of launch_loss_ls (0: none compiled)
  cudaError_t (*launch_loss_ls)(const sde4_model_desc&, const sde4_cost_desc&, const sde4_cost_desc&, const CostK&, int, int,
                                cudaStream_t);
  cudaError_t (*launch_loss)(const sde4_model_desc&, const sde4_cost_desc&, const sde4_cost_desc&, const CostK&,
                             int blocks, int threads, cudaStream_t);
  int net_dims[3][4];   // IN, H1, H2, OUT of density / net A / net B (0 = absent)
  int net_off[3];       // offsets of the three networks in the packed block
  int n_scaler, off_scaler;
  // distance-aware regulariser (reg_kernels.cuh); mu network of the compiled instance: D -> 8 -> 8 -> 1, tanh
  cudaError_t (*launch_reg)(const sde4_model_desc&, const RegK&, cudaStream_t);
  int act_density;
};

std::vector<ArchEntry>& arch_table();

// warp-specialised cost + gradient kernels (kernels_ws.cuh), registered per architecture from their own translation
// units; looked up by descriptor at dispatch time
struct WsEntry {
  const char* name;
  bool (*matches)(const sde4_model_desc&);
  int npart;  // particles per CTA = samples per cost / gradient partial (sizes the partial buffers)
  cudaError_t (*launch)(const sde4_model_desc&, const sde4_cost_desc&, const sde4_cost_desc&, const CostK&, float* ftape,
                        int dev, cudaStream_t);
};
std::vector<WsEntry>& ws_table();
struct WsRegistrar {
  explicit WsRegistrar(const WsEntry& e) { ws_table().push_back(e); }
};

// tensor-core forward rollout (kernels_tc.cuh): one thread per sample, CTA = one 128-sample tile
struct TcEntry {
  const char* name;
  bool (*matches)(const sde4_model_desc&);
  cudaError_t (*launch_rollout)(const sde4_model_desc&, const RolloutK&, int dev, cudaStream_t);
};
std::vector<TcEntry>& tc_table();
struct TcRegistrar {
  explicit TcRegistrar(const TcEntry& e) { tc_table().push_back(e); }
};

// cudaFuncSetAttribute is per device and per function: remember which devices a function has been configured on
// (bit mask, relaxed atomics: setting the attribute twice is harmless, skipping it is not)
struct DevOnce {
  unsigned long long done = 0;
  bool need(int dev) const { return dev < 0 || dev >= 64 || !((__atomic_load_n(&done, __ATOMIC_RELAXED) >> dev) & 1ull); }
  void mark(int dev) {
    if (dev >= 0 && dev < 64) __atomic_fetch_or(&done, 1ull << dev, __ATOMIC_RELAXED);
  }
};
inline int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
  return dev;
}

// dynamic shared memory: parameter image + derived constants + per-thread scratch columns
template <class M>
constexpr size_t smem_bytes(int threads, bool backward = true) {
  using A = typename M::A;
  constexpr int L = M::Ev::L;
  const size_t slots = backward ? Scratch::NSLOTS : Scratch::NSLOTS_FWD;
  return sizeof(float) * ((size_t)A::smem_floats(L) + (L == 1 ? slots * MAXW * threads : 0));
}

template <class M>
cudaError_t launch_rollout_t(const sde4_model_desc& d, const RolloutK& a, int blocks, int threads, cudaStream_t s) {
  using A = typename M::A;
  const size_t sm = smem_bytes<M>(threads, false);
  const int dev = current_device();
  static DevOnce attr_set;
  if (attr_set.need(dev)) {
    cudaFuncSetAttribute(k_rollout<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<M>(256, false));
    if constexpr (M::Ev::L > 1)
      cudaFuncSetAttribute(k_rollout<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<M>(256, false));
    attr_set.mark(dev);
  }
  if (a.last) {  // simulator summary (sde4_sim_rollout): its own instantiation, generic noise path
    static DevOnce attr_sim;
    if (attr_sim.need(dev)) {
      cudaFuncSetAttribute(k_rollout<M, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem_bytes<M>(256, false));
      attr_sim.mark(dev);
    }
    k_rollout<M, false, false, true><<<blocks, threads, sm, s>>>(d, a);
    return cudaGetLastError();
  }
  if constexpr (M::Ev::L > 1) {
    if (a.noise_mode == SDE4_NOISE_THREEFRY && a.solver == SDE4_SOLVER_EULER_MARUYAMA) {  // pipelined-noise instantiation
      k_rollout<M, true><<<blocks, threads, sm, s>>>(d, a);
      return cudaGetLastError();
    }
  }
  if (a.solver != SDE4_SOLVER_EULER_MARUYAMA) {
    static DevOnce attr_set2;
    if (attr_set2.need(dev)) {
      cudaFuncSetAttribute(k_rollout<M, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem_bytes<M>(256, false));
      attr_set2.mark(dev);
    }
    k_rollout<M, false, true><<<blocks, threads, sm, s>>>(d, a);
    return cudaGetLastError();
  }
  k_rollout<M><<<blocks, threads, sm, s>>>(d, a);
  return cudaGetLastError();
}

template <class M>
cudaError_t launch_cost_t(const sde4_model_desc& d, const sde4_cost_desc& cd, const sde4_cost_desc& td,
                          const CostK& a, bool grad, bool constr, int blocks, int threads, cudaStream_t s) {
  using A = typename M::A;
  const size_t sm = smem_bytes<M>(threads, grad);
  const int dev = current_device();
  static DevOnce attr_set;
  if (attr_set.need(dev)) {
    const int mx = (int)smem_bytes<M>(256);
    cudaFuncSetAttribute(k_cost<M, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_cost<M, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_cost<M, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(k_cost<M, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    attr_set.mark(dev);
  }
  if constexpr (M::Ev::L > 1) {
    if (a.noise_mode == SDE4_NOISE_THREEFRY && (!grad || a.dwbuf)) {
      // lane-split + in-kernel noise: the straight-line instantiations (kernels.cuh, MODE 1 / 2)
      // the unshaped cost that goes with the model (cost.cuh, PLAIN) has a branch-free instantiation
      bool plain = false;
      if constexpr (M::NX == 13) plain = cd.kind == SDE4_COST_ROTOR_TRACK && cd.sig_mask == 0 && !cd.use_res_sig;
      else if constexpr (M::NX == 4) plain = cd.kind == SDE4_COST_QUADRATIC && cd.feat == 1 && !cd.use_res_sig;
      else plain = cd.kind == SDE4_COST_QUADRATIC && cd.feat == 0 && !cd.use_res_sig;
      {
        if (plain) {
          if (grad) {
            if (constr)
              k_cost<M, true, true, 2><<<blocks, threads, sm, s>>>(d, cd, td, a);
            else
              k_cost<M, true, false, 2><<<blocks, threads, sm, s>>>(d, cd, td, a);
          } else {
            if (constr)
              k_cost<M, false, true, 2><<<blocks, threads, sm, s>>>(d, cd, td, a);
            else
              k_cost<M, false, false, 2><<<blocks, threads, sm, s>>>(d, cd, td, a);
          }
          return cudaGetLastError();
        }
      }
      if (grad) {
        if (constr)
          k_cost<M, true, true, 1><<<blocks, threads, sm, s>>>(d, cd, td, a);
        else
          k_cost<M, true, false, 1><<<blocks, threads, sm, s>>>(d, cd, td, a);
      } else {
        if (constr)
          k_cost<M, false, true, 1><<<blocks, threads, sm, s>>>(d, cd, td, a);
        else
          k_cost<M, false, false, 1><<<blocks, threads, sm, s>>>(d, cd, td, a);
      }
      return cudaGetLastError();
    }
  }
  if constexpr (M::Ev::L == 1) {
    // one thread per sample with the model's plain stage cost: the smaller instantiation (MODE 3)
    bool plain1 = false;
    if constexpr (M::NX == 13) plain1 = cd.kind == SDE4_COST_ROTOR_TRACK && cd.sig_mask == 0 && !cd.use_res_sig;
    else if constexpr (M::NX == 4) plain1 = cd.kind == SDE4_COST_QUADRATIC && cd.feat == 1 && !cd.use_res_sig;
    else plain1 = cd.kind == SDE4_COST_QUADRATIC && cd.feat == 0 && !cd.use_res_sig;
    if (plain1) {
      static DevOnce attr3;
      if (attr3.need(dev)) {
        const int mx = (int)smem_bytes<M>(256);
        cudaFuncSetAttribute(k_cost<M, true, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(k_cost<M, true, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(k_cost<M, false, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(k_cost<M, false, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        attr3.mark(dev);
      }
      if (grad) {
        if (constr)
          k_cost<M, true, true, 3><<<blocks, threads, sm, s>>>(d, cd, td, a);
        else
          k_cost<M, true, false, 3><<<blocks, threads, sm, s>>>(d, cd, td, a);
      } else {
        if (constr)
          k_cost<M, false, true, 3><<<blocks, threads, sm, s>>>(d, cd, td, a);
        else
          k_cost<M, false, false, 3><<<blocks, threads, sm, s>>>(d, cd, td, a);
      }
      return cudaGetLastError();
    }
  }
  if (grad) {
    if (constr)
      k_cost<M, true, true><<<blocks, threads, sm, s>>>(d, cd, td, a);
    else
      k_cost<M, true, false><<<blocks, threads, sm, s>>>(d, cd, td, a);
  } else {
    if (constr)
      k_cost<M, false, true><<<blocks, threads, sm, s>>>(d, cd, td, a);
    else
      k_cost<M, false, false><<<blocks, threads, sm, s>>>(d, cd, td, a);
  }
  return cudaGetLastError();
}

struct NoLaneSplit {};

template <class M>
cudaError_t launch_loss_t(const sde4_model_desc& d, const sde4_cost_desc& cd, const sde4_cost_desc& td, const CostK& a,
                          int blocks, int threads, cudaStream_t s) {
  const size_t sm = smem_bytes<M>(threads);
  if constexpr (M::Ev::L > 1) {
    if (a.noise_mode == SDE4_NOISE_THREEFRY) {  // straight-line instantiation (noise generated one step ahead)
      k_cost<M, true, false, 1><<<blocks, threads, sm, s>>>(d, cd, td, a);
      return cudaGetLastError();
    }
  }
  const int dev = current_device();
  static DevOnce attr_set;
  if (attr_set.need(dev)) {
    cudaFuncSetAttribute(k_cost<M, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<M>(256));
    attr_set.mark(dev);
  }
  k_cost<M, true, false><<<blocks, threads, sm, s>>>(d, cd, td, a);
  return cudaGetLastError();
}

// distance-aware regulariser: R1 centres -> R2 ball samples -> R3 scalars -> R4 second-order columns (reg_kernels.cuh)
template <class M, class MuNet>
cudaError_t launch_reg_t(const sde4_model_desc& d, const RegK& a, cudaStream_t s) {
  using C = RegCfg<M, MuNet>;
  using DNet = typename C::DNet;
  const int K0 = a.B * a.Hd;
  const long long KQ = (long long)K0 * a.N;
  const int dev = current_device();
  static DevOnce attr_set;
  const size_t sm1 = (size_t)(up4(C::W_FLOATS) + Scratch::NSLOTS * MAXW * 128) * 4;
  const int bs2 = a.ns <= 160 ? (160 / a.ns) * a.ns : 0;
  if (bs2 == 0) return cudaErrorInvalidValue;
  const size_t sm2 = (size_t)(up4(DNet::SIZE) + (3 * MAXW + C::D + 3) * bs2) * 4;
  const size_t sm4 = (size_t)(up4(C::W_FLOATS) + 4 * MAXW * 64) * 4;
  if (attr_set.need(dev)) {
    cudaFuncSetAttribute(k_reg_centre<M, MuNet>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
    cudaFuncSetAttribute(k_reg_ball<M, MuNet>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)((up4(DNet::SIZE) + (3 * MAXW + C::D + 3) * 160) * 4));
    cudaFuncSetAttribute(k_reg_second<M, MuNet>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4);
    attr_set.mark(dev);
  }
  k_reg_centre<M, MuNet><<<(K0 + 127) / 128, 128, sm1, s>>>(d, a);
  const int gpc = bs2 / a.ns;
  k_reg_ball<M, MuNet><<<(unsigned)((KQ + gpc - 1) / gpc), bs2, sm2, s>>>(d, a);
  k_reg_scalars<C::D><<<1, 1024, 0, s>>>(a);
  k_reg_second<M, MuNet><<<(K0 + 63) / 64, 64, sm4, s>>>(d, a);
  return cudaGetLastError();
}

constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v / 2); }

template <class MLS>
void add_lane_variant(ArchEntry& e) {
  if constexpr (!std::is_same<MLS, NoLaneSplit>::value) {
    e.launch_rollout[ilog2(MLS::Ev::L)] = &launch_rollout_t<MLS>;
    e.launch_cost[ilog2(MLS::Ev::L)] = &launch_cost_t<MLS>;
  }
}

template <class M, class MLS = NoLaneSplit, class MLS2 = NoLaneSplit, class MTRAIN = NoLaneSplit, class MLS3 = NoLaneSplit,
          class MTRAINLS = NoLaneSplit>
ArchEntry make_entry(const char* name) {
  using A = typename M::A;
  ArchEntry e;
  e.name = name;
  e.matches = &A::matches;
  e.nx = M::NX;
  e.nu = M::NU;
  e.naux = M::NAUX;
  e.param_total = A::PARAM_TOTAL;
  e.offsets[0] = A::OFF_SCALARS;
  e.offsets[1] = A::OFF_SCALER;
  e.offsets[2] = A::OFF_DENSITY;
  e.offsets[3] = A::OFF_NET_A;
  e.offsets[4] = A::OFF_NET_B;
  e.offsets[5] = A::PARAM_TOTAL;
  e.smem_bytes = (int)smem_bytes<M>(32);
  e.macs = A::MACS;
  for (int i = 0; i < 5; ++i) {
    e.launch_rollout[i] = nullptr;
    e.launch_cost[i] = nullptr;
  }
  e.launch_rollout[0] = &launch_rollout_t<M>;
  e.launch_cost[0] = &launch_cost_t<M>;
  add_lane_variant<MLS>(e);
  add_lane_variant<MLS2>(e);
  add_lane_variant<MLS3>(e);
  e.launch_loss = nullptr;
  if constexpr (!std::is_same<MTRAIN, NoLaneSplit>::value) e.launch_loss = &launch_loss_t<MTRAIN>;
  e.n_phys = 0;
  if constexpr (!std::is_same<MTRAIN, NoLaneSplit>::value) e.n_phys = MTRAIN::NPHYS;
  e.launch_loss_ls = nullptr;
  e.loss_lanes = 0;
  if constexpr (!std::is_same<MTRAINLS, NoLaneSplit>::value) {
    e.launch_loss_ls = &launch_loss_t<MTRAINLS>;
    e.loss_lanes = MTRAINLS::Ev::L;
  }
  using D = typename A::DNet;
  using NA = typename A::NetA;
  using NB = typename A::NetB;
  const int dims[3][4] = {{D::IN, D::H1, D::H2, D::OUT}, {NA::IN, NA::H1, NA::H2, NA::OUT}, {NB::IN, NB::H1, NB::H2, NB::OUT}};
  for (int s = 0; s < 3; ++s)
    for (int q = 0; q < 4; ++q) e.net_dims[s][q] = dims[s][q];
  e.net_off[0] = A::OFF_DENSITY;
  e.net_off[1] = A::OFF_NET_A;
  e.net_off[2] = A::OFF_NET_B;
  e.n_scaler = A::N_SCALER;
  e.off_scaler = A::OFF_SCALER;
  e.launch_reg = nullptr;
  e.act_density = D::ACT;
  if constexpr (!std::is_same<MTRAIN, NoLaneSplit>::value && A::HAS_DENSITY) {
    // every learn_mucoeff network of the reference's configurations is [8, 8] tanh (cartpole_sde.yaml:103-107, ...)
    e.launch_reg = &launch_reg_t<MTRAIN, Net<D::IN, 8, 8, 1, SDE4_ACT_TANH>>;
  }
  return e;
}

struct ArchRegistrar {
  explicit ArchRegistrar(const ArchEntry& e) { arch_table().push_back(e); }
};

#define SDE4_REGISTER(NAME, MODEL) static ::sde4::ArchRegistrar reg_##NAME(::sde4::make_entry<MODEL>(#NAME));
// also instantiate the lane-split kernels (LANES adjacent lanes per sample) for latency-bound launches
// one-thread-per-sample kernels + the training-loss kernel (parameter-gradient streams)
#define SDE4_REGISTER_TRAIN(NAME, MODEL_T, ARCH)                                                     \
  static ::sde4::ArchRegistrar reg_##NAME(::sde4::make_entry<MODEL_T<ARCH>, ::sde4::NoLaneSplit, ::sde4::NoLaneSplit, \
                                                              MODEL_T<ARCH, ::sde4::Ev1P>>(#NAME));
// one-thread-per-sample + LANES-lane kernels for rollout / cost AND for the training loss (latency-bound minibatches)
#define SDE4_REGISTER_TRAIN_LS(NAME, MODEL_T, ARCH, LANES)                                             \
  static ::sde4::ArchRegistrar reg_##NAME(                                                             \
      ::sde4::make_entry<MODEL_T<ARCH>, MODEL_T<ARCH, ::sde4::EvL<LANES>>, ::sde4::NoLaneSplit, MODEL_T<ARCH, ::sde4::Ev1P>, \
                         ::sde4::NoLaneSplit, MODEL_T<ARCH, ::sde4::EvLT<LANES, true>>>(#NAME));
#define SDE4_REGISTER_LS(NAME, MODEL_T, ARCH, LANES) \
  static ::sde4::ArchRegistrar reg_##NAME(::sde4::make_entry<MODEL_T<ARCH>, MODEL_T<ARCH, ::sde4::EvL<LANES>>>(#NAME));
#define SDE4_REGISTER_LS2(NAME, MODEL_T, ARCH, LANES_A, LANES_B)                                          \
  static ::sde4::ArchRegistrar reg_##NAME(                                                                \
      ::sde4::make_entry<MODEL_T<ARCH>, MODEL_T<ARCH, ::sde4::EvL<LANES_A>>, MODEL_T<ARCH, ::sde4::EvL<LANES_B>>>(#NAME));

#define SDE4_REGISTER_LS3(NAME, MODEL_T, ARCH, LANES_A, LANES_B, LANES_C)                                  \
  static ::sde4::ArchRegistrar reg_##NAME(                                                                \
      ::sde4::make_entry<MODEL_T<ARCH>, MODEL_T<ARCH, ::sde4::EvL<LANES_A>>, MODEL_T<ARCH, ::sde4::EvL<LANES_B>>, \
                         ::sde4::NoLaneSplit, MODEL_T<ARCH, ::sde4::EvL<LANES_C>>>(#NAME));

// two lane-split variants for rollout / cost + training kernels (one thread per sample and LT lanes)
#define SDE4_REGISTER_LS2_TRAIN(NAME, MODEL_T, ARCH, LANES_A, LANES_B, LT)                                 \
  static ::sde4::ArchRegistrar reg_##NAME(                                                                \
      ::sde4::make_entry<MODEL_T<ARCH>, MODEL_T<ARCH, ::sde4::EvL<LANES_A>>, MODEL_T<ARCH, ::sde4::EvL<LANES_B>>, \
                         MODEL_T<ARCH, ::sde4::Ev1P>, ::sde4::NoLaneSplit, MODEL_T<ARCH, ::sde4::EvLT<LT, true>>>(#NAME));

// shorthand for the architecture zoo
constexpr int TANH = SDE4_ACT_TANH, SWISH = SDE4_ACT_SWISH;
constexpr uint32_t ROTOR_NOISE_MASK = (7u << 3) | (7u << 10);  // states 3,4,5,10,11,12

}  // namespace sde4

// inst_tc.cu -- tensor-core forward rollouts (kernels_tc.cuh) of the architectures with [32, 32] networks that run in
// the throughput regime: the cartpole simulator model (sde4mbrlExamples/cartpole/cartpole_sde.yaml:60-63, BASELINE
// configs[4]) and the multirotor SDE with the 32-wide density network (hexa_ahg/optimizer_sde.yaml).
#include "kernels_tc.cuh"
#include "registry.cuh"
namespace sde4 {
using CartD32 = Net<3, 32, 32, 1, SWISH>;
using CartResSi = Net<3, 8, 24, 2, TANH>;
using CartResBb = Net<4, 8, 24, 2, TANH>;
using CartCtrl = Net<2, 6, 8, 2, TANH>;
using CartSiD32 = Arch<SDE4_MODEL_CARTPOLE, 4, 1, SDE4_CART_SIDE_INFO, CartD32, 0xE, 0xA, true, CartResSi, 0, CartCtrl, 0>;
using CartBbD32 = Arch<SDE4_MODEL_CARTPOLE, 4, 1, 0, CartD32, 0xE, 0xA, true, CartResBb, 0, NoNet, 0>;
using ResF = Net<3, 8, 16, 3, TANH>;
using ResM = Net<6, 8, 16, 3, TANH>;
using HexaD32s = Arch<SDE4_MODEL_ROTOR, 13, 6, SDE4_ROTOR_AERO_DRAG, Net<6, 32, 32, 1, SWISH>, ROTOR_NOISE_MASK, ROTOR_NOISE_MASK, true, ResF, 0x7, ResM, 0x3F>;

template <class M>
cudaError_t launch_rollout_tc_t(const sde4_model_desc& d, const RolloutK& a, int dev, cudaStream_t s) {
  using A = typename M::A;
  constexpr int sm = tc_smem_bytes<A>();
  static DevOnce once;
  if (once.need(dev)) {
    cudaFuncSetAttribute(k_rollout_tc<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaFuncSetAttribute(k_rollout_tc<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    once.mark(dev);
  }
  const int blocks = (a.G + tc::TM - 1) / tc::TM;
  if (a.last) k_rollout_tc<M, true><<<blocks, tc::TM, sm, s>>>(d, a);
  else k_rollout_tc<M, false><<<blocks, tc::TM, sm, s>>>(d, a);
  return cudaGetLastError();
}

#define SDE4_REGISTER_TC(NAME, MODEL, ARCH) \
  static TcRegistrar tcreg_##NAME(TcEntry{#NAME, &ARCH::matches, &launch_rollout_tc_t<MODEL<ARCH, Ev1TC>>});
SDE4_REGISTER_TC(cartpole_si_d32, CartpoleModel, CartSiD32)
SDE4_REGISTER_TC(cartpole_bb_d32, CartpoleModel, CartBbD32)
SDE4_REGISTER_TC(hexa_d32_swish, RotorModel, HexaD32s)
}  // namespace sde4

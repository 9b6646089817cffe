// inst_hexa_ws.cu -- warp-specialised cost + gradient kernels (kernels_ws.cuh) of the multirotor SDE architectures
// whose single-problem launches are bound by the serial instruction chain (BASELINE configs[2]).
#include <cstdlib>
#include "kernels_ws.cuh"
#include "registry.cuh"
namespace sde4 {
using ResF = Net<3, 8, 16, 3, TANH>;
using ResM = Net<6, 8, 16, 3, TANH>;
using HexaD32s = Arch<SDE4_MODEL_ROTOR, 13, 6, SDE4_ROTOR_AERO_DRAG, Net<6, 32, 32, 1, SWISH>, ROTOR_NOISE_MASK, ROTOR_NOISE_MASK, true, ResF, 0x7, ResM, 0x3F>;
using HexaD16t = Arch<SDE4_MODEL_ROTOR, 13, 6, SDE4_ROTOR_AERO_DRAG, Net<6, 16, 16, 1, TANH>, ROTOR_NOISE_MASK, ROTOR_NOISE_MASK, true, ResF, 0x7, ResM, 0x3F>;
using IrisD32s = Arch<SDE4_MODEL_ROTOR, 13, 4, SDE4_ROTOR_AERO_DRAG, Net<6, 32, 32, 1, SWISH>, ROTOR_NOISE_MASK, ROTOR_NOISE_MASK, true, ResF, 0x7, ResM, 0x3F>;

template <class A>
cudaError_t launch_cost_ws_t(const sde4_model_desc& d, const sde4_cost_desc& cd, const sde4_cost_desc& td, const CostK& a,
                             float* ftape, int dev, cudaStream_t s) {
  using Cfg = WsCfg<A>;
  const size_t sm = Cfg::smem_bytes();
  static DevOnce once;
  if (once.need(dev)) {
    cudaFuncSetAttribute(k_cost_ws<A, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k_cost_ws<A, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k_cost_ws<A, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k_cost_ws<A, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k_cost_ws<A, true, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k_cost_ws<A, false, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    once.mark(dev);
  }
  const int blocks = (a.G + Cfg::NPART - 1) / Cfg::NPART;
  const bool plain = cd.kind == SDE4_COST_ROTOR_TRACK && cd.sig_mask == 0 && !cd.use_res_sig;
  int sms = 148;
  if (dev < 0 || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  static const int phymb_max = [] {
    // CTAs per SM up to which the latency-regime kernel runs: one wave of up to three CTAs per SM is still bound by the chain of
    // a step (speculative APG launches of 5-7 k samples: 2.80 -> 2.69 ms and 2.88 -> 2.76 ms per 20 iterations against the
    // many-CTAs instantiation); from the second wave on the busiest role binds.  SDE4_WS_PHYMB_MAX overrides (tuning aid).
    const char* e = std::getenv("SDE4_WS_PHYMB_MAX");
    return e ? std::atoi(e) : 3;
  }();
  const bool phymb = blocks <= phymb_max * sms;  // latency regime (see k_cost_ws)
  // four CTAs per SM where that saves a wave of CTAs (up to two waves; beyond, three per SM have the higher throughput)
  static const int force_minb = [] {
    const char* e = std::getenv("SDE4_WS_MINB");  // tuning aid: 3 / 4 forces the instantiation
    return e ? std::atoi(e) : 0;
  }();
  const int w3 = (blocks + 3 * sms - 1) / (3 * sms), w4 = (blocks + 4 * sms - 1) / (4 * sms);
  const bool four = !phymb && (force_minb ? force_minb == 4 : (w4 < w3 && w4 <= 2));
  if (plain) {
    if (phymb) k_cost_ws<A, true, true><<<blocks, Cfg::THREADS, sm, s>>>(d, cd, td, a, ftape);
    else if (four) k_cost_ws<A, true, false, 4><<<blocks, Cfg::THREADS, sm, s>>>(d, cd, td, a, ftape);
    else k_cost_ws<A, true, false><<<blocks, Cfg::THREADS, sm, s>>>(d, cd, td, a, ftape);
  } else {
    if (phymb) k_cost_ws<A, false, true><<<blocks, Cfg::THREADS, sm, s>>>(d, cd, td, a, ftape);
    else if (four) k_cost_ws<A, false, false, 4><<<blocks, Cfg::THREADS, sm, s>>>(d, cd, td, a, ftape);
    else k_cost_ws<A, false, false><<<blocks, Cfg::THREADS, sm, s>>>(d, cd, td, a, ftape);
  }
  return cudaGetLastError();
}

#define SDE4_REGISTER_WS(NAME, ARCH) \
  static WsRegistrar wsreg_##NAME(WsEntry{#NAME, &ARCH::matches, WsCfg<ARCH>::NPART, &launch_cost_ws_t<ARCH>});
SDE4_REGISTER_WS(hexa_d32_swish, HexaD32s)
}  // namespace sde4
#ifdef SDE4_WS_TRACE
extern "C" int sde4_debug_ws_trace(long long* out /*[4][1024]*/, int* counts /*[4]*/) {
  int zero[4] = {0, 0, 0, 0};
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, sde4::g_ws_trace, sizeof(long long) * 4 * 1024);
  cudaMemcpyFromSymbol(counts, sde4::g_ws_trace_n, sizeof(int) * 4);
  cudaMemcpyToSymbol(sde4::g_ws_trace_n, zero, sizeof(zero));
  return 0;
}
#endif

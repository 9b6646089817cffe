// exp_tcgen05.cu -- the tensor-core experiment of the hidden->hidden product, as a measured micro-kernel.
//
// Question (profiles/r2_tcgen05_experiment.md): is the 32 x 32 hidden->hidden product of the reference's MLPs
// (hk.nets.MLP, hidden_layers [32, 32]: sde4mbrl/nsde.py:387, sde_rotor_model.py:400-409) faster on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulator in tensor memory) than on the FP32 pipe, at float32 parity?
// Both kernels evaluate a CHAIN of `layers` layers  x <- act(x W + b)  (swish) for tiles of 128 samples, one thread
// per sample, which is the dependency structure of a rollout step (the next product needs the activation of this one).
//
//   variant 0  FP32: weights in shared memory (warp-uniform LDS.128), 32 accumulators per thread, 1024 FFMA per layer.
//   variant 1  tcgen05: per layer the 128 threads split their activations into TF32 high / low parts
//              (x_hi = x & 0xffffe000, x_lo = x - x_hi), store them to shared memory in the K-major no-swizzle
//              core-matrix layout, one thread issues 12 x tcgen05.mma.cta_group::1.kind::tf32 (M 128, N 32, K 8; four
//              K steps x {hi*hi, lo*hi, hi*lo} = 3xTF32), tcgen05.commit -> mbarrier, and every thread reads its own
//              row of the accumulator back with one tcgen05.ld.32x32b.x32.
//
// Debug / measurement entry points only (not declared in include/sde4b200.h, not used by the product path).
#include <cuda_runtime.h>
#include <cstdint>

namespace {

constexpr int TM = 128, TK = 32, TN = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float fast_sigmoid(float a) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * -1.44269504088896341f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float swish(float a) { return a * fast_sigmoid(a); }

// ---- variant 0: FP32 FMA ------------------------------------------------------------------------------------------
// (the product's one-thread-per-sample evaluator, nets.cuh dense_in_smem: activations in a per-thread column of shared
//  memory, the loop over the input rows ROLLED, 32 accumulators in registers)
__global__ void __launch_bounds__(TM) k_layer_chain_fp32(const float* __restrict__ w, const float* __restrict__ b,
                                                         const float* __restrict__ x, float* __restrict__ y, int layers) {
  __shared__ __align__(16) float sw[TK * TN + TN];
  __shared__ float col[TK * TM];  // element k of thread t at col[k*TM + t]
  for (int i = threadIdx.x; i < TK * TN; i += TM) sw[i] = w[i];
  if (threadIdx.x < TN) sw[TK * TN + threadIdx.x] = b[threadIdx.x];
  const size_t row = (size_t)blockIdx.x * TM + threadIdx.x;
  float* mine = col + threadIdx.x;
#pragma unroll
  for (int k = 0; k < TK; ++k) mine[k * TM] = x[row * TK + k];
  __syncthreads();
  float acc[TN];
  for (int l = 0; l < layers; ++l) {
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[j] = sw[TK * TN + j];
#pragma unroll 4
    for (int k = 0; k < TK; ++k) {
      const float a = mine[k * TM];
      const float4* wr = reinterpret_cast<const float4*>(sw + k * TN);
#pragma unroll
      for (int j4 = 0; j4 < TN / 4; ++j4) {
        const float4 q = wr[j4];
        acc[4 * j4 + 0] = fmaf(a, q.x, acc[4 * j4 + 0]);
        acc[4 * j4 + 1] = fmaf(a, q.y, acc[4 * j4 + 1]);
        acc[4 * j4 + 2] = fmaf(a, q.z, acc[4 * j4 + 2]);
        acc[4 * j4 + 3] = fmaf(a, q.w, acc[4 * j4 + 3]);
      }
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      acc[j] = swish(acc[j]);
      mine[j * TM] = acc[j];
    }
  }
#pragma unroll
  for (int k = 0; k < TK; ++k) y[row * TK + k] = acc[k];
}

// ---- variant 1: tcgen05 ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (K-major, no swizzle): core matrix = 8 rows x 16 bytes, rows 16 bytes apart;
// LBO = byte distance of the two 16-byte K chunks of one instruction (K = 8 tf32 = 32 bytes), SBO = byte distance of
// consecutive 8-row groups.  Bits: [0,14) address >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4, [46,48) version = 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor: D = F32 (bits [4,6) = 1), A = B = TF32 (bits [7,10) = [10,13) = 2), both K-major, N >> 3 at
// [17,23), M >> 4 at [24,29).
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

constexpr uint32_t A_LBO = (TM / 8) * 128, A_SBO = 128, A_TILE = TM * TK * 4;  // 2048, 128, 16 KB
constexpr uint32_t B_LBO = (TN / 8) * 128, B_SBO = 128, B_TILE = TN * TK * 4;  // 512, 128, 4 KB

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(TM) k_layer_chain_tc(const float* __restrict__ w, const float* __restrict__ b,
                                                       const float* __restrict__ x, float* __restrict__ y, int layers,
                                                       int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_hi = smem;
  unsigned char* a_lo = smem + A_TILE;
  unsigned char* b_hi = smem + 2 * A_TILE;
  unsigned char* b_lo = b_hi + B_TILE;
  float* sb = reinterpret_cast<float*>(b_lo + B_TILE);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sb + TN);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int t = threadIdx.x, warp = t >> 5;

  // B[n][k] = w[k][n] (hk stores [in][out]); element (n, k) -> chunk k/4, row group n/8, row n%8, k%4 inside 16 bytes
  for (int i = t; i < TK * TN; i += TM) {
    const int k = i / TN, n = i - k * TN;
    const float v = w[i];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    const uint32_t off = (uint32_t)(k >> 2) * B_LBO + (uint32_t)(n >> 3) * B_SBO + (uint32_t)(n & 7) * 16 + (uint32_t)(k & 3) * 4;
    *reinterpret_cast<float*>(b_hi + off) = hi;
    *reinterpret_cast<float*>(b_lo + off) = v - hi;
  }
  if (t < TN) sb[t] = b[t];
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {  // tensor memory: 32 columns (128 lanes x 32 x 4 bytes = the accumulator tile)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = *tmem_slot;

  const size_t row = (size_t)blockIdx.x * TM + t;
  float v[TK];
#pragma unroll
  for (int k = 0; k < TK; ++k) v[k] = x[row * TK + k];
  const uint32_t a_off = (uint32_t)(t >> 3) * A_SBO + (uint32_t)(t & 7) * 16;
  const uint64_t da_hi = umma_desc(smem_u32(a_hi), A_LBO, A_SBO), da_lo = umma_desc(smem_u32(a_lo), A_LBO, A_SBO);
  const uint64_t db_hi = umma_desc(smem_u32(b_hi), B_LBO, B_SBO), db_lo = umma_desc(smem_u32(b_lo), B_LBO, B_SBO);
  uint32_t phase = 0;
  for (int l = 0; l < layers; ++l) {
#pragma unroll
    for (int c = 0; c < TK / 4; ++c) {  // 8 chunks of 4 values: one 16-byte store each for the high and the low part
      float4 hi, lo;
      hi.x = __uint_as_float(__float_as_uint(v[4 * c + 0]) & 0xffffe000u);
      hi.y = __uint_as_float(__float_as_uint(v[4 * c + 1]) & 0xffffe000u);
      hi.z = __uint_as_float(__float_as_uint(v[4 * c + 2]) & 0xffffe000u);
      hi.w = __uint_as_float(__float_as_uint(v[4 * c + 3]) & 0xffffe000u);
      lo.x = v[4 * c + 0] - hi.x;
      lo.y = v[4 * c + 1] - hi.y;
      lo.z = v[4 * c + 2] - hi.z;
      lo.w = v[4 * c + 3] - hi.w;
      *reinterpret_cast<float4*>(a_hi + c * A_LBO + a_off) = hi;
      *reinterpret_cast<float4*>(a_lo + c * A_LBO + a_off) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;");           // (orders the previous layer's tcgen05.ld too)
    __syncthreads();
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
      for (int ks = 0; ks < TK / 8; ++ks) {  // start address advances by two K chunks per step
        const uint64_t sa = (uint64_t)((2 * ks * A_LBO) >> 4), sbb = (uint64_t)((2 * ks * B_LBO) >> 4);
        mma_tf32(tmem_d, da_hi + sa, db_hi + sbb, ks > 0 ? 1u : 0u);
        mma_tf32(tmem_d, da_lo + sa, db_hi + sbb, 1u);
        mma_tf32(tmem_d, da_hi + sa, db_lo + sbb, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    }
    {  // wait for the accumulator (bounded: a wrong descriptor must not hang the device)
      uint32_t done = 0;
      for (int spin = 0; spin < (1 << 20) && !done; ++spin) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(mbar)), "r"(phase)
            : "memory");
      }
      if (!done) {  // report and abort the launch (an error code, not a hung device)
        if (status) atomicExch(status, 1);
        __trap();
      }
      phase ^= 1u;
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t r[TN];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);  // lane = row of the tile, column = output unit
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
#pragma unroll
    for (int j = 0; j < TN; ++j) v[j] = swish(__uint_as_float(r[j]) + sb[j]);
  }
#pragma unroll
  for (int k = 0; k < TK; ++k) y[row * TK + k] = v[k];
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(32u));
}

}  // namespace

// variant 0 FP32, 1 tcgen05.  x, y: device float[tiles*128][32]; w: device float[32][32] ([in][out]); b: device float[32].
// ms (host, may be null): average duration of `reps` launches (CUDA events on `stream`).  Returns 0, a CUDA error
// code, or -2 when the tcgen05 kernel gave up waiting for its accumulator.
extern "C" int sde4_debug_layer_chain(int variant, int tiles, int layers, int reps, const float* w, const float* b,
                                      const float* x, float* y, float* ms, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const size_t sm = 2 * A_TILE + 2 * B_TILE + TN * 4 + 8 + 16;
  int* status = nullptr;
  if (variant == 1) {
    if (cudaFuncSetAttribute(k_layer_chain_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return (int)cudaGetLastError();
    if (cudaMalloc(&status, 4) != cudaSuccess) return (int)cudaGetLastError();
    cudaMemsetAsync(status, 0, 4, s);
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto launch = [&]() {
    if (variant == 0) k_layer_chain_fp32<<<tiles, TM, 0, s>>>(w, b, x, y, layers);
    else k_layer_chain_tc<<<tiles, TM, sm, s>>>(w, b, x, y, layers, status);
  };
  launch();  // warm-up
  cudaEventRecord(e0, s);
  for (int r = 0; r < reps; ++r) launch();
  cudaEventRecord(e1, s);
  cudaError_t err = cudaStreamSynchronize(s);
  float t = 0.0f;
  cudaEventElapsedTime(&t, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  int rc = 0;
  if (err != cudaSuccess) rc = (int)err;
  if (variant == 1) {
    int h = 0;
    if (rc == 0 && cudaMemcpy(&h, status, 4, cudaMemcpyDeviceToHost) == cudaSuccess && h) rc = -2;
    cudaFree(status);
  }
  if (ms) *ms = reps > 0 ? t / reps : 0.0f;
  return rc;
}

// FFMA2 (fma.rn.f32x2, sm_100a) vs FFMA: throughput (many independent chains, full occupancy) and dependent-chain
// latency (one warp per SM, one chain).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_bench scripts/ffma2_bench.cu && /tmp/ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int CH, bool PACKED>
__global__ void __launch_bounds__(1024) k_tp(int iters, float* sink) {
  const float a = 1.000001f, b = 1e-7f;
  if (PACKED) {
    u64 acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = pk(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    const u64 aa = pk(a, a), bb = pk(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < CH; ++i) acc[i] = fma2(acc[i], aa, bb);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) { float x, y; upk(acc[i], x, y); s += x + y; }
    if (s == 123.456f) *sink = s;
  } else {
    float acc[2 * CH];
#pragma unroll
    for (int i = 0; i < 2 * CH; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 2 * CH; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 2 * CH; ++i) s += acc[i];
    if (s == 123.456f) *sink = s;
  }
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  float* sink; cudaMalloc(&sink, 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  {  // throughput: 2 blocks of 1024 threads per SM, 8 chains (16 FMAs per iteration per thread either way)
    float m0 = time_ms([&] { k_tp<8, false><<<sms * 2, 1024>>>(iters, sink); });
    float m1 = time_ms([&] { k_tp<8, true><<<sms * 2, 1024>>>(iters, sink); });
    double fl = 2.0 * 16 * iters * 1024.0 * sms * 2;
    printf("throughput  FFMA %.1f TFLOP/s   FFMA2 %.1f TFLOP/s\n", fl / m0 * 1e-9, fl / m1 * 1e-9);
  }
  {  // one warp per SM: issue-limited single-warp rate with 8 independent chains, and the 1-chain dependent latency
    float m0 = time_ms([&] { k_tp<8, false><<<sms, 32>>>(iters, sink); });
    float m1 = time_ms([&] { k_tp<8, true><<<sms, 32>>>(iters, sink); });
    printf("one warp/SM, 16 FMAs/iter: FFMA %.2f cycles/iter (16 instr)   FFMA2 %.2f cycles/iter (8 instr)\n",
           m0 * 1e-3 * clk * 1e3 / iters, m1 * 1e-3 * clk * 1e3 / iters);
    float l0 = time_ms([&] { k_tp<1, false><<<sms, 32>>>(iters, sink); });
    float l1 = time_ms([&] { k_tp<1, true><<<sms, 32>>>(iters, sink); });
    printf("dependent chain: FFMA (2 chains) %.2f cycles/iter   FFMA2 (1 chain) %.2f cycles/iter\n",
           l0 * 1e-3 * clk * 1e3 / iters, l1 * 1e-3 * clk * 1e3 / iters);
  }
  return 0;
}

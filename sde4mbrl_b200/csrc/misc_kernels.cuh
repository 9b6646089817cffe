// misc_kernels.cuh -- non-template kernels (K3 reduction, K0 RNG self-test, FMA peak);
// included by sde4b200.cu only.
#pragma once
#include "kernels.cuh"

namespace sde4 {

// ----------------------------------------------------------------------------
// K3: fixed-order reduction of the partials of each problem, 1/N, + control-slew term.
// item <-> (row = t*n_opt + j, p), p fastest.  Row index H*n_opt is the cost.  a.lpi lanes work on
// one item: 1 (many problems, few partials each) or 32 (one warp per item: lane-strided, coalesced
// partial sums in a fixed order, then a butterfly) -- deterministic either way.
// ----------------------------------------------------------------------------
SDE4_DEV float reduce_lanes(float v, int lpi) {
  if (lpi == 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;
}

SDE4_DEV float strided_sum(const float* __restrict__ src, int n, int sub, int lpi) {
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
  int k = sub;
  for (; k + 3 * lpi < n; k += 4 * lpi) {
    s0 += src[k];
    s1 += src[k + lpi];
    s2 += src[k + 2 * lpi];
    s3 += src[k + 3 * lpi];
  }
  for (; k < n; k += lpi) s0 += src[k];
  return (s0 + s1) + (s2 + s3);
}

SDE4_DEV void k_reduce_body(const ReduceK& a);

// ---- cross-GPU tail of K3 (sde4_xgpu_reduce): run by the LAST CTA of the grid -------------------------------------------
SDE4_DEV void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
SDE4_DEV unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
SDE4_DEV float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
SDE4_DEV void xgpu_tail(const XgpuK& x) {
  // every CTA's stores to this rank's buffer were ordered before its ticket (threadfence + atomic); this CTA took the
  // last ticket, so after the fence below they are all visible at system scope before the flags go out
  if (threadIdx.x < x.world && (int)threadIdx.x != x.rank) {
    __threadfence_system();
    st_release_sys(x.flag[threadIdx.x] + x.rank, x.epoch);               // "rank's partial of this epoch is in place"
    while ((int)(ld_acquire_sys(x.flag[x.rank] + threadIdx.x) - x.epoch) < 0) {  // peer's partial is in place
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < x.n; i += blockDim.x) {
    float s = 0.0f;
    for (int r = 0; r < x.world; ++r) s += ld_relaxed_sys(x.buf[r] + i);  // rank order: identical bits on every rank
    x.out[i] = s;
  }
}
SDE4_DEV void xgpu_ticket(const XgpuK& x) {
  if (x.world <= 1) return;
  __shared__ unsigned last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(x.counter, 1u);
    last = t == gridDim.x - 1 ? 1u : 0u;
    if (last) *x.counter = 0u;  // ready for the next launch (stream ordered)
  }
  __syncthreads();
  if (last) {
    __threadfence();
    xgpu_tail(x);
  }
}

__global__ void k_reduce(const __grid_constant__ ReduceK a) {
  // programmatic dependent launch: when launched with the serialization attribute behind a kernel that calls
  // griddepcontrol.launch_dependents (k_cost_ws), this grid is already resident when the cost kernel's last CTA retires
  // and only waits here for its memory to be visible; a no-op for ordinary launches
  asm volatile("griddepcontrol.wait;" ::: "memory");
  k_reduce_body(a);
  xgpu_ticket(a.x);
}

SDE4_DEV void k_reduce_body(const ReduceK& a) {
  const int lpi = a.lpi;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long idx = gid / lpi;
  const int sub = (int)(gid - idx * lpi);
  const int rows = a.grad ? a.H * a.n_opt + 1 : 1;  // value only: just the cost row
  if (idx >= (long long)rows * a.P) return;  // warp-uniform for lpi = 32 (blockDim is a multiple of 32)
  const int p = (int)(idx % a.P);
  const int row = a.grad ? (int)(idx / a.P) : a.H * a.n_opt;
  if (a.mask && a.mask[p] == 0) return;  // skipped problem (uniform over the lanes of an item)
  const float invN = 1.0f / (float)a.N;
  const float* op = a.opt + (size_t)p * a.o_sp;
  if (row == a.H * a.n_opt) {
    float s = reduce_lanes(strided_sum(a.cpart + (size_t)p * a.NP, a.NP, sub, lpi), lpi) * invN;
    if (a.has_slew) {  // sum_t sum_j coeff_j ((u_{t+1,j}-u_{t,j})/dt_t)^2 * res_mult
      float sl = 0.0f;
      for (int t = sub; t + 1 < a.H; t += lpi) {
        const float idt = 1.0f / a.dt[t];
        for (int j = 0; j < a.n_u; ++j) {
          const float dv = (op[(size_t)(t + 1) * a.o_st + (size_t)j * a.o_sj] - op[(size_t)t * a.o_st + (size_t)j * a.o_sj]) * idt;
          sl = fmaf(a.slew[j] * dv, dv, sl);
        }
      }
      s = fmaf(reduce_lanes(sl, lpi), a.res_mult, s);
    }
    if (sub == 0) a.cost[p] = s;
    return;
  }
  if (a.grad == nullptr) return;
  const int t = row / a.n_opt, j = row - t * a.n_opt;
  float s = reduce_lanes(strided_sum(a.gpart + (size_t)row * a.PW + (size_t)p * a.NP, a.NP, sub, lpi), lpi) * invN;
  if (sub != 0) return;
  if (a.has_slew && j < a.n_u) {
    const float uj = op[(size_t)t * a.o_st + (size_t)j * a.o_sj];
    float gsl = 0.0f;
    if (t > 0) {
      const float idt = 1.0f / a.dt[t - 1];
      gsl += 2.0f * a.slew[j] * (uj - op[(size_t)(t - 1) * a.o_st + (size_t)j * a.o_sj]) * idt * idt;
    }
    if (t + 1 < a.H) {
      const float idt = 1.0f / a.dt[t];
      gsl -= 2.0f * a.slew[j] * (op[(size_t)(t + 1) * a.o_st + (size_t)j * a.o_sj] - uj) * idt * idt;
    }
    s = fmaf(gsl, a.res_mult, s);
  }
  a.grad[(size_t)p * a.o_sp + (size_t)t * a.o_st + (size_t)j * a.o_sj] = s;
}

// ----------------------------------------------------------------------------
// simulator summary: mean / population std over the particles of the last state, mean accumulated reward, reward and
// termination of the mean state (cartpole_sde.py:219-223, ModelEnv.step / evaluate_action_sequences).
// One thread per (state i, problem p), p fastest; fixed summation order.
// ----------------------------------------------------------------------------
struct SimReduceK {
  const float* last;   // [n_x][G]
  const float* racc;   // [G] or null
  const float* u_last; // control of the last step, element (p, j) at u_last[p*u_sp + j*u_sj]
  long long u_sp, u_sj;
  float* mean;
  float* std;
  float* reward_mean;
  float* reward_of_mean;
  int32_t* done_of_mean;
  int P, N, n_x, reward_kind;
};
__global__ void k_sim_reduce(const __grid_constant__ SimReduceK a) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)a.P * a.n_x) return;
  const int p = (int)(gid % a.P), i = (int)(gid / a.P);
  const size_t G = (size_t)a.P * a.N;
  const float* src = a.last + (size_t)i * G + (size_t)p * a.N;
  float m = 0.0f;
  for (int n = 0; n < a.N; ++n) m += src[n];
  m /= (float)a.N;
  if (a.mean) a.mean[(size_t)p * a.n_x + i] = m;
  if (a.std) {
    float v = 0.0f;
    for (int n = 0; n < a.N; ++n) {
      const float dlt = src[n] - m;
      v = fmaf(dlt, dlt, v);
    }
    a.std[(size_t)p * a.n_x + i] = sqrtf(v / (float)a.N);
  }
  if (i != 0) return;
  if (a.reward_mean && a.racc) {
    float r = 0.0f;
    for (int n = 0; n < a.N; ++n) r += a.racc[(size_t)p * a.N + n];
    a.reward_mean[p] = r / (float)a.N;
  }
  if ((a.reward_of_mean || a.done_of_mean) && a.reward_kind == SDE4_REWARD_CARTPOLE_SWINGUP && a.n_x == 4) {
    float xm[4];
    for (int k = 0; k < 4; ++k) {
      const float* s2 = a.last + (size_t)k * G + (size_t)p * a.N;
      float mk = 0.0f;
      for (int n = 0; n < a.N; ++n) mk += s2[n];
      xm[k] = mk / (float)a.N;
    }
    if (a.reward_of_mean) a.reward_of_mean[p] = sim_reward_cartpole(xm);
    if (a.done_of_mean) a.done_of_mean[p] = sim_done_cartpole(xm) ? 1 : 0;
  }
}

// ----------------------------------------------------------------------------
// K0: RNG self-test kernels
// ----------------------------------------------------------------------------
__global__ void k_rng_bits(uint32_t k0, uint32_t k1, unsigned long long total, unsigned long long offset,
                           unsigned long long count, int mode, uint32_t* out, float* outf) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t b = jax_bits(k0, k1, offset + i, total, mode);
  if (out) out[i] = b;
  if (outf) outf[i] = bits_to_normal(b);
}

__global__ void k_rng_split(const uint32_t* key, int num, int mode, uint32_t* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num) return;
  uint32_t c0, c1;
  jax_split(key[0], key[1], (uint32_t)i, (uint32_t)num, mode, c0, c1);
  out[2 * i] = c0;
  out[2 * i + 1] = c1;
}

// batched forms (training-loss key chains, nsde.py:615,709,781): K keys at once
__global__ void k_rng_split_many(const uint32_t* __restrict__ keys, unsigned long long K, int num, int mode,
                                 uint32_t* __restrict__ out) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * (unsigned long long)num) return;
  const unsigned long long k = i / num;
  const uint32_t j = (uint32_t)(i - k * num);
  uint32_t c0, c1;
  jax_split(keys[2 * k], keys[2 * k + 1], j, (uint32_t)num, mode, c0, c1);
  out[2 * i] = c0;
  out[2 * i + 1] = c1;
}

__global__ void k_rng_normal_many(const uint32_t* __restrict__ keys, unsigned long long K, unsigned long long n, int mode,
                                  float* __restrict__ out) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * n) return;
  const unsigned long long k = i / n;
  out[i] = bits_to_normal(jax_bits(keys[2 * k], keys[2 * k + 1], i - k * n, n, mode));
}

// jax.random.choice(key, n, (m,), replace=False) as a 0/1 indicator over [0, n): permutation(key, n) is one round of
// a stable sort of arange(n) by random_bits(split(key)[1], 32, (n,)); index t is drawn iff its rank is below m.
// One thread per (t, key); n is a training horizon (tens), so the O(n) rank count per thread is noise.
__global__ void k_rng_choice_mask(const uint32_t* __restrict__ keys, unsigned long long K, int n, int m, int mode,
                                  float* __restrict__ out) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * (unsigned long long)n) return;
  const unsigned long long k = i % K;
  const int t = (int)(i / K);
  uint32_t s0, s1;
  jax_split(keys[2 * k], keys[2 * k + 1], 1, 2, mode, s0, s1);
  const uint32_t mine = jax_bits(s0, s1, t, n, mode);
  int rank = 0;
  for (int j = 0; j < n; ++j) {
    const uint32_t b = jax_bits(s0, s1, j, n, mode);
    rank += (b < mine || (b == mine && j < t)) ? 1 : 0;
  }
  out[i] = rank < m ? 1.0f : 0.0f;
}

// jax.random.randint(key, (n,), minval, minval + span) for int32 (sde4mbrl/nsde.py:65, :1308 -- the 'random'
// u_sampling_strategy): k1, k2 = split(key); hi = random_bits(k1), lo = random_bits(k2);
// offset = ((hi % span) * (2^32 % span) + lo % span) % span in wrapping uint32 arithmetic, with 2^32 % span formed as
// ((2^16 % span)^2) % span.  One thread per (key, element).
__global__ void k_rng_randint_many(const uint32_t* __restrict__ keys, unsigned long long K, unsigned long long n,
                                   int minval, uint32_t span, int mode, int32_t* __restrict__ out) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * n) return;
  const unsigned long long k = i / n;
  const unsigned long long j = i - k * n;
  uint32_t a0, a1, b0, b1;
  jax_split(keys[2 * k], keys[2 * k + 1], 0u, 2u, mode, a0, a1);
  jax_split(keys[2 * k], keys[2 * k + 1], 1u, 2u, mode, b0, b1);
  const uint32_t hi = jax_bits(a0, a1, j, n, mode);
  const uint32_t lo = jax_bits(b0, b1, j, n, mode);
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  const uint32_t off = ((hi % span) * mult + lo % span) % span;
  out[i] = minval + (int32_t)off;
}

// ----------------------------------------------------------------------------
// FP32 FMA peak micro-benchmark (roofline denominator)
// ----------------------------------------------------------------------------
template <int VARIANT>
__global__ void __launch_bounds__(1024) k_fma_peak(int iters, float* sink) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i) * 1e-3f;
  const float a = 1.000001f, b = 1e-7f;
  if (VARIANT == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        asm volatile(
            "{\n"
            ".reg .b64 ra, rb, rc;\n"
            "mov.b64 ra, {%0, %1};\n"
            "mov.b64 rb, {%2, %2};\n"
            "mov.b64 rc, {%3, %3};\n"
            "fma.rn.f32x2 ra, ra, rb, rc;\n"
            "mov.b64 {%0, %1}, ra;\n"
            "}\n"
            : "+f"(acc[i]), "+f"(acc[i + 1])
            : "f"(a), "f"(b));
      }
    }
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456f) sink[0] = s;
}


// ----------------------------------------------------------------------------
// K5: parameter gradients of the training loss from the streams written by the DUMP evaluators:
//   out[a*ldo + b] = scale * sum_k A[a][k] * B[b][k]     (A == nullptr: A = 1, i.e. bias / scaler sums)
// Tiny GEMMs with a long K = H*G (n_a, n_b <= 33).  Stage 1: grid (K chunks, tasks); a CTA stages its
// [n_a + n_b] x WG_KC slab of the streams in shared memory with coalesced row reads (every stream element is read
// from global memory ONCE), each thread forms whole (a, b) dot products over the slab (rows padded by one float:
// conflict-free).  Stage 2 sums the per-chunk partials in a fixed order -- deterministic, no atomics.
// ----------------------------------------------------------------------------
constexpr int WG_KC = 256;
struct WgradTask {
  const float* A;
  const float* B;
  float* out;
  int n_a, n_b, ldo, pair0;  // pair0: first (a, b) pair of this task in the partial buffer
  // has_e: the LAST of the n_a rows is not A's but the row E (nullptr: ones) -- the bias sums ride along with the
  // weight task (the packed layouts keep b right behind w), so the B rows are read once for both
  const float* E;
  int has_e;
};
struct WgradK {
  WgradTask task[24];
  int n_tasks, n_pairs, n_chunks;
  unsigned long long K;  // columns; row stride of A and B
  float scale;
  float* partial;  // [n_chunks][n_pairs]
};

// Register tiling: a thread owns a 4 x 4 tile of (a, b) pairs over one quarter of the slab's columns (8 LDS feed 16
// FMA; rows of one tile are 4 apart, so the 32 tiles of a warp read distinct banks); the four column quarters are
// summed through shared memory in a fixed order.
constexpr int WG_ROWS_MAX = 72;  // up4(n_a) + up4(n_b), n_a, n_b <= 33
__global__ void __launch_bounds__(256) k_wgrad_partial(const __grid_constant__ WgradK a) {
  extern __shared__ float slab[];  // [up4(n_a) + up4(n_b)][WG_KC + 1]
  const WgradTask& t = a.task[blockIdx.y];
  const unsigned long long k0 = (unsigned long long)blockIdx.x * WG_KC;
  const int kn = (int)min((unsigned long long)WG_KC, a.K - k0);
  constexpr int LD = WG_KC + 1;
  const int na4 = (t.n_a + 3) & ~3, nb4 = (t.n_b + 3) & ~3;
  const int rows = na4 + nb4;
  {  // thread = column of the slab (blockDim.x == WG_KC); 8 independent row loads in flight per thread
    const int c = threadIdx.x;
    const bool in_k = c < kn;
    const size_t col = (size_t)k0 + c;
    const int n_own = t.has_e ? t.n_a - 1 : t.n_a;  // rows that come from A
    for (int r0 = 0; r0 < rows; r0 += 8) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = r0 + q;
        float x = 0.0f;
        if (in_k && r < rows) {
          if (r < na4) {
            if (r < n_own) x = t.A ? t.A[(size_t)r * a.K + col] : 1.0f;
            else if (r < t.n_a) x = t.E ? t.E[col] : 1.0f;
          } else if (r - na4 < t.n_b) {
            x = t.B[(size_t)(r - na4) * a.K + col];
          }
        }
        v[q] = x;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (r0 + q < rows) slab[(r0 + q) * LD + c] = v[q];
    }
  }
  __syncthreads();
  const int tiles_b = nb4 >> 2, n_tiles = (na4 >> 2) * tiles_b;
  const int kq = threadIdx.x >> 6, c0 = kq * (WG_KC / 4);
  float acc[2][16];  // n_tiles <= 81: at most two tiles per thread
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const int tile = (threadIdx.x & 63) + 64 * m;
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[m][q] = 0.0f;
    if (tile < n_tiles) {
      const int ta = tile / tiles_b, tb = tile - ta * tiles_b;
      const float* sa = slab + (4 * ta) * LD + c0;
      const float* sb = slab + (na4 + 4 * tb) * LD + c0;
#pragma unroll 4
      for (int c = 0; c < WG_KC / 4; ++c) {  // the slab is zero-padded beyond kn and beyond n_a / n_b
        float va[4], vb[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          va[q] = sa[q * LD + c];
          vb[q] = sb[q * LD + c];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[m][i * 4 + jj] = fmaf(va[i], vb[jj], acc[m][i * 4 + jj]);
      }
    }
  }
  __syncthreads();  // every thread is done with the slab: reuse it as red[4][pairs]
  const int pairs = t.n_a * t.n_b;
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const int tile = (threadIdx.x & 63) + 64 * m;
    if (tile < n_tiles) {
      const int ta = tile / tiles_b, tb = tile - ta * tiles_b;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int ia = 4 * ta + i, ib = 4 * tb + jj;
          if (ia < t.n_a && ib < t.n_b) slab[kq * pairs + ia * t.n_b + ib] = acc[m][i * 4 + jj];
        }
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < pairs; p += blockDim.x)
    a.partial[(size_t)blockIdx.x * a.n_pairs + t.pair0 + p] = (slab[p] + slab[pairs + p]) + (slab[2 * pairs + p] + slab[3 * pairs + p]);
}

// Stage 2: a CTA sums 32 adjacent pairs over all chunks: 8 chunk lanes per pair (coalesced 128-byte rows of the partial
// buffer), combined in a fixed order.
__global__ void __launch_bounds__(256) k_wgrad_final(const __grid_constant__ WgradK a) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, cl = threadIdx.x >> 5;
  const int q = blockIdx.x * 32 + lane;
  float s = 0.0f;
  if (q < a.n_pairs)
    for (int c = cl; c < a.n_chunks; c += 8) s += a.partial[(size_t)c * a.n_pairs + q];
  red[cl][lane] = s;
  __syncthreads();
  if (cl != 0 || q >= a.n_pairs) return;
#pragma unroll
  for (int w = 1; w < 8; ++w) s += red[w][lane];
  int ti = 0;
  while (ti + 1 < a.n_tasks && q >= a.task[ti + 1].pair0) ++ti;
  const WgradTask& t = a.task[ti];
  const int p = q - t.pair0;
  const int ia = p / t.n_b, ib = p - ia * t.n_b;
  t.out[(size_t)ia * t.ldo + ib] = s * a.scale;
}

}  // namespace sde4

// sde4b200.cu -- C-ABI of libsde4b200.so (see include/sde4b200.h for the contract).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "registry.cuh"
#include "misc_kernels.cuh"
#include "apg_kernels.cuh"

namespace sde4 {

std::vector<ArchEntry>& arch_table() {
  static std::vector<ArchEntry> table;
  return table;
}

std::vector<WsEntry>& ws_table() {
  static std::vector<WsEntry> table;
  return table;
}

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}

static const ArchEntry* find_arch(const sde4_model_desc& d) {
  for (const ArchEntry& e : arch_table())
    if (e.matches(d)) return &e;
  return nullptr;
}
// first registered variant of the architecture; SDE4_WS_NPART (environment, tuning experiments only) picks another
// particles-per-CTA variant when one is compiled
std::vector<TcEntry>& tc_table() {
  static std::vector<TcEntry> table;
  return table;
}
static const TcEntry* find_tc(const sde4_model_desc& d) {
  for (const TcEntry& e : tc_table())
    if (e.matches(d)) return &e;
  return nullptr;
}

static const WsEntry* find_ws(const sde4_model_desc& d) {
  static const int want = [] {
    const char* v = getenv("SDE4_WS_NPART");
    return v ? atoi(v) : 0;
  }();
  const WsEntry* first = nullptr;
  for (const WsEntry& e : ws_table()) {
    if (!e.matches(d)) continue;
    if (want > 0 && e.npart == want) return &e;
    if (!first) first = &e;
  }
  return first;
}

// number of SMs of the current device (queried once per device; 148 on a B200)
static int sm_count() {
  static int cached[64] = {0};
  const int dev = current_device();
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  int n = 0;
  if (dev < 0 || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    (void)cudaGetLastError();
    return 148;
  }
  if (dev < 64) cached[dev] = n;
  return n;
}

static std::string describe(const sde4_model_desc& d) {
  char buf[384];
  snprintf(buf, sizeof(buf),
           "model_id=%d n_x=%d n_u=%d flags=0x%x density=[%d,%d,%d,%d act %d] net_a=[%d,%d,%d,%d act %d] "
           "net_b=[%d,%d,%d,%d act %d] in_mask=0x%x out_mask=0x%x n_scaler=%d a_mask=0x%x b_mask=0x%x",
           d.model_id, d.n_x, d.n_u, d.flags, d.density.n_in, d.density.h1, d.density.h2, d.density.n_out,
           d.density.act, d.net_a.n_in, d.net_a.h1, d.net_a.h2, d.net_a.n_out, d.net_a.act, d.net_b.n_in, d.net_b.h1,
           d.net_b.h2, d.net_b.n_out, d.net_b.act, d.noise_in_mask, d.noise_out_mask, d.n_scaler, d.net_a_in_mask,
           d.net_b_in_mask);
  return buf;
}

static int pick_threads(int requested, long long threads_total) {
  if (requested > 0) return requested;
  // few threads: one warp per CTA so the warps spread over all SMs; many: fatter CTAs
  const long long sms = sm_count();
  if (threads_total <= sms * 32 * 8) return 32;
  if (threads_total <= sms * 64 * 8) return 64;
  return 128;
}

// lanes per sample: with few samples one thread per sample leaves most of the 592 warp schedulers idle and each
// resident warp latency bound, so a lane-split kernel is used when it exists.  Rule fitted to scripts/lane_sweep.py
// (hexacopter, H = 20, G problems): the widest compiled split whose launch stays near one wave of 7 warps per SM
// (G*L <= 40960 threads); otherwise the narrowest compiled split while the launch stays below ~98 k threads
// (value only) / ~229 k threads (value + gradient); beyond that the one-thread-per-sample throughput kernel.
static int pick_lanes(const ArchEntry& e, int requested, long long G, bool grad) {
  if (requested == 1) return 1;
  if (requested > 1) {
    const int i = ilog2(requested);
    return (i < 5 && (1 << i) == requested && e.launch_cost[i]) ? requested : -1;
  }
  for (int i = 4; i >= 1; --i)
    if (e.launch_cost[i] && G * (1LL << i) <= 40960) return 1 << i;
  const long long limit = grad ? 229376 : 98304;
  for (int i = 1; i <= 4; ++i)
    if (e.launch_cost[i] && G * (1LL << i) <= limit) return 1 << i;
  return 1;
}

static int max_lanes(const ArchEntry& e) {
  for (int i = 4; i >= 1; --i)
    if (e.launch_cost[i]) return 1 << i;
  return 1;
}

static sde4_cost_desc empty_cost() {
  sde4_cost_desc c;
  memset(&c, 0, sizeof(c));
  for (int i = 0; i < SDE4_MAX_NX; ++i) c.slack_col[i] = -1;
  return c;
}

struct CostLayout {
  size_t xbuf, dwbuf, auxbuf, gpart, cpart, ftape, total;
  int PW, NP, warp_reduce;
};

// lanes: lanes per sample of the kernel that writes the partials (32 / lanes samples share one partial when they all
// belong to one problem); ws: the warp-specialised kernel also keeps the residual-net inputs of every step
static CostLayout cost_layout(const ArchEntry& e, int P, int N, int H, int n_slack, bool grad, bool have_xtraj,
                              int noise_mode, int lanes, bool ws = false) {
  CostLayout L;
  const size_t G = (size_t)P * N;
  const int spw = 32 / lanes;  // samples per warp
  L.warp_reduce = (N % spw == 0) ? 1 : 0;
  L.NP = L.warp_reduce ? N / spw : N;
  L.PW = P * L.NP;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  size_t off = 0;
  L.xbuf = off;
  if (grad && !have_xtraj) off += al((size_t)(H + 1) * e.nx * G * 4);
  L.dwbuf = off;
  if (grad && noise_mode == SDE4_NOISE_THREEFRY) off += al((size_t)H * e.nx * G * 4);
  L.auxbuf = off;
  if (grad && e.naux) off += al((size_t)H * G * 4);
  L.gpart = off;
  if (grad) off += al((size_t)H * (e.nu + n_slack) * L.PW * 4);
  L.cpart = off;
  off += al((size_t)L.PW * 4);
  L.ftape = off;
  if (ws) off += al((size_t)H * 6 * G * 4);
  L.total = off;
  return L;
}

}  // namespace sde4

using namespace sde4;

extern "C" {

int sde4_version(void) { return SDE4_ABI_VERSION; }
void sde4_struct_sizes(int32_t out[4]) {
  out[0] = (int32_t)sizeof(sde4_model_desc);
  out[1] = (int32_t)sizeof(sde4_cost_desc);
  out[2] = (int32_t)sizeof(sde4_rollout_args);
  out[3] = (int32_t)sizeof(sde4_cost_args);
}
void sde4_apg_struct_sizes(int32_t out[2]) {
  out[0] = (int32_t)sizeof(sde4_apg_params);
  out[1] = (int32_t)sizeof(sde4_apg_state);
}
void sde4_aux_struct_sizes(int32_t out[3]) {
  out[0] = (int32_t)sizeof(sde4_sim_out);
  out[1] = (int32_t)sizeof(sde4_xgpu_reduce);
  out[2] = (int32_t)sizeof(sde4_density_reg_args);
}
const char* sde4_last_error(void) { return g_err; }

const char* sde4_supported_archs(void) {
  static std::string s;
  s.clear();
  for (const ArchEntry& e : arch_table()) {
    s += e.name;
    s += " ";
  }
  return s.c_str();
}

int sde4_model_supported(const sde4_model_desc* model) { return model && find_arch(*model) ? 1 : 0; }

int sde4_param_layout(const sde4_model_desc* model, int32_t* offsets) {
  if (!model) return fail(SDE4_ERR_INVALID, "null model");
  const ArchEntry* e = find_arch(*model);
  if (!e) return fail(SDE4_ERR_UNSUPPORTED, "no compiled specialisation for %s; compiled: %s", describe(*model).c_str(), sde4_supported_archs());
  if (offsets) memcpy(offsets, e->offsets, sizeof(e->offsets));
  return e->param_total;
}

static int rollout_impl(const sde4_model_desc* model, const sde4_rollout_args* a, const sde4_sim_out* so, void* stream);

int sde4_rollout(const sde4_model_desc* model, const sde4_rollout_args* a, void* stream) {
  return rollout_impl(model, a, nullptr, stream);
}

int sde4_sim_rollout(const sde4_model_desc* model, const sde4_rollout_args* a, const sde4_sim_out* so, void* stream) {
  if (!so) return fail(SDE4_ERR_INVALID, "null sde4_sim_out");
  if (!so->last) return fail(SDE4_ERR_INVALID, "sim: `last` is required");
  if (so->reward_kind != SDE4_REWARD_NONE && so->reward_kind != SDE4_REWARD_CARTPOLE_SWINGUP)
    return fail(SDE4_ERR_UNSUPPORTED, "sim: unknown reward kind");
  if (so->reward_kind != SDE4_REWARD_NONE && !so->racc) return fail(SDE4_ERR_INVALID, "sim: `racc` is required with a reward");
  if (so->reward_kind == SDE4_REWARD_CARTPOLE_SWINGUP && (!model || model->n_x != 4))
    return fail(SDE4_ERR_UNSUPPORTED, "sim: the cartpole swing-up reward needs the 4-state cartpole model");
  if (a && a->solver != SDE4_SOLVER_EULER_MARUYAMA)
    return fail(SDE4_ERR_UNSUPPORTED, "sim: the simulator summary integrates with Euler-Maruyama only");
  return rollout_impl(model, a, so, stream);
}

static int rollout_impl(const sde4_model_desc* model, const sde4_rollout_args* a, const sde4_sim_out* so, void* stream) {
  if (!model || !a) return fail(SDE4_ERR_INVALID, "null argument");
  if (a->struct_size != (int)sizeof(sde4_rollout_args)) return fail(SDE4_ERR_INVALID, "sde4_rollout_args size mismatch (ABI)");
  const ArchEntry* e = find_arch(*model);
  if (!e) return fail(SDE4_ERR_UNSUPPORTED, "no compiled specialisation for %s; compiled: %s", describe(*model).c_str(), sde4_supported_archs());
  if (a->P <= 0 || a->N <= 0 || a->H <= 0) return fail(SDE4_ERR_INVALID, "P, N, H must be positive");
  if (!a->params || !a->y0 || !a->u || !a->dt || (!so && !a->xevol)) return fail(SDE4_ERR_INVALID, "null buffer");
  if (((uintptr_t)a->params & 15) != 0) return fail(SDE4_ERR_INVALID, "params must be 16-byte aligned");
  if (a->x_rows < e->nx) return fail(SDE4_ERR_INVALID, "x_rows < n_x");
  if (a->noise_mode == SDE4_NOISE_THREEFRY && !a->key) return fail(SDE4_ERR_INVALID, "threefry noise needs a key");
  if (a->noise_mode == SDE4_NOISE_GIVEN && !a->noise) return fail(SDE4_ERR_INVALID, "given noise needs a buffer");
  if (a->noise_mode < 0 || a->noise_mode > 2 || a->rng_mode < 0 || a->rng_mode > 1) return fail(SDE4_ERR_INVALID, "bad noise/rng mode");
  const long long G = (long long)a->P * a->N;
  if (G > 0x7fffffffLL) return fail(SDE4_ERR_INVALID, "P*N exceeds 2^31-1");
  RolloutK k;
  k.params = a->params;
  k.y0 = a->y0;
  k.u = a->u;
  k.u_sp = a->u_stride_p;
  k.u_st = a->u_stride_t;
  k.u_sj = a->u_stride_j;
  k.dt = a->dt;
  k.key = a->key;
  k.key_sp = a->key_stride_p;
  k.noise = a->noise;
  k.xevol = a->xevol;
  k.P = a->P;
  k.N = a->N;
  k.H = a->H;
  k.G = (int)G;
  k.x_rows = a->x_rows;
  k.rng_mode = a->rng_mode;
  k.noise_mode = a->noise_mode;
  k.n_off = a->particle_total > 0 ? a->particle_offset : 0;
  k.n_tot = a->particle_total > 0 ? a->particle_total : a->N;
  if (k.n_off < 0 || k.n_off + a->N > k.n_tot) return fail(SDE4_ERR_INVALID, "bad particle shard");
  if (a->solver < SDE4_SOLVER_EULER_MARUYAMA || a->solver > SDE4_SOLVER_STRAT_MILSTEIN) return fail(SDE4_ERR_INVALID, "unknown solver");
  k.solver = a->solver;
  k.last = so ? so->last : nullptr;
  k.racc = so ? so->racc : nullptr;
  k.reward_kind = so ? so->reward_kind : SDE4_REWARD_NONE;
  // tensor-core forward rollout (kernels_tc.cuh): on request, or by itself in the throughput regime (at least two
  // 128-sample tiles per SM; profiles/r2_tcgen05_experiment.md), Euler-Maruyama only
  const TcEntry* tce = nullptr;
  if (a->lanes == SDE4_LANES_TC) {
    tce = a->solver == SDE4_SOLVER_EULER_MARUYAMA ? find_tc(*model) : nullptr;
    if (!tce) return fail(SDE4_ERR_UNSUPPORTED, "no tensor-core rollout kernel for this call (%s; needs [32, 32] networks, Euler-Maruyama)", e->name);
  } else if (a->lanes == 0 && a->solver == SDE4_SOLVER_EULER_MARUYAMA && G >= 2ll * 128 * sm_count()) {
    tce = find_tc(*model);
  }
  cudaError_t err;
  if (tce) {
    err = tce->launch_rollout(*model, k, current_device(), (cudaStream_t)stream);
  } else {
    const int lanes = pick_lanes(*e, a->lanes, G, false);
    if (lanes < 0) return fail(SDE4_ERR_UNSUPPORTED, "requested lane count has no compiled kernel for %s", e->name);
    const long long nthr = G * lanes;
    const int threads = pick_threads(a->block_threads, nthr);
    const int blocks = (int)((nthr + threads - 1) / threads);
    err = e->launch_rollout[ilog2(lanes)](*model, k, blocks, threads, (cudaStream_t)stream);
  }
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "rollout launch (%s): %s", e->name, cudaGetErrorString(err));
  if (so && (so->mean || so->std || so->reward_mean || so->reward_of_mean || so->done_of_mean)) {
    SimReduceK r;
    r.last = so->last;
    r.racc = so->racc;
    r.u_last = a->u + (size_t)(a->H - 1) * a->u_stride_t;
    r.u_sp = a->u_stride_p;
    r.u_sj = a->u_stride_j;
    r.mean = so->mean;
    r.std = so->std;
    r.reward_mean = so->reward_mean;
    r.reward_of_mean = so->reward_of_mean;
    r.done_of_mean = so->done_of_mean;
    r.P = a->P;
    r.N = a->N;
    r.n_x = e->nx;
    r.reward_kind = so->reward_kind;
    const long long items = (long long)a->P * e->nx;
    k_sim_reduce<<<(unsigned)((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(r);
    err = cudaGetLastError();
    if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "sim reduce launch: %s", cudaGetErrorString(err));
  }
  return SDE4_OK;
}

size_t sde4_cost_workspace_bytes(const sde4_model_desc* model, int32_t P, int32_t N, int32_t H, int32_t n_slack,
                                 int32_t want_grad, int32_t have_xtraj) {
  if (!model) return 0;
  const ArchEntry* e = find_arch(*model);
  if (!e) return 0;
  // worst case over noise modes (threefry keeps a dw buffer)
  size_t a = cost_layout(*e, P, N, H, n_slack, want_grad != 0, have_xtraj != 0, SDE4_NOISE_THREEFRY, 1).total;
  for (int l = 2; l <= max_lanes(*e); l *= 2) {
    size_t b = cost_layout(*e, P, N, H, n_slack, want_grad != 0, have_xtraj != 0, SDE4_NOISE_THREEFRY, l).total;
    if (b > a) a = b;
  }
  if (const WsEntry* ws = find_ws(*model)) {
    if (want_grad) {
      size_t b = cost_layout(*e, P, N, H, n_slack, true, have_xtraj != 0, SDE4_NOISE_THREEFRY, 32 / ws->npart, true).total;
      if (b > a) a = b;
    }
  }
  return a;
}

struct LossOpts {
  bool enabled = false;
  float* grad_params = nullptr;
};

static int cost_grad_impl(const sde4_model_desc* model, const sde4_cost_args* a, void* stream, const LossOpts& lo);

int sde4_cost_grad(const sde4_model_desc* model, const sde4_cost_args* a, void* stream) {
  return cost_grad_impl(model, a, stream, LossOpts());
}

// ---- host-buffer variant -----------------------------------------------------------------------
// staging layout (floats): [ y0 P*n_x | opt P*H*n_opt | xref Px*(H+1)*n_x | key 2*Pk ] in, [ cost P | grad P*H*n_opt ] out
namespace {
struct HostStage {
  size_t y0, opt, xref, key, in_floats, cost, grad, total;
};
HostStage host_stage(int P, int H, int nx, int n_opt, int xref_rows, int key_rows) {
  HostStage h;
  h.y0 = 0;
  h.opt = h.y0 + (size_t)P * nx;
  h.xref = h.opt + (size_t)P * H * n_opt;
  h.key = h.xref + (size_t)xref_rows * (H + 1) * nx;
  h.in_floats = (h.key + 2 * (size_t)key_rows + 3) & ~(size_t)3;
  h.cost = h.in_floats;
  h.grad = h.cost + (((size_t)P + 3) & ~(size_t)3);
  h.total = h.grad + (size_t)P * H * n_opt;
  return h;
}
}  // namespace

size_t sde4_cost_grad_host_staging_bytes(const sde4_model_desc* model, int32_t P, int32_t H, int32_t n_slack) {
  if (!model) return 0;
  const ArchEntry* e = find_arch(*model);
  if (!e) return 0;
  return host_stage(P, H, e->nx, e->nu + n_slack, P, P).total * sizeof(float);
}

int sde4_cost_grad_host(const sde4_model_desc* model, const sde4_cost_args* a, void* staging_pinned,
                        void* staging_device, size_t staging_bytes, void* stream) {
  if (!model || !a) return fail(SDE4_ERR_INVALID, "null argument");
  if (a->struct_size != (int)sizeof(sde4_cost_args)) return fail(SDE4_ERR_INVALID, "sde4_cost_args size mismatch (ABI)");
  const ArchEntry* e = find_arch(*model);
  if (!e) return fail(SDE4_ERR_UNSUPPORTED, "no compiled specialisation for %s; compiled: %s", describe(*model).c_str(), sde4_supported_archs());
  if (!a->y0 || !a->opt || !a->xref || !a->stage || !a->cost || !staging_pinned || !staging_device)
    return fail(SDE4_ERR_INVALID, "null buffer");
  if (a->P <= 0 || a->N <= 0 || a->H <= 0) return fail(SDE4_ERR_INVALID, "P, N, H must be positive");
  if (a->data_mod != 0 && a->data_mod != a->P) return fail(SDE4_ERR_UNSUPPORTED, "host variant: data_mod must be 0 or P");
  const int P = a->P, H = a->H, nx = e->nx, n_opt = e->nu + a->stage->n_slack;
  if (a->opt_stride_j != 1 || a->opt_stride_t != n_opt || (P > 1 && a->opt_stride_p != (long long)H * n_opt))
    return fail(SDE4_ERR_INVALID, "host variant: opt / grad must be contiguous [P][H][n_opt]");
  const int xrows = a->xref_stride_p == 0 ? 1 : P;
  if (a->xref_stride_p != 0 && a->xref_stride_p != (long long)(H + 1) * nx)
    return fail(SDE4_ERR_INVALID, "host variant: xref must be contiguous [P][H+1][n_x] or shared");
  const bool keyed = a->noise_mode == SDE4_NOISE_THREEFRY;
  if (keyed && !a->key) return fail(SDE4_ERR_INVALID, "threefry noise needs a key");
  const int krows = !keyed ? 0 : (a->key_stride_p == 0 ? 1 : P);
  if (keyed && a->key_stride_p != 0 && a->key_stride_p != 2) return fail(SDE4_ERR_INVALID, "host variant: keys must be contiguous [P][2]");
  const HostStage h = host_stage(P, H, nx, n_opt, xrows, krows);
  if (staging_bytes < h.total * sizeof(float)) return fail(SDE4_ERR_WORKSPACE, "staging buffers too small (see sde4_cost_grad_host_staging_bytes)");
  float* hp = static_cast<float*>(staging_pinned);
  float* dp = static_cast<float*>(staging_device);
  std::memcpy(hp + h.y0, a->y0, (size_t)P * nx * 4);
  std::memcpy(hp + h.opt, a->opt, (size_t)P * H * n_opt * 4);
  std::memcpy(hp + h.xref, a->xref, (size_t)xrows * (H + 1) * nx * 4);
  if (keyed) std::memcpy(hp + h.key, a->key, (size_t)krows * 8);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t err = cudaMemcpyAsync(dp, hp, h.in_floats * 4, cudaMemcpyHostToDevice, s);
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "H2D: %s", cudaGetErrorString(err));
  sde4_cost_args d = *a;
  d.y0 = dp + h.y0;
  d.opt = dp + h.opt;
  d.xref = dp + h.xref;
  d.key = keyed ? reinterpret_cast<const uint32_t*>(dp + h.key) : nullptr;
  // [cost | grad] (P*(1 + H*n_opt) floats, written once by K3): when the pinned staging buffer is mapped into the
  // device's address space (always the case under unified addressing) K3 stores them straight into it -- posted
  // writes over the host link, visible after the stream synchronise -- and the separate D2H copy disappears
  float* hp_dev = nullptr;
  if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&hp_dev), hp, 0) != cudaSuccess) {
    (void)cudaGetLastError();
    hp_dev = nullptr;
  }
  float* outp = hp_dev ? hp_dev : dp;
  d.cost = outp + h.cost;
  d.grad = a->grad ? outp + h.grad : nullptr;
  // particles of the problem sharded over GPUs (sde4_xgpu_reduce): K3's last CTA sums the world's partials; its output
  // -- the contiguous [cost | grad] vector of x.n floats -- goes straight into the staging buffer as well
  sde4_xgpu_reduce xg;
  size_t goff = h.grad;
  if (a->xgpu) {
    xg = *a->xgpu;
    xg.out = outp + h.cost;
    d.xgpu = &xg;
    goff = h.cost + (size_t)P;
  }
  if (int rc = cost_grad_impl(model, &d, stream, LossOpts())) return rc;
  const size_t out_floats = a->grad ? h.total - h.cost : (size_t)P;
  err = cudaSuccess;
  if (!hp_dev) err = cudaMemcpyAsync(hp + h.cost, dp + h.cost, out_floats * 4, cudaMemcpyDeviceToHost, s);
  if (err == cudaSuccess) err = cudaStreamSynchronize(s);
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "D2H: %s", cudaGetErrorString(err));
  std::memcpy(a->cost, hp + h.cost, (size_t)P * 4);
  if (a->grad) std::memcpy(a->grad, hp + goff, (size_t)P * H * n_opt * 4);
  return SDE4_OK;
}

// rows of the parameter-gradient streams: per network [in | h1 | h2 | c1 | g2 | gout] (+ scaler rows after the density
// net), then the physical-scalar rows (rotor); off[3] = offset of the latter
static size_t loss_dump_floats(const ArchEntry& e, size_t K, size_t off[4]) {
  size_t tot = 0;
  for (int s = 0; s < 3; ++s) {
    off[s] = tot;
    const int* d = e.net_dims[s];
    if (d[0] > 0) tot += (size_t)(d[0] + 2 * d[1] + 2 * d[2] + d[3] + (s == 0 ? e.n_scaler : 0)) * K;
  }
  off[3] = tot;
  tot += (size_t)e.n_phys * K;
  return tot;
}

// number of (a, b) output pairs of the K5 tasks (weights, biases, scaler) of an architecture
static int loss_wgrad_pairs(const ArchEntry& e) {
  int pairs = 0;
  for (int s = 0; s < 3; ++s) {
    const int* d = e.net_dims[s];
    if (d[0] == 0) continue;
    pairs += d[0] * d[1] + d[1] + d[1] * d[2] + d[2] + d[3] * d[2] + d[3];
    if (s == 0) pairs += e.n_scaler;
  }
  return pairs + e.n_phys;
}

size_t sde4_loss_workspace_bytes(const sde4_model_desc* model, int32_t B, int32_t N, int32_t H) {
  if (!model) return 0;
  const ArchEntry* e = find_arch(*model);
  if (!e || !e->launch_loss) return 0;
  size_t off[4];
  const size_t K = (size_t)B * N * H;
  const size_t dump = loss_dump_floats(*e, K, off) * 4;
  const size_t partial = ((K + WG_KC - 1) / WG_KC) * (size_t)loss_wgrad_pairs(*e) * 4;
  size_t lay = cost_layout(*e, B, N, H, 0, true, false, SDE4_NOISE_THREEFRY, 1).total;
  if (e->loss_lanes > 1) {
    const size_t l2 = cost_layout(*e, B, N, H, 0, true, false, SDE4_NOISE_THREEFRY, e->loss_lanes).total;
    if (l2 > lay) lay = l2;
  }
  return lay + ((dump + 255) & ~(size_t)255) + partial;
}

int sde4_loss_grad(const sde4_model_desc* model, const sde4_cost_args* a, float* grad_params, void* stream) {
  if (!grad_params) return fail(SDE4_ERR_INVALID, "null grad_params");
  LossOpts lo;
  lo.enabled = true;
  lo.grad_params = grad_params;
  return cost_grad_impl(model, a, stream, lo);
}

static int cost_grad_impl(const sde4_model_desc* model, const sde4_cost_args* a, void* stream, const LossOpts& lo) {
  if (!model || !a) return fail(SDE4_ERR_INVALID, "null argument");
  if (a->struct_size != (int)sizeof(sde4_cost_args)) return fail(SDE4_ERR_INVALID, "sde4_cost_args size mismatch (ABI)");
  const ArchEntry* e = find_arch(*model);
  if (!e) return fail(SDE4_ERR_UNSUPPORTED, "no compiled specialisation for %s; compiled: %s", describe(*model).c_str(), sde4_supported_archs());
  if (a->P <= 0 || a->N <= 0 || a->H <= 0) return fail(SDE4_ERR_INVALID, "P, N, H must be positive");
  if (!a->params || !a->y0 || !a->opt || !a->dt || !a->xref || !a->stage || !a->cost || !a->workspace)
    return fail(SDE4_ERR_INVALID, "null buffer");
  if (((uintptr_t)a->params & 15) != 0) return fail(SDE4_ERR_INVALID, "params must be 16-byte aligned");
  if (a->noise_mode == SDE4_NOISE_THREEFRY && !a->key) return fail(SDE4_ERR_INVALID, "threefry noise needs a key");
  if (a->noise_mode == SDE4_NOISE_GIVEN && !a->noise) return fail(SDE4_ERR_INVALID, "given noise needs a buffer");
  if (a->noise_mode < 0 || a->noise_mode > 2 || a->rng_mode < 0 || a->rng_mode > 1) return fail(SDE4_ERR_INVALID, "bad noise/rng mode");
  const sde4_cost_desc& cd = *a->stage;
  if (cd.kind == SDE4_COST_ROTOR_TRACK && e->nx != 13) return fail(SDE4_ERR_UNSUPPORTED, "rotor tracking cost needs a 13-state model");
  if (cd.n_slack < 0 || e->nu + cd.n_slack > SDE4_MAX_NOPT) return fail(SDE4_ERR_INVALID, "bad n_slack");
  if (cd.constr_mode != SDE4_CONSTR_WITHPROX && cd.n_slack != 0) return fail(SDE4_ERR_INVALID, "n_slack without WITHPROX constraints");
  const long long G = (long long)a->P * a->N;
  if (G > 0x7fffffffLL) return fail(SDE4_ERR_INVALID, "P*N exceeds 2^31-1");
  const bool grad = a->grad != nullptr || lo.enabled;
  const bool constr = cd.constr_mode != SDE4_CONSTR_NONE;
  if (lo.enabled) {
    if (!e->launch_loss) return fail(SDE4_ERR_UNSUPPORTED, "no training-loss kernel compiled for %s", e->name);
    if (constr || cd.kind != SDE4_COST_QUADRATIC || a->xtraj) return fail(SDE4_ERR_INVALID, "training loss: QUADRATIC data cost, no constraints, no xtraj");
  }
  // warp-specialised value + gradient kernel (kernels_ws.cuh): on request, or by itself while the launch is bound by
  // the serial instruction chain of a step (the regime where the widest lane split would be chosen)
  const WsEntry* wse = nullptr;
  {
    const bool ws_ok = !lo.enabled && a->grad != nullptr && !constr && a->noise_mode == SDE4_NOISE_THREEFRY;
    if (a->lanes == SDE4_LANES_WS) {
      wse = ws_ok ? find_ws(*model) : nullptr;
      if (!wse)
        return fail(SDE4_ERR_UNSUPPORTED, "no warp-specialised kernel for this call (%s; needs value + gradient, "
                    "threefry noise, no state constraints)", e->name);
    } else if (a->lanes == 0 && ws_ok && G >= 2048 && G <= 98304) {
      // scripts/ws_sizes.py (hexacopter, H = 20): level with the 8-lane kernel up to ~1 k samples, 13-25 % ahead of the
      // best lane-split kernel from 3.5 k to 65 k samples, behind the one-thread-per-sample kernel at 131 k
      wse = find_ws(*model);
    }
  }
  // training loss: the lane-split kernel while the minibatch is latency bound (G*L <= 64 k threads)
  const int lanes = wse ? 32 / wse->npart
                       : lo.enabled ? ((e->launch_loss_ls && a->lanes != 1 && G * e->loss_lanes <= 65536) ? e->loss_lanes : 1)
                                    : pick_lanes(*e, a->lanes, G, a->grad != nullptr);
  if (lanes < 0) return fail(SDE4_ERR_UNSUPPORTED, "requested lane count has no compiled kernel for %s", e->name);
  const CostLayout L = cost_layout(*e, a->P, a->N, a->H, cd.n_slack, grad, a->xtraj != nullptr, a->noise_mode, lanes, wse != nullptr);
  size_t dump_off[4] = {0, 0, 0, 0};
  const size_t Kcols = (size_t)G * a->H;
  const size_t dump_bytes = lo.enabled ? loss_dump_floats(*e, Kcols, dump_off) * 4 : 0;
  const size_t wg_bytes = lo.enabled ? ((Kcols + WG_KC - 1) / WG_KC) * (size_t)loss_wgrad_pairs(*e) * 4 : 0;
  if (a->workspace_bytes < L.total + ((dump_bytes + 255) & ~(size_t)255) + wg_bytes)
    return fail(SDE4_ERR_WORKSPACE, "workspace too small");
  if (((uintptr_t)a->workspace & 255) != 0) return fail(SDE4_ERR_INVALID, "workspace must be 256-byte aligned");
  char* ws = (char*)a->workspace;
  CostK k;
  k.params = a->params;
  k.y0 = a->y0;
  k.opt = a->opt;
  k.o_sp = a->opt_stride_p;
  k.o_st = a->opt_stride_t;
  k.o_sj = a->opt_stride_j;
  k.dt = a->dt;
  k.xref = a->xref;
  k.xref_sp = a->xref_stride_p;
  k.key = a->key;
  k.key_sp = a->key_stride_p;
  k.noise = a->noise;
  if (a->xtraj) {
    k.xbuf = a->xtraj;
    k.x_rows = e->nx + 1;
    k.write_stage_cost = 1;
  } else if (grad) {
    k.xbuf = (float*)(ws + L.xbuf);
    k.x_rows = e->nx;
    k.write_stage_cost = 0;
  } else {
    k.xbuf = nullptr;
    k.x_rows = e->nx;
    k.write_stage_cost = 0;
  }
  k.dwbuf = (grad && a->noise_mode == SDE4_NOISE_THREEFRY) ? (float*)(ws + L.dwbuf) : nullptr;
  k.auxbuf = (grad && e->naux) ? (float*)(ws + L.auxbuf) : nullptr;
  k.gpart = grad ? (float*)(ws + L.gpart) : nullptr;
  k.cpart = (float*)(ws + L.cpart);
  k.P = a->P;
  k.N = a->N;
  k.H = a->H;
  k.G = (int)G;
  k.rng_mode = a->rng_mode;
  k.noise_mode = a->noise_mode;
  k.n_slack = cd.n_slack;
  k.warp_reduce = L.warp_reduce;
  k.PW = L.PW;
  k.n_off = a->particle_total > 0 ? a->particle_offset : 0;
  k.n_tot = a->particle_total > 0 ? a->particle_total : a->N;
  if (k.n_off < 0 || k.n_off + a->N > k.n_tot) return fail(SDE4_ERR_INVALID, "bad particle shard");
  k.data_mod = a->data_mod > 0 ? a->data_mod : a->P;
  k.mask = a->problem_mask;
  k.step_w = lo.enabled ? a->step_weight : nullptr;
  k.sum_mode = lo.enabled ? 1 : 0;
  k.rng_chain = lo.enabled ? 1 : 0;
  k.pd_stride = Kcols;
  float* dump_base = (float*)(ws + L.total);
  for (int s2 = 0; s2 < 3; ++s2) k.pd_net[s2] = lo.enabled ? dump_base + dump_off[s2] : nullptr;
  k.pd_phys = (lo.enabled && e->n_phys) ? dump_base + dump_off[3] : nullptr;
  k.has_terminal = (a->terminal && a->terminal->kind != SDE4_COST_NONE) ? 1 : 0;
  const sde4_cost_desc td = k.has_terminal ? *a->terminal : empty_cost();
  const long long nthr = G * lanes;
  int threads = pick_threads(a->block_threads, nthr);
  const int blocks = (int)((nthr + threads - 1) / threads);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t err;
  if (lo.enabled) {
    err = cudaMemsetAsync(dump_base, 0, dump_bytes, s);  // rows of networks that are not evaluated (ODE mode) stay 0
    if (err == cudaSuccess) err = (lanes > 1 ? e->launch_loss_ls : e->launch_loss)(*model, cd, td, k, blocks, threads, s);
  } else if (wse) {
    err = wse->launch(*model, cd, td, k, (float*)(ws + L.ftape), current_device(), s);
  } else {
    err = e->launch_cost[ilog2(lanes)](*model, cd, td, k, grad, constr, blocks, threads, s);
  }
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "cost launch (%s): %s", e->name, cudaGetErrorString(err));
  ReduceK r;
  r.gpart = k.gpart;
  r.cpart = k.cpart;
  r.opt = a->opt;
  r.o_sp = a->opt_stride_p;
  r.o_st = a->opt_stride_t;
  r.o_sj = a->opt_stride_j;
  r.dt = a->dt;
  r.cost = a->cost;
  r.grad = a->grad;
  r.P = a->P;
  r.N = k.n_tot;
  r.H = a->H;
  r.n_u = e->nu;
  r.n_opt = e->nu + cd.n_slack;
  r.PW = L.PW;
  r.NP = L.NP;
  r.has_slew = (cd.has_slew && k.n_off == 0) ? 1 : 0;
  r.mask = a->problem_mask;
  r.x.world = 0;
  if (a->xgpu && !lo.enabled) {  // cross-GPU sum folded into K3 (sde4_xgpu_reduce)
    const sde4_xgpu_reduce& x = *a->xgpu;
    const long long n_opt_ll = e->nu + cd.n_slack;
    if (x.world < 1 || x.world > 8 || x.rank < 0 || x.rank >= x.world) return fail(SDE4_ERR_INVALID, "xgpu: bad world / rank");
    if (!a->grad && x.world > 1) return fail(SDE4_ERR_INVALID, "xgpu: needs value + gradient (pass any non-null grad)");
    if (x.n != a->P * (1 + a->H * (int)n_opt_ll)) return fail(SDE4_ERR_INVALID, "xgpu: n must be P * (1 + H * n_opt)");
    if (a->opt_stride_j != 1 || a->opt_stride_t != n_opt_ll || (a->P > 1 && a->opt_stride_p != (long long)a->H * n_opt_ll))
      return fail(SDE4_ERR_INVALID, "xgpu: opt must be contiguous [P][H][n_opt] (the partial buffer is written with opt's strides)");
    if (a->problem_mask) return fail(SDE4_ERR_UNSUPPORTED, "xgpu: problem_mask is not supported");
    if (!x.out || !x.counter) return fail(SDE4_ERR_INVALID, "xgpu: null out / counter");
    for (int q = 0; q < x.world; ++q)
      if (!x.buf[q] || !x.flag[q]) return fail(SDE4_ERR_INVALID, "xgpu: null peer buffer / flag pointer");
    r.x.world = x.world;
    r.x.rank = x.rank;
    r.x.n = x.n;
    r.x.epoch = x.epoch;
    for (int q = 0; q < 8; ++q) {
      r.x.buf[q] = q < x.world ? x.buf[q] : nullptr;
      r.x.flag[q] = q < x.world ? x.flag[q] : nullptr;
    }
    r.x.out = x.out;
    r.x.counter = x.counter;
    r.cost = x.buf[x.rank];
    r.grad = x.buf[x.rank] + a->P;
    r.o_sp = (long long)a->H * n_opt_ll;  // (P == 1 callers may pass any problem stride)
  }
  r.res_mult = cd.res_mult;
  for (int j = 0; j < SDE4_MAX_NU; ++j) r.slew[j] = cd.u_slew_coeff[j];
  // value-only launches reduce the cost row only
  const long long items = (long long)(r.grad ? a->H * r.n_opt + 1 : 1) * a->P;
  r.lpi = (r.NP >= 16 && items * 32 <= (1ll << 24)) ? 32 : 1;  // few problems, many partials: a warp per item
  const int rthreads = 128;
  {
    // behind the warp-specialised cost kernel (which releases its dependents at its start) K3 is launched as a
    // programmatic dependent: its CTAs are in place before K2 retires (not while the stream is being captured)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    const bool pdl = wse != nullptr && cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((items * r.lpi + rthreads - 1) / rthreads));
    cfg.blockDim = dim3(rthreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    err = cudaLaunchKernelEx(&cfg, k_reduce, r);
    if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "reduce launch: %s", cudaGetErrorString(err));
  }
  err = cudaGetLastError();
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "reduce launch: %s", cudaGetErrorString(err));
  if (lo.enabled) {
    // K5: parameter gradients from the streams; scale = 1/(B*N) (mean over minibatch and particles)
    err = cudaMemsetAsync(lo.grad_params, 0, (size_t)e->param_total * 4, s);
    if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "memset: %s", cudaGetErrorString(err));
    WgradK wk;
    wk.n_tasks = 0;
    wk.K = Kcols;
    wk.scale = 1.0f / (float)((double)a->P * k.n_tot);
    int pair0 = 0, max_rows = 0;
    auto add = [&](const float* A, int n_a, const float* B, int n_b, float* out, int ldo, bool ones_row = false) {
      WgradTask& t = wk.task[wk.n_tasks++];
      t.A = A;
      t.B = B;
      t.out = out;
      t.n_a = n_a;
      t.n_b = n_b;
      t.ldo = ldo;
      t.pair0 = pair0;
      t.E = nullptr;
      t.has_e = ones_row ? 1 : 0;
      pair0 += n_a * n_b;
      if (n_a + n_b > max_rows) max_rows = n_a + n_b;
    };
    auto up4l = [](int n) { return (n + 3) & ~3; };
    for (int s2 = 0; s2 < 3; ++s2) {
      const int* dm = e->net_dims[s2];
      if (dm[0] == 0) continue;
      const int IN = dm[0], H1 = dm[1], H2 = dm[2], OUT = dm[3];
      const float* base = k.pd_net[s2];
      const float* r_in = base;
      const float* r_h1 = base + (size_t)IN * Kcols;
      const float* r_h2 = base + (size_t)(IN + H1) * Kcols;
      const float* r_c1 = base + (size_t)(IN + H1 + H2) * Kcols;
      const float* r_g2 = base + (size_t)(IN + 2 * H1 + H2) * Kcols;
      const float* r_go = base + (size_t)(IN + 2 * H1 + 2 * H2) * Kcols;
      float* gp = lo.grad_params + e->net_off[s2];
      const int h1p = up4l(H1), h2p = up4l(H2);
      float* w0 = gp;
      float* b0 = w0 + IN * h1p;
      float* w1 = b0 + h1p;
      float* b1 = w1 + H1 * h2p;
      float* w2 = b1 + h2p;  // transposed: [OUT][h2p]
      float* b2 = w2 + OUT * h2p;
      // (b0 sits right behind w0 and b1 behind w1 in the packed layout: the bias sums are the extra all-ones A row)
      (void)b0;
      (void)b1;
      add(r_in, IN + 1, r_c1, H1, w0, h1p, true);
      add(r_h1, H1 + 1, r_g2, H2, w1, h2p, true);
      add(r_go, OUT, r_h2, H2, w2, h2p);
      add(nullptr, 1, r_go, OUT, b2, 4);
      if (s2 == 0 && e->n_scaler) add(nullptr, 1, r_go + (size_t)OUT * Kcols, e->n_scaler, lo.grad_params + e->off_scaler, e->n_scaler);
    }
    if (e->n_phys) add(nullptr, 1, k.pd_phys, e->n_phys, lo.grad_params + e->offsets[0], e->n_phys);
    wk.n_pairs = pair0;
    wk.n_chunks = (int)((Kcols + WG_KC - 1) / WG_KC);
    wk.partial = (float*)(ws + L.total + ((dump_bytes + 255) & ~(size_t)255));
    if (pair0 != loss_wgrad_pairs(*e)) return fail(SDE4_ERR_INVALID, "internal: wgrad pair count");
    const size_t slab = (size_t)(max_rows + 6) * (WG_KC + 1) * sizeof(float);  // rows padded to multiples of 4 per operand
    static DevOnce wg_attr;
    const int wg_dev = current_device();
    if (wg_attr.need(wg_dev)) {
      cudaFuncSetAttribute(k_wgrad_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(WG_ROWS_MAX * (WG_KC + 1) * sizeof(float)));
      wg_attr.mark(wg_dev);
    }
    k_wgrad_partial<<<dim3(wk.n_chunks, wk.n_tasks), 256, slab, s>>>(wk);
    k_wgrad_final<<<(wk.n_pairs + 31) / 32, 256, 0, s>>>(wk);
    err = cudaGetLastError();
    if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "wgrad launch: %s", cudaGetErrorString(err));
  }
  return SDE4_OK;
}

// ---- distance-aware regulariser (reg_kernels.cuh) -----------------------------------------------------------------
struct RegLayout {
  size_t cen, part, scal, dstream, mstream, partial, total;
  unsigned long long K0, KQ, K2, Ktot;
  int d_rows, m_rows, d_pairs, m_pairs;
};
static RegLayout reg_layout(const ArchEntry& e, int B, int N, int Hd, int ns, int mu_type) {
  RegLayout L;
  const int* dm = e.net_dims[0];
  const int D = dm[0], H1 = dm[1], H2 = dm[2];
  L.K0 = (unsigned long long)B * Hd;
  L.KQ = L.K0 * N;
  L.K2 = L.KQ * ns;
  L.Ktot = L.K2 + 2 * L.K0;
  L.d_rows = D + 2 * H1 + 2 * H2 + 1 + 1;  // [in | h1 | h2 | c1 | g2 | gout | one]
  L.m_rows = mu_type == SDE4_MU_DNN ? D + 2 * 8 + 2 * 8 + 1 : (mu_type == SDE4_MU_GLOBAL ? 1 : 0);
  L.d_pairs = D * H1 + H1 + H1 * H2 + H2 + H2 + 1;
  L.m_pairs = mu_type == SDE4_MU_DNN ? D * 8 + 8 + 8 * 8 + 8 + 8 + 1 : (mu_type == SDE4_MU_GLOBAL ? 1 : 0);
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t o = 0;
  L.cen = o;
  o = al(o + (size_t)(2 * D + 3) * L.K0 * 4);
  L.part = o;
  o = al(o + (size_t)(D + 3) * L.KQ * 4);
  L.scal = o;
  o = al(o + 64);
  L.dstream = o;
  o = al(o + (size_t)L.d_rows * L.Ktot * 4);
  L.mstream = o;
  o = al(o + (size_t)L.m_rows * L.K0 * 4);
  L.partial = o;
  const size_t pd = ((L.Ktot + WG_KC - 1) / WG_KC) * (size_t)L.d_pairs * 4;
  const size_t pm = ((L.K0 + WG_KC - 1) / WG_KC) * (size_t)L.m_pairs * 4;
  o = al(o + (pd > pm ? pd : pm));
  L.total = o;
  return L;
}

size_t sde4_density_reg_workspace_bytes(const sde4_model_desc* model, int32_t B, int32_t N, int32_t Hd, int32_t ns) {
  if (!model || B <= 0 || N <= 0 || Hd <= 0 || ns <= 0) return 0;
  const ArchEntry* e = find_arch(*model);
  if (!e || !e->launch_reg) return 0;
  return reg_layout(*e, B, N, Hd, ns, SDE4_MU_DNN).total;
}

static cudaError_t run_wgrad(WgradK& wk, int max_rows, cudaStream_t s) {
  const size_t slab = (size_t)(max_rows + 6) * (WG_KC + 1) * sizeof(float);  // rows padded to multiples of 4 per operand
  static DevOnce wg_attr;
  const int wg_dev = current_device();
  if (wg_attr.need(wg_dev)) {
    cudaFuncSetAttribute(k_wgrad_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(WG_ROWS_MAX * (WG_KC + 1) * sizeof(float)));
    wg_attr.mark(wg_dev);
  }
  k_wgrad_partial<<<dim3(wk.n_chunks, wk.n_tasks), 256, slab, s>>>(wk);
  k_wgrad_final<<<(wk.n_pairs + 31) / 32, 256, 0, s>>>(wk);
  return cudaGetLastError();
}

// K5 tasks of one 3-layer network whose stream rows are [in | h1 | h2 | c1 | g2 | gout (| one)] with row stride K:
// `one` (may be null: all columns count) weights the bias sums
static void add_net_tasks(WgradK& wk, int& pair0, int& max_rows, const float* base, size_t K, int IN, int H1, int H2,
                          int OUT, const float* one, float* gp) {
  auto add = [&](const float* A, int n_a, const float* B, int n_b, float* out, int ldo, bool extra) {
    WgradTask& t = wk.task[wk.n_tasks++];
    t.A = A;
    t.B = B;
    t.out = out;
    t.n_a = n_a;
    t.n_b = n_b;
    t.ldo = ldo;
    t.pair0 = pair0;
    t.E = one;
    t.has_e = extra ? 1 : 0;
    pair0 += n_a * n_b;
    if (n_a + n_b > max_rows) max_rows = n_a + n_b;
  };
  auto up4l = [](int n) { return (n + 3) & ~3; };
  const float* r_in = base;
  const float* r_h1 = base + (size_t)IN * K;
  const float* r_h2 = base + (size_t)(IN + H1) * K;
  const float* r_c1 = base + (size_t)(IN + H1 + H2) * K;
  const float* r_g2 = base + (size_t)(IN + 2 * H1 + H2) * K;
  const float* r_go = base + (size_t)(IN + 2 * H1 + 2 * H2) * K;
  const int h1p = up4l(H1), h2p = up4l(H2);
  float* w0 = gp;                    // w0[IN][h1p] b0[h1p]: the bias row rides behind the weight rows
  float* w1 = w0 + IN * h1p + h1p;   // w1[H1][h2p] b1[h2p]
  float* w2 = w1 + H1 * h2p + h2p;   // transposed: [OUT][h2p]
  float* b2 = w2 + OUT * h2p;
  add(r_in, IN + 1, r_c1, H1, w0, h1p, true);
  add(r_h1, H1 + 1, r_g2, H2, w1, h2p, true);
  add(r_go, OUT, r_h2, H2, w2, h2p, false);
  add(one, 1, r_go, OUT, b2, 4, false);  // (A = the `one` row itself; nullptr: every column counts)
}

int sde4_density_reg(const sde4_model_desc* model, const sde4_density_reg_args* a, void* stream) {
  if (!model || !a) return fail(SDE4_ERR_INVALID, "null argument");
  if (a->struct_size != (uint32_t)sizeof(sde4_density_reg_args)) return fail(SDE4_ERR_INVALID, "sde4_density_reg_args size mismatch (ABI)");
  const ArchEntry* e = find_arch(*model);
  if (!e) return fail(SDE4_ERR_UNSUPPORTED, "no compiled specialisation for %s; compiled: %s", describe(*model).c_str(), sde4_supported_archs());
  if (!e->launch_reg) return fail(SDE4_ERR_UNSUPPORTED, "no regulariser kernels compiled for %s (needs a density network and a training kernel)", e->name);
  if (a->B <= 0 || a->N <= 0 || a->Hd <= 0) return fail(SDE4_ERR_INVALID, "B, N, Hd must be positive");
  if (a->ball_nsamples <= 0 || a->ball_nsamples > 160) return fail(SDE4_ERR_INVALID, "ball_nsamples must be in 1..160");
  if (!a->params || !a->y || !a->u || !a->key || !a->out || !a->grad_params || !a->workspace) return fail(SDE4_ERR_INVALID, "null buffer");
  if (a->rng_mode < 0 || a->rng_mode > 1) return fail(SDE4_ERR_INVALID, "bad rng mode");
  if (a->mu_type < SDE4_MU_CONSTANT || a->mu_type > SDE4_MU_DNN) return fail(SDE4_ERR_INVALID, "bad mu_type");
  if (a->pen_mu_type < 0 || a->pen_mu_type > 2) return fail(SDE4_ERR_INVALID, "bad pen_mu_type");
  const int D = e->net_dims[0][0], H1 = e->net_dims[0][1], H2 = e->net_dims[0][2];
  if (a->mu_type != SDE4_MU_CONSTANT && (!a->mu_params || !a->grad_mu)) return fail(SDE4_ERR_INVALID, "mu parameters / gradient buffer missing");
  if (a->mu_type == SDE4_MU_DNN) {
    const sde4_net_desc& m = a->mu_net;
    if (m.n_in != D || m.h1 != 8 || m.h2 != 8 || m.n_out != 1 || m.act != SDE4_ACT_TANH)
      return fail(SDE4_ERR_UNSUPPORTED, "mu_coeff network: compiled for hidden_layers [8, 8], tanh, on the density input");
  }
  const unsigned long long kq = (unsigned long long)a->B * a->N * a->Hd;
  if (kq * (unsigned long long)a->ball_nsamples > 0x7fffffffull) return fail(SDE4_ERR_INVALID, "B*N*Hd*ball_nsamples exceeds 2^31-1");
  const RegLayout L = reg_layout(*e, a->B, a->N, a->Hd, a->ball_nsamples, a->mu_type);
  if (a->workspace_bytes < L.total) return fail(SDE4_ERR_WORKSPACE, "workspace too small");
  if (((uintptr_t)a->workspace & 255) != 0) return fail(SDE4_ERR_INVALID, "workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)a->workspace;
  RegK k;
  k.params = a->params;
  k.mu_params = a->mu_params;
  k.y = a->y;
  k.u = a->u;
  k.key = a->key;
  k.rng_mode = a->rng_mode;
  k.B = a->B;
  k.N = a->N;
  k.Hd = a->Hd;
  k.ns = a->ball_nsamples;
  k.mu_type = a->mu_type;
  k.mu0 = a->mu_coeff;
  for (int i = 0; i < REG_MAXD; ++i) k.radius[i] = a->ball_radius[i];
  k.c_g = a->w_grad;
  k.c_s = a->w_scvex;
  k.c_mu = a->w_mu;
  k.a_g = a->w_grad / (float)L.K0;
  k.a_s = a->w_scvex / (float)((double)a->B * a->N);
  k.pen_mu_type = a->pen_mu_type;
  k.pen_mu_temp = a->pen_mu_temp;
  k.cen = (float*)(ws + L.cen);
  k.part = (float*)(ws + L.part);
  k.scal = (float*)(ws + L.scal);
  k.out = a->out;
  k.dstream = (float*)(ws + L.dstream);
  k.mstream = (float*)(ws + L.mstream);
  k.Ktot = L.Ktot;
  k.K2 = L.K2;
  cudaError_t err = cudaMemsetAsync(a->grad_params, 0, (size_t)e->param_total * 4, s);
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "memset: %s", cudaGetErrorString(err));
  if (a->mu_type != SDE4_MU_CONSTANT) {
    const size_t n_mu = a->mu_type == SDE4_MU_DNN ? (size_t)(D * 8 + 8 + 8 * 8 + 8 + 8 + 4) : 1;
    err = cudaMemsetAsync(a->grad_mu, 0, n_mu * 4, s);
    if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "memset: %s", cudaGetErrorString(err));
  }
  err = e->launch_reg(*model, k, s);
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "regulariser launch: %s", cudaGetErrorString(err));
  {  // density network: ball columns + primal / tangent columns, biases over the `one` row
    WgradK wk;
    wk.n_tasks = 0;
    wk.K = L.Ktot;
    wk.scale = 1.0f;
    int pair0 = 0, max_rows = 0;
    const float* one = k.dstream + (size_t)(L.d_rows - 1) * L.Ktot;
    add_net_tasks(wk, pair0, max_rows, k.dstream, L.Ktot, D, H1, H2, 1, one, a->grad_params + e->net_off[0]);
    wk.n_pairs = pair0;
    wk.n_chunks = (int)((L.Ktot + WG_KC - 1) / WG_KC);
    wk.partial = (float*)(ws + L.partial);
    if (pair0 != L.d_pairs) return fail(SDE4_ERR_INVALID, "internal: regulariser pair count");
    err = run_wgrad(wk, max_rows, s);
    if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "wgrad launch: %s", cudaGetErrorString(err));
  }
  if (a->mu_type != SDE4_MU_CONSTANT) {  // (stream-ordered after the density pass: the partial buffer is shared)
    WgradK wk;
    wk.n_tasks = 0;
    wk.K = L.K0;
    wk.scale = 1.0f;
    int pair0 = 0, max_rows = 0;
    if (a->mu_type == SDE4_MU_DNN) {
      add_net_tasks(wk, pair0, max_rows, k.mstream, L.K0, D, 8, 8, 1, nullptr, a->grad_mu);
    } else {
      WgradTask& t = wk.task[wk.n_tasks++];
      t.A = nullptr;
      t.B = k.mstream;
      t.out = a->grad_mu;
      t.n_a = 1;
      t.n_b = 1;
      t.ldo = 1;
      t.pair0 = 0;
      t.E = nullptr;
      t.has_e = 0;
      pair0 = 1;
      max_rows = 2;
    }
    wk.n_pairs = pair0;
    wk.n_chunks = (int)((L.K0 + WG_KC - 1) / WG_KC);
    wk.partial = (float*)(ws + L.partial);
    err = run_wgrad(wk, max_rows, s);
    if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "wgrad launch: %s", cudaGetErrorString(err));
  }
  return SDE4_OK;
}

int sde4_rng_bits(const uint32_t* key_host, uint64_t total, uint64_t offset, uint64_t count, int32_t rng_mode,
                  uint32_t* out, void* stream) {
  if (!key_host || !out) return fail(SDE4_ERR_INVALID, "null argument");
  if (count == 0) return SDE4_OK;
  if (rng_mode == SDE4_RNG_ORIGINAL && total > 0xffffffffull) return fail(SDE4_ERR_UNSUPPORTED, "original layout beyond 2^32 elements");
  k_rng_bits<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(key_host[0], key_host[1], total, offset,
                                                                              count, rng_mode, out, nullptr);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "rng launch: %s", cudaGetErrorString(err));
  return SDE4_OK;
}

int sde4_rng_normal(const uint32_t* key_host, uint64_t total, uint64_t offset, uint64_t count, int32_t rng_mode,
                    float* out, void* stream) {
  if (!key_host || !out) return fail(SDE4_ERR_INVALID, "null argument");
  if (count == 0) return SDE4_OK;
  if (rng_mode == SDE4_RNG_ORIGINAL && total > 0xffffffffull) return fail(SDE4_ERR_UNSUPPORTED, "original layout beyond 2^32 elements");
  k_rng_bits<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(key_host[0], key_host[1], total, offset,
                                                                              count, rng_mode, nullptr, out);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "rng launch: %s", cudaGetErrorString(err));
  return SDE4_OK;
}

int sde4_rng_split(const uint32_t* key_dev, int32_t num, int32_t rng_mode, uint32_t* out, void* stream) {
  if (!key_dev || !out || num <= 0) return fail(SDE4_ERR_INVALID, "bad argument");
  k_rng_split<<<(num + 127) / 128, 128, 0, (cudaStream_t)stream>>>(key_dev, num, rng_mode, out);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "rng launch: %s", cudaGetErrorString(err));
  return SDE4_OK;
}

int sde4_rng_split_many(const uint32_t* keys_dev, uint64_t K, int32_t num, int32_t rng_mode, uint32_t* out, void* stream) {
  if (!keys_dev || !out || num <= 0) return fail(SDE4_ERR_INVALID, "bad argument");
  if (K == 0) return SDE4_OK;
  const uint64_t tot = K * (uint64_t)num;
  k_rng_split_many<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(keys_dev, K, num, rng_mode, out);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "rng launch: %s", cudaGetErrorString(err));
  return SDE4_OK;
}

int sde4_rng_normal_many(const uint32_t* keys_dev, uint64_t K, uint64_t n, int32_t rng_mode, float* out, void* stream) {
  if (!keys_dev || !out) return fail(SDE4_ERR_INVALID, "bad argument");
  if (K == 0 || n == 0) return SDE4_OK;
  if (rng_mode == SDE4_RNG_ORIGINAL && n > 0xffffffffull) return fail(SDE4_ERR_UNSUPPORTED, "original layout beyond 2^32 elements");
  const uint64_t tot = K * n;
  k_rng_normal_many<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(keys_dev, K, n, rng_mode, out);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "rng launch: %s", cudaGetErrorString(err));
  return SDE4_OK;
}

int sde4_rng_choice_mask(const uint32_t* keys_dev, uint64_t K, int32_t n, int32_t m, int32_t rng_mode, float* out, void* stream) {
  if (!keys_dev || !out) return fail(SDE4_ERR_INVALID, "bad argument");
  if (n <= 0 || m < 0 || m > n) return fail(SDE4_ERR_INVALID, "choice: need 0 <= m <= n, n > 0");
  if (n > 1625) return fail(SDE4_ERR_UNSUPPORTED, "choice: jax shuffles in several rounds beyond n = 1625");
  if (K == 0) return SDE4_OK;
  const uint64_t tot = K * (uint64_t)n;
  k_rng_choice_mask<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(keys_dev, K, n, m, rng_mode, out);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "rng launch: %s", cudaGetErrorString(err));
  return SDE4_OK;
}

int sde4_rng_randint_many(const uint32_t* keys_dev, uint64_t K, uint64_t n, int32_t minval, int32_t maxval, int32_t rng_mode,
                          int32_t* out, void* stream) {
  if (!keys_dev || !out) return fail(SDE4_ERR_INVALID, "bad argument");
  if (K == 0 || n == 0) return SDE4_OK;
  // jax: span = 1 when maxval <= minval, so minval is always returned
  const uint32_t span = maxval <= minval ? 1u : (uint32_t)((int64_t)maxval - (int64_t)minval);
  const uint64_t tot = K * n;
  k_rng_randint_many<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(keys_dev, K, n, minval, span, rng_mode, out);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "rng launch: %s", cudaGetErrorString(err));
  return SDE4_OK;
}

static int apg_check(const sde4_apg_params* prm, const sde4_apg_state* st) {
  if (!prm) return fail(SDE4_ERR_INVALID, "null apg params");
  if (prm->P <= 0 || prm->H <= 0 || prm->n_opt <= 0 || prm->n_opt > SDE4_MAX_NOPT || prm->maxls < 0)
    return fail(SDE4_ERR_INVALID, "bad apg sizes");
  if (st && (!st->num_steps || !st->not_done || !st->xk || !st->yk || !st->grad_yk || !st->x_opt))
    return fail(SDE4_ERR_INVALID, "null apg state buffer");
  return SDE4_OK;
}
static inline int apg_chunks(const sde4_apg_params* prm) { return (prm->H * prm->n_opt + APG_CH - 1) / APG_CH; }
// VEC: 2-D grid (problems x element chunks); FIN: one thread per problem
#define SDE4_APG_LAUNCH(KERNEL, VEC, ...)                                                     \
  do {                                                                                        \
    const int th = 128;                                                                       \
    const dim3 grid((prm->P + th - 1) / th, (VEC) ? apg_chunks(prm) : 1);                     \
    KERNEL<<<grid, th, 0, (cudaStream_t)stream>>>(__VA_ARGS__);                               \
    cudaError_t err = cudaGetLastError();                                                     \
    if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, #KERNEL ": %s", cudaGetErrorString(err)); \
  } while (0)

int sde4_apg_scratch_floats(const sde4_apg_params* prm) {
  if (int rc = apg_check(prm, nullptr)) return rc;
  return apg_chunks(prm) * prm->P;
}
int sde4_apg_init(const sde4_apg_params* prm, const sde4_apg_state* st, const float* x0, const float* cost0,
                  const float* grad0, const float* momentum_or_null, const float* stepsize_or_null,
                  float init_stepsize, float beta_init, float* scratch, void* stream) {
  if (int rc = apg_check(prm, st)) return rc;
  if (!st || !x0 || !cost0 || !grad0 || !scratch) return fail(SDE4_ERR_INVALID, "null argument");
  SDE4_APG_LAUNCH(k_apg_init, 1, *prm, *st, x0, grad0, scratch);
  SDE4_APG_LAUNCH(k_apg_init_fin, 0, *prm, *st, cost0, momentum_or_null, stepsize_or_null, init_stepsize, beta_init,
                  scratch, apg_chunks(prm));
  return SDE4_OK;
}
int sde4_apg_candidates(const sde4_apg_params* prm, const sde4_apg_state* st, float* cand, float* svals,
                        int32_t* xmask_or_null, void* stream) {
  if (int rc = apg_check(prm, st)) return rc;
  if (!st || !cand || !svals) return fail(SDE4_ERR_INVALID, "null argument");
  SDE4_APG_LAUNCH(k_apg_candidates, 1, *prm, *st, cand, svals, xmask_or_null);
  return SDE4_OK;
}
int sde4_apg_select(const sde4_apg_params* prm, const sde4_apg_state* st, const float* cand, const float* svals,
                    const float* cand_cost, float* xv, float* sel_step, int32_t* sel_ls, int32_t* xmask_or_null,
                    float* xv_cost_or_null, void* stream) {
  if (int rc = apg_check(prm, st)) return rc;
  if (!st || !cand || !svals || !cand_cost || !xv || !sel_step || !sel_ls) return fail(SDE4_ERR_INVALID, "null argument");
  if ((xmask_or_null == nullptr) != (xv_cost_or_null == nullptr)) return fail(SDE4_ERR_INVALID, "xmask and xv_cost go together");
  SDE4_APG_LAUNCH(k_apg_select, 1, *prm, *st, cand, svals, cand_cost, xv, sel_step, sel_ls, xmask_or_null, xv_cost_or_null);
  return SDE4_OK;
}
int sde4_apg_ls_step(const sde4_apg_params* prm, const sde4_apg_state* st, const float* svals, const float* cand_cost,
                     int32_t k, int32_t* active, void* stream) {
  if (int rc = apg_check(prm, st)) return rc;
  if (!st || !svals || !cand_cost || !active || k < 0 || k > prm->maxls) return fail(SDE4_ERR_INVALID, "bad argument");
  SDE4_APG_LAUNCH(k_apg_ls_step, 0, *prm, *st, svals, cand_cost, (int)k, active);
  return SDE4_OK;
}
int sde4_apg_pick(const sde4_apg_params* prm, const float* xv, const float* xv_cost, float* y, float* y_cost,
                  int32_t* pick, const float* xv_grad_or_null, float* y_grad_or_null, void* stream) {
  if (int rc = apg_check(prm, nullptr)) return rc;
  if (!xv || !xv_cost || !y || !y_cost || !pick) return fail(SDE4_ERR_INVALID, "null argument");
  if ((xv_grad_or_null == nullptr) != (y_grad_or_null == nullptr)) return fail(SDE4_ERR_INVALID, "xv_grad and y_grad go together");
  SDE4_APG_LAUNCH(k_apg_pick, 1, *prm, xv, xv_cost, y, y_cost, pick, xv_grad_or_null, y_grad_or_null);
  return SDE4_OK;
}
int sde4_apg_spec_points(const sde4_apg_params* prm, const sde4_apg_state* st, const float* cand, float* spec, void* stream) {
  if (int rc = apg_check(prm, st)) return rc;
  if (!st || !cand || !spec) return fail(SDE4_ERR_INVALID, "null argument");
  SDE4_APG_LAUNCH(k_apg_spec_points, 1, *prm, *st, cand, spec);
  return SDE4_OK;
}
int sde4_apg_spec_gather(const sde4_apg_params* prm, const int32_t* sel_ls, const float* spec_cost, const float* spec_grad,
                         float* xv_cost, float* xv_grad, void* stream) {
  if (int rc = apg_check(prm, nullptr)) return rc;
  if (!sel_ls || !spec_cost || !spec_grad || !xv_cost || !xv_grad) return fail(SDE4_ERR_INVALID, "null argument");
  SDE4_APG_LAUNCH(k_apg_spec_gather, 1, *prm, sel_ls, spec_cost, spec_grad, xv_cost, xv_grad);
  return SDE4_OK;
}
int sde4_apg_update(const sde4_apg_params* prm, const sde4_apg_state* st, const float* xv, const float* y,
                    const float* y_cost, const float* y_grad, const int32_t* pick, const float* sel_step,
                    const int32_t* sel_ls, float* scratch, void* stream) {
  if (int rc = apg_check(prm, st)) return rc;
  if (!st || !xv || !y || !y_cost || !y_grad || !pick || !sel_step || !sel_ls || !scratch)
    return fail(SDE4_ERR_INVALID, "null argument");
  SDE4_APG_LAUNCH(k_apg_update, 1, *prm, *st, xv, y, y_cost, y_grad, scratch);
  SDE4_APG_LAUNCH(k_apg_update_fin, 0, *prm, *st, y_cost, pick, sel_step, sel_ls, scratch, apg_chunks(prm));
  return SDE4_OK;
}

int sde4_fma_peak(int32_t variant, int32_t blocks, int32_t threads, int32_t iters, float* sink, double* flops,
                  void* stream) {
  if (!sink || blocks <= 0 || threads <= 0 || threads > 1024) return fail(SDE4_ERR_INVALID, "bad argument");
  if (variant == 0)
    k_fma_peak<0><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink);
  else
    k_fma_peak<1><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(SDE4_ERR_CUDA, "fma_peak launch: %s", cudaGetErrorString(err));
  if (flops) *flops = 2.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
  return SDE4_OK;
}

}  // extern "C"

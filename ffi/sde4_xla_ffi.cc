// sde4_xla_ffi.cc -- XLA FFI (jax.ffi) custom-call handlers over the C-ABI of libsde4b200.so.
//
// This is the thin layer the north star asks for: a jitted JAX program (the reference's own host code:
// sde4mbrl/nsde.py multi_sampling closures, sde4mbrl/apg.py) reaches the CUDA kernels through three custom calls,
//
//   sde4_rollout    create_sampling_fn(...).multi_sampling                    sde4mbrl/nsde.py:1024-1028
//   sde4_cost_grad  create_online_cost_sampling_fn(...).multi_sampling and
//                   jax.value_and_grad of it                                  sde4mbrl/nsde.py:1576-1586, apg.py:82,224
//   sde4_loss_grad  create_model_loss_fn(...).loss_fn data term + jax.grad    sde4mbrl/nsde.py:1185-1203, train_sde.py:248-262
//
// each of which unpacks XLA buffers / attributes into the POD argument structs of include/sde4b200.h and calls the
// launcher on XLA's stream.  Descriptors (sde4_model_desc / sde4_cost_desc) travel as byte-array attributes; scratch
// space is an extra RESULT buffer that XLA allocates (sized on the Python side with sde4_cost_workspace_bytes), so the
// library still allocates nothing.  The Python half (registration, custom_vjp) is ffi/nsde_b200_jax.py.
//
// The XLA FFI headers ship with jaxlib (jax.ffi.include_dir()); they are not in the build image of this repository, so
// everything below is guarded: without the header this file compiles to a stub that only exports
// sde4_xla_ffi_available() == 0.  tests/test_ffi_shim.py compiles BOTH variants (the real one against a minimal mock of
// the header's binding API, tests/mock_xla/) so the guard and the handler signatures are exercised on every run.
#include <stdint.h>

#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define SDE4_HAVE_XLA_FFI 1
#endif
#endif

#ifndef SDE4_HAVE_XLA_FFI

extern "C" int sde4_xla_ffi_available(void) { return 0; }

#else  // ------------------------------------------------------------------------------------------------------------

#include <cuda_runtime.h>

#include <cstring>
#include <string>

#include "../include/sde4b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

extern "C" int sde4_xla_ffi_available(void) { return 1; }

namespace {

template <class T>
bool load_pod(ffi::Span<const uint8_t> bytes, T* out) {
  if (bytes.size() != sizeof(T)) return false;
  std::memcpy(out, bytes.data(), sizeof(T));
  return true;
}

ffi::Error from_status(int rc, const char* what) {
  if (rc == SDE4_OK) return ffi::Error::Success();
  const std::string msg = std::string(what) + ": " + sde4_last_error();
  if (rc == SDE4_ERR_INVALID || rc == SDE4_ERR_UNSUPPORTED) return ffi::Error::InvalidArgument(msg);
  return ffi::Error::Internal(msg);
}

// strides (in elements) of a row-major [P, H, n] or [H, n] control array; a leading extent of 1 broadcasts
bool control_strides(ffi::Span<const int64_t> dims, int64_t P, int64_t* H, int64_t* n, int64_t* sp, int64_t* st) {
  if (dims.size() == 2) {
    *H = dims[0], *n = dims[1], *sp = 0, *st = dims[1];
    return true;
  }
  if (dims.size() != 3 || (dims[0] != P && dims[0] != 1)) return false;
  *H = dims[1], *n = dims[2], *st = dims[2];
  *sp = dims[0] == 1 ? 0 : dims[1] * dims[2];
  return true;
}

// ---- sde4_rollout ---------------------------------------------------------------------------------------------------
// args   params f32[NP], y0 f32[P, n_x], u f32[P | 1, H, n_u] or [H, n_u], dt f32[H], key u32[P | 1, 2]
// attrs  model_desc (bytes of sde4_model_desc), num_particles, rng_mode, noise_mode, solver
// result xevol f32[H + 1, n_x, P * N]   (sample-minor; the reference's [P, N, H+1, n_x] is a transpose of it)
ffi::Error RolloutImpl(cudaStream_t stream, ffi::Span<const uint8_t> model_desc, int32_t num_particles, int32_t rng_mode,
                       int32_t noise_mode, int32_t solver, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> y0,
                       ffi::Buffer<ffi::F32> u, ffi::Buffer<ffi::F32> dt, ffi::Buffer<ffi::U32> key,
                       ffi::ResultBuffer<ffi::F32> xevol) {
  sde4_model_desc md;
  if (!load_pod(model_desc, &md)) return ffi::Error::InvalidArgument("sde4_rollout: model_desc has the wrong size (ABI)");
  if (y0.dimensions().size() != 2 || y0.dimensions()[1] != md.n_x)
    return ffi::Error::InvalidArgument("sde4_rollout: y0 must be [P, n_x]");
  const int64_t P = y0.dimensions()[0];
  int64_t H, nu, sp, st;
  if (!control_strides(u.dimensions(), P, &H, &nu, &sp, &st) || nu != md.n_u || (int64_t)dt.element_count() != H)
    return ffi::Error::InvalidArgument("sde4_rollout: u must be [P | 1, H, n_u] or [H, n_u] with H = len(dt)");
  const int64_t G = P * num_particles;
  if ((int64_t)xevol->element_count() != (H + 1) * md.n_x * G)
    return ffi::Error::InvalidArgument("sde4_rollout: xevol must be [H + 1, n_x, P * N]");
  const int64_t nkeys = key.element_count() / 2;
  if (noise_mode == SDE4_NOISE_THREEFRY && nkeys != 1 && nkeys != P)
    return ffi::Error::InvalidArgument("sde4_rollout: one key or one key per problem");
  sde4_rollout_args a;
  std::memset(&a, 0, sizeof(a));
  a.struct_size = (int32_t)sizeof(a);
  a.P = (int32_t)P, a.N = num_particles, a.H = (int32_t)H;
  a.params = params.typed_data();
  a.y0 = y0.typed_data();
  a.u = u.typed_data();
  a.u_stride_p = sp, a.u_stride_t = st, a.u_stride_j = 1;
  a.dt = dt.typed_data();
  a.key = key.typed_data();
  a.key_stride_p = nkeys == 1 ? 0 : 2;
  a.rng_mode = rng_mode, a.noise_mode = noise_mode, a.solver = solver;
  a.xevol = xevol->typed_data();
  a.x_rows = md.n_x;
  return from_status(sde4_rollout(&md, &a, stream), "sde4_rollout");
}

// ---- sde4_cost_grad -------------------------------------------------------------------------------------------------
// args   params f32[NP], y0 f32[P, n_x], opt f32[P, H, n_opt], dt f32[H], xref f32[P | 1, H + 1, n_x], key u32[P | 1, 2]
// attrs  model_desc, stage_desc, terminal_desc (empty = none), num_particles, rng_mode, noise_mode
// result cost f32[P], grad f32[P, H, n_opt], xtraj f32[H + 1, n_x + 1, P * N], workspace u8[>= sde4_cost_workspace_bytes]
ffi::Error CostGradImpl(cudaStream_t stream, ffi::Span<const uint8_t> model_desc, ffi::Span<const uint8_t> stage_desc,
                        ffi::Span<const uint8_t> terminal_desc, int32_t num_particles, int32_t rng_mode, int32_t noise_mode,
                        ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> y0, ffi::Buffer<ffi::F32> opt,
                        ffi::Buffer<ffi::F32> dt, ffi::Buffer<ffi::F32> xref, ffi::Buffer<ffi::U32> key,
                        ffi::ResultBuffer<ffi::F32> cost, ffi::ResultBuffer<ffi::F32> grad, ffi::ResultBuffer<ffi::F32> xtraj,
                        ffi::ResultBuffer<ffi::U8> workspace) {
  sde4_model_desc md;
  sde4_cost_desc stage, terminal;
  if (!load_pod(model_desc, &md)) return ffi::Error::InvalidArgument("sde4_cost_grad: model_desc has the wrong size (ABI)");
  if (!load_pod(stage_desc, &stage)) return ffi::Error::InvalidArgument("sde4_cost_grad: stage_desc has the wrong size (ABI)");
  const bool has_terminal = terminal_desc.size() != 0;
  if (has_terminal && !load_pod(terminal_desc, &terminal))
    return ffi::Error::InvalidArgument("sde4_cost_grad: terminal_desc has the wrong size (ABI)");
  if (y0.dimensions().size() != 2 || y0.dimensions()[1] != md.n_x)
    return ffi::Error::InvalidArgument("sde4_cost_grad: y0 must be [P, n_x]");
  const int64_t P = y0.dimensions()[0];
  const int64_t n_opt = md.n_u + stage.n_slack;
  if (opt.dimensions().size() != 3 || opt.dimensions()[0] != P || opt.dimensions()[2] != n_opt)
    return ffi::Error::InvalidArgument("sde4_cost_grad: opt must be [P, H, n_u + n_slack]");
  const int64_t H = opt.dimensions()[1];
  if ((int64_t)dt.element_count() != H) return ffi::Error::InvalidArgument("sde4_cost_grad: len(dt) must be H");
  const int64_t xrows = xref.element_count() / ((H + 1) * md.n_x);
  if ((int64_t)xref.element_count() != xrows * (H + 1) * md.n_x || (xrows != 1 && xrows != P))
    return ffi::Error::InvalidArgument("sde4_cost_grad: xref must be [P | 1, H + 1, n_x]");
  const int64_t nkeys = key.element_count() / 2;
  if (noise_mode == SDE4_NOISE_THREEFRY && nkeys != 1 && nkeys != P)
    return ffi::Error::InvalidArgument("sde4_cost_grad: one key or one key per problem");
  const int64_t G = P * num_particles;
  if ((int64_t)cost->element_count() != P || grad->element_count() != opt.element_count() ||
      (int64_t)xtraj->element_count() != (H + 1) * (md.n_x + 1) * G)
    return ffi::Error::InvalidArgument("sde4_cost_grad: result shapes (cost[P], grad like opt, xtraj[H + 1, n_x + 1, P * N])");
  sde4_cost_args a;
  std::memset(&a, 0, sizeof(a));
  a.struct_size = (int32_t)sizeof(a);
  a.P = (int32_t)P, a.N = num_particles, a.H = (int32_t)H;
  a.params = params.typed_data();
  a.y0 = y0.typed_data();
  a.opt = opt.typed_data();
  a.opt_stride_p = H * n_opt, a.opt_stride_t = n_opt, a.opt_stride_j = 1;
  a.dt = dt.typed_data();
  a.xref = xref.typed_data();
  a.xref_stride_p = xrows == 1 ? 0 : (H + 1) * md.n_x;
  a.key = key.typed_data();
  a.key_stride_p = nkeys == 1 ? 0 : 2;
  a.rng_mode = rng_mode, a.noise_mode = noise_mode;
  a.stage = &stage;
  a.terminal = has_terminal ? &terminal : nullptr;
  a.cost = cost->typed_data();
  a.grad = grad->typed_data();
  a.xtraj = xtraj->typed_data();
  // XLA allocations are at least 256-byte aligned; the launcher checks it and the size
  a.workspace = workspace->typed_data();
  a.workspace_bytes = workspace->element_count();
  return from_status(sde4_cost_grad(&md, &a, stream), "sde4_cost_grad");
}

// ---- sde4_loss_grad -------------------------------------------------------------------------------------------------
// args   params f32[NP], ymeas f32[B, H + 1, n_x], u f32[B, H, n_u], dt f32[H], key u32[B, 2] (= split(rng, B))
// attrs  model_desc, data_cost_desc (QUADRATIC), num_particles, rng_mode, noise_mode
// result loss f32[B], grad_params f32[NP] (d mean_b loss / d params, packed layout), workspace u8[>= sde4_loss_workspace_bytes]
ffi::Error LossGradImpl(cudaStream_t stream, ffi::Span<const uint8_t> model_desc, ffi::Span<const uint8_t> data_cost_desc,
                        int32_t num_particles, int32_t rng_mode, int32_t noise_mode, ffi::Buffer<ffi::F32> params,
                        ffi::Buffer<ffi::F32> ymeas, ffi::Buffer<ffi::F32> u, ffi::Buffer<ffi::F32> dt,
                        ffi::Buffer<ffi::U32> key, ffi::ResultBuffer<ffi::F32> loss, ffi::ResultBuffer<ffi::F32> grad_params,
                        ffi::ResultBuffer<ffi::U8> workspace) {
  sde4_model_desc md;
  sde4_cost_desc dc;
  if (!load_pod(model_desc, &md)) return ffi::Error::InvalidArgument("sde4_loss_grad: model_desc has the wrong size (ABI)");
  if (!load_pod(data_cost_desc, &dc)) return ffi::Error::InvalidArgument("sde4_loss_grad: data_cost_desc has the wrong size (ABI)");
  if (ymeas.dimensions().size() != 3 || ymeas.dimensions()[2] != md.n_x || u.dimensions().size() != 3 ||
      u.dimensions()[2] != md.n_u || u.dimensions()[0] != ymeas.dimensions()[0] || u.dimensions()[1] + 1 != ymeas.dimensions()[1])
    return ffi::Error::InvalidArgument("sde4_loss_grad: ymeas must be [B, H + 1, n_x], u [B, H, n_u]");
  const int64_t B = u.dimensions()[0], H = u.dimensions()[1];
  if ((int64_t)dt.element_count() != H || (int64_t)key.element_count() != 2 * B || (int64_t)loss->element_count() != B ||
      grad_params->element_count() != params.element_count())
    return ffi::Error::InvalidArgument("sde4_loss_grad: dt[H], key[B, 2], loss[B], grad_params like params");
  sde4_cost_args a;
  std::memset(&a, 0, sizeof(a));
  a.struct_size = (int32_t)sizeof(a);
  a.P = (int32_t)B, a.N = num_particles, a.H = (int32_t)H;
  a.params = params.typed_data();
  a.opt = u.typed_data();
  a.opt_stride_p = H * md.n_u, a.opt_stride_t = md.n_u, a.opt_stride_j = 1;
  a.dt = dt.typed_data();
  a.xref = ymeas.typed_data();
  a.xref_stride_p = (H + 1) * md.n_x;
  a.key = key.typed_data();
  a.key_stride_p = 2;
  a.rng_mode = rng_mode, a.noise_mode = noise_mode;
  a.stage = &dc;
  a.cost = loss->typed_data();
  a.workspace = workspace->typed_data();
  a.workspace_bytes = workspace->element_count();
  // y0 must be contiguous [B][n_x] for the launcher: the first n_x floats of each element of ymeas are not (stride
  // (H + 1) n_x), so they are gathered into the head of the scratch buffer first (B * n_x floats, 256-byte rounded)
  const size_t y0_bytes = ((size_t)B * md.n_x * sizeof(float) + 255) & ~(size_t)255;
  if (a.workspace_bytes <= y0_bytes) return ffi::Error::InvalidArgument("sde4_loss_grad: workspace too small");
  float* y0 = reinterpret_cast<float*>(workspace->typed_data());
  if (cudaMemcpy2DAsync(y0, md.n_x * sizeof(float), ymeas.typed_data(), (size_t)(H + 1) * md.n_x * sizeof(float),
                        md.n_x * sizeof(float), (size_t)B, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return ffi::Error::Internal("sde4_loss_grad: gathering y0 failed");
  a.y0 = y0;
  a.workspace = workspace->typed_data() + y0_bytes;
  a.workspace_bytes -= y0_bytes;
  return from_status(sde4_loss_grad(&md, &a, grad_params->typed_data(), stream), "sde4_loss_grad");
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(Sde4Rollout, RolloutImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<ffi::Span<const uint8_t>>("model_desc")
                                  .Attr<int32_t>("num_particles")
                                  .Attr<int32_t>("rng_mode")
                                  .Attr<int32_t>("noise_mode")
                                  .Attr<int32_t>("solver")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // params
                                  .Arg<ffi::Buffer<ffi::F32>>()   // y0
                                  .Arg<ffi::Buffer<ffi::F32>>()   // u
                                  .Arg<ffi::Buffer<ffi::F32>>()   // dt
                                  .Arg<ffi::Buffer<ffi::U32>>()   // key
                                  .Ret<ffi::Buffer<ffi::F32>>()); // xevol

XLA_FFI_DEFINE_HANDLER_SYMBOL(Sde4CostGrad, CostGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<ffi::Span<const uint8_t>>("model_desc")
                                  .Attr<ffi::Span<const uint8_t>>("stage_desc")
                                  .Attr<ffi::Span<const uint8_t>>("terminal_desc")
                                  .Attr<int32_t>("num_particles")
                                  .Attr<int32_t>("rng_mode")
                                  .Attr<int32_t>("noise_mode")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // params
                                  .Arg<ffi::Buffer<ffi::F32>>()   // y0
                                  .Arg<ffi::Buffer<ffi::F32>>()   // opt
                                  .Arg<ffi::Buffer<ffi::F32>>()   // dt
                                  .Arg<ffi::Buffer<ffi::F32>>()   // xref
                                  .Arg<ffi::Buffer<ffi::U32>>()   // key
                                  .Ret<ffi::Buffer<ffi::F32>>()   // cost
                                  .Ret<ffi::Buffer<ffi::F32>>()   // grad
                                  .Ret<ffi::Buffer<ffi::F32>>()   // xtraj
                                  .Ret<ffi::Buffer<ffi::U8>>());  // workspace (scratch)

XLA_FFI_DEFINE_HANDLER_SYMBOL(Sde4LossGrad, LossGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<ffi::Span<const uint8_t>>("model_desc")
                                  .Attr<ffi::Span<const uint8_t>>("data_cost_desc")
                                  .Attr<int32_t>("num_particles")
                                  .Attr<int32_t>("rng_mode")
                                  .Attr<int32_t>("noise_mode")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // params
                                  .Arg<ffi::Buffer<ffi::F32>>()   // ymeas
                                  .Arg<ffi::Buffer<ffi::F32>>()   // u
                                  .Arg<ffi::Buffer<ffi::F32>>()   // dt
                                  .Arg<ffi::Buffer<ffi::U32>>()   // key
                                  .Ret<ffi::Buffer<ffi::F32>>()   // loss
                                  .Ret<ffi::Buffer<ffi::F32>>()   // grad_params
                                  .Ret<ffi::Buffer<ffi::U8>>());  // workspace (scratch)

#endif  // SDE4_HAVE_XLA_FFI

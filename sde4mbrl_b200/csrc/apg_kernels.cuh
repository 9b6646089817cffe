// apg_kernels.cuh -- per-problem bookkeeping of the batched accelerated proximal gradient solver.
// Restates sde4mbrl/apg.py for P independent problems with masks instead of lax.cond / lax.scan:
//   init_apg :53-113, one_step_apg :164-274, armijo_line_search :332-379, update_search :312-321,
//   wolfe_cond_violated :324-329, _while_loop :451-464.
// Every [H][n_opt][*] array is problem-minor.  The vector kernels run on a 2-D grid: x covers the problems,
// y covers chunks of APG_CH consecutive elements e = t*n_opt + j, so a warp touches full 128 B lines and the
// serial work per thread is APG_CH elements; the (cheap) per-problem scalar logic is recomputed by every
// chunk and written by chunk 0 only.  Kernels that both read and write per-problem scalars are split into a
// vector kernel and a one-thread-per-problem `_fin` kernel, which also sums the |grad|^2 partials of the
// chunks in a fixed order.
#pragma once
#include "common.cuh"

namespace sde4 {

constexpr int APG_CH = 4;

__global__ void k_apg_init(const __grid_constant__ sde4_apg_params prm, const __grid_constant__ sde4_apg_state st,
                           const float* __restrict__ x0, const float* __restrict__ grad0, float* __restrict__ gsq_part) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  const int P = prm.P, E = prm.H * prm.n_opt;
  const int e0 = blockIdx.y * APG_CH, e1 = min(e0 + APG_CH, E);
  float gsq = 0.0f;
  for (int e = e0; e < e1; ++e) {
    const float x = x0[(size_t)e * P + p], g = grad0[(size_t)e * P + p];
    st.x_opt[(size_t)e * P + p] = x;
    st.xk[(size_t)e * P + p] = x;
    st.yk[(size_t)e * P + p] = x;
    st.grad_yk[(size_t)e * P + p] = g;
    gsq = fmaf(g, g, gsq);
  }
  gsq_part[(size_t)blockIdx.y * P + p] = gsq;
}

__global__ void k_apg_init_fin(const __grid_constant__ sde4_apg_params prm, const __grid_constant__ sde4_apg_state st,
                               const float* cost0, const float* momentum, const float* stepsize, float init_stepsize,
                               float beta_init, const float* __restrict__ gsq_part, int chunks) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  float gsq = 0.0f;
  for (int c = 0; c < chunks; ++c) gsq += gsq_part[(size_t)c * prm.P + p];
  const float mom = momentum ? momentum[p] : beta_init;
  const float s = stepsize ? stepsize[p] : init_stepsize;
  st.num_steps[p] = 1;
  st.num_iter_no_improv[p] = 0;
  st.not_done[p] = 1;
  st.momentum[p] = mom;
  st.avg_momentum[p] = mom;
  st.avg_linesearch[p] = 1.0f;
  st.stepsize[p] = s;
  st.avg_stepsize[p] = s;
  st.opt_cost[p] = cost0[p];
  st.curr_cost[p] = cost0[p];
  st.init_cost[p] = cost0[p];
  st.grad_sqr[p] = gsq;
}

// step sizes of the backtracking sequence: s_0 = reset(stepsize), s_k = min(s_{k-1}*dec, max)
SDE4_DEV float apg_first_step(const sde4_apg_params& prm, float s) {
  if (prm.reset_increase) s = fminf(s * prm.increase_factor, prm.max_stepsize);  // apg.py:179-184
  return s;
}

__global__ void k_apg_candidates(const __grid_constant__ sde4_apg_params prm,
                                 const __grid_constant__ sde4_apg_state st, float* __restrict__ cand,
                                 float* __restrict__ svals, int32_t* __restrict__ xmask) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  if (xmask && blockIdx.y == 0) xmask[p] = 0;  // "x must be re-evaluated" flags, raised by k_apg_select
  const int P = prm.P, E = prm.H * prm.n_opt, K = prm.maxls + 1;
  const int e0 = blockIdx.y * APG_CH, e1 = min(e0 + APG_CH, E);
  float y[APG_CH], g[APG_CH];
#pragma unroll
  for (int i = 0; i < APG_CH; ++i) {
    const int e = min(e0 + i, E - 1);
    y[i] = st.yk[(size_t)e * P + p];
    g[i] = st.grad_yk[(size_t)e * P + p];
  }
  float s = apg_first_step(prm, st.stepsize[p]);
  for (int k = 0; k < K; ++k) {
    if (k > 0) s = fminf(s * prm.decrease_factor, prm.max_stepsize);  // update_search, apg.py:318
    if (blockIdx.y == 0) svals[(size_t)k * P + p] = s;
#pragma unroll
    for (int i = 0; i < APG_CH; ++i)
      if (e0 + i < e1) cand[(size_t)(e0 + i) * K * P + (size_t)k * P + p] = y[i] - s * g[i];  // apg.py:319,358
  }
}

__global__ void k_apg_select(const __grid_constant__ sde4_apg_params prm, const __grid_constant__ sde4_apg_state st,
                             const float* __restrict__ cand, const float* __restrict__ svals,
                             const float* __restrict__ cand_cost, float* __restrict__ xv, float* __restrict__ sel_step,
                             int32_t* __restrict__ sel_ls, int32_t* __restrict__ xmask, float* __restrict__ xv_cost) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  const int P = prm.P, E = prm.H * prm.n_opt, K = prm.maxls + 1;
  const int e0 = blockIdx.y * APG_CH, e1 = min(e0 + APG_CH, E);
  const float f_cur = st.curr_cost[p], gsq = st.grad_sqr[p];
  const float eps = 1.1920929e-07f;  // finfo(float32).eps, apg.py:328
  int ks = prm.maxls;
  for (int k = 0; k < prm.maxls; ++k) {
    const float s = svals[(size_t)k * P + p], f = cand_cost[(size_t)k * P + p];
    const bool violated = s * prm.coef * gsq > f_cur - f + eps;  // wolfe_cond_violated
    if (!violated) {
      ks = k;
      break;
    }
  }
  if (blockIdx.y == 0) {
    sel_step[p] = svals[(size_t)ks * P + p];
    sel_ls[p] = 1 + ks;  // num_iter of armijo_line_search
    if (xmask) {
      // cost(x) = cost of the accepted candidate when the prox leaves it untouched: reuse it (the x / v launch then
      // skips x); the fall-back candidate maxls was never evaluated, so x has to be
      xv_cost[p] = cand_cost[(size_t)(ks < prm.maxls ? ks : 0) * P + p];
      if (ks >= prm.maxls) atomicOr(xmask + p, 1);
    }
  }
  const float mom = st.momentum[p];
  bool moved = false;
  for (int e = e0; e < e1; ++e) {
    float x = cand[(size_t)e * K * P + (size_t)ks * P + p];
    if (prm.has_prox) {
      const int j = e % prm.n_opt;
      const float xc = fminf(fmaxf(x, prm.lb[j]), prm.ub[j]);  // constr_vect, nsde.py:1422
      moved = moved || (xc != x);
      x = xc;
    }
    const float xp = st.xk[(size_t)e * P + p];
    xv[(size_t)e * 2 * P + p] = x;
    xv[(size_t)e * 2 * P + P + p] = xp + mom * (x - xp);  // vcurr, apg.py:215
  }
  if (xmask && moved) atomicOr(xmask + p, 1);  // (an OR: order independent, deterministic)
}

// ynext = where(cost_x <= cost_v, x, v).  When the x / v evaluation also produced gradients (xv_grad != nullptr:
// ONE value+gradient launch over the 2P problems instead of a value launch followed by a gradient launch of the
// chosen point -- same arithmetic work, one sequential rollout less per iteration) the chosen gradient is copied too.
__global__ void k_apg_pick(const __grid_constant__ sde4_apg_params prm, const float* __restrict__ xv,
                           const float* __restrict__ xv_cost, float* __restrict__ y, float* __restrict__ y_cost,
                           int32_t* __restrict__ pick, const float* __restrict__ xv_grad, float* __restrict__ y_grad) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  const int P = prm.P, E = prm.H * prm.n_opt;
  const int e0 = blockIdx.y * APG_CH, e1 = min(e0 + APG_CH, E);
  const float cx = xv_cost[p], cv = xv_cost[P + p];
  const int x_le_v = cx <= cv;  // apg.py:219
  if (blockIdx.y == 0) {
    pick[p] = x_le_v;
    y_cost[p] = x_le_v ? cx : cv;
  }
  for (int e = e0; e < e1; ++e) {
    const size_t src = (size_t)e * 2 * P + (x_le_v ? 0 : P) + p;
    y[(size_t)e * P + p] = xv[src];
    if (xv_grad) y_grad[(size_t)e * P + p] = xv_grad[src];
  }
}

// ---- speculative iteration (latency-bound sizes): every point the iteration can end at is laid out BEFORE the line
// search is known, so that ONE value+gradient launch replaces "candidate values -> select -> x / v value+gradient":
//   spec[H][n_opt][Q*P]: blocks of P problems  [ x_0..x_maxls | v_0..v_maxls | (has_prox) raw_0..raw_{maxls-1} ]
// with x_k = prox(raw_k), v_k = xk + momentum (x_k - xk)   (apg.py:201-215; same expressions as k_apg_select).
// The Armijo test needs the cost of the UN-projected candidate (apg.py:362-379), hence the raw block.
__global__ void k_apg_spec_points(const __grid_constant__ sde4_apg_params prm, const __grid_constant__ sde4_apg_state st,
                                  const float* __restrict__ cand, float* __restrict__ spec) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  const int P = prm.P, E = prm.H * prm.n_opt, K = prm.maxls + 1;
  const int Q = 2 * K + (prm.has_prox ? prm.maxls : 0);
  const int e0 = blockIdx.y * APG_CH, e1 = min(e0 + APG_CH, E);
  const float mom = st.momentum[p];
  for (int e = e0; e < e1; ++e) {
    const float xp = st.xk[(size_t)e * P + p];
    const int j = e % prm.n_opt;
    float* row = spec + (size_t)e * Q * P + p;
    for (int k = 0; k < K; ++k) {
      float x = cand[(size_t)e * K * P + (size_t)k * P + p];
      if (prm.has_prox) {
        if (k < prm.maxls) row[(size_t)(2 * K + k) * P] = x;
        x = fminf(fmaxf(x, prm.lb[j]), prm.ub[j]);
      }
      row[(size_t)k * P] = x;
      row[(size_t)(K + k) * P] = xp + mom * (x - xp);
    }
  }
}

// after k_apg_select: cost and gradient of [xcurr | vcurr] are those of the speculated points of the accepted step
__global__ void k_apg_spec_gather(const __grid_constant__ sde4_apg_params prm, const int32_t* __restrict__ sel_ls,
                                  const float* __restrict__ spec_cost, const float* __restrict__ spec_grad,
                                  float* __restrict__ xv_cost, float* __restrict__ xv_grad) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  const int P = prm.P, E = prm.H * prm.n_opt, K = prm.maxls + 1;
  const int Q = 2 * K + (prm.has_prox ? prm.maxls : 0);
  const int e0 = blockIdx.y * APG_CH, e1 = min(e0 + APG_CH, E);
  const int ks = sel_ls[p] - 1;
  if (blockIdx.y == 0) {
    xv_cost[p] = spec_cost[(size_t)ks * P + p];
    xv_cost[P + p] = spec_cost[(size_t)(K + ks) * P + p];
  }
  for (int e = e0; e < e1; ++e) {
    const float* row = spec_grad + (size_t)e * Q * P + p;
    xv_grad[(size_t)e * 2 * P + p] = row[(size_t)ks * P];
    xv_grad[(size_t)e * 2 * P + P + p] = row[(size_t)(K + ks) * P];
  }
}

// progressive line search (apg.py:362-379 evaluated one step size at a time over the whole batch)
__global__ void k_apg_ls_step(const __grid_constant__ sde4_apg_params prm, const __grid_constant__ sde4_apg_state st,
                              const float* __restrict__ svals, const float* __restrict__ cand_cost, int k,
                              int32_t* __restrict__ active) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  const int on = k == 0 ? st.not_done[p] : active[p];
  if (!on) {
    active[p] = 0;
    return;
  }
  const float s = svals[(size_t)k * prm.P + p], f = cand_cost[(size_t)k * prm.P + p];
  const bool violated = s * prm.coef * st.grad_sqr[p] > st.curr_cost[p] - f + 1.1920929e-07f;
  active[p] = violated ? 1 : 0;
}

// vector part of the state update: reads the OLD per-problem scalars, writes vectors and |grad|^2 partials
__global__ void k_apg_update(const __grid_constant__ sde4_apg_params prm, const __grid_constant__ sde4_apg_state st,
                             const float* __restrict__ xv, const float* __restrict__ y,
                             const float* __restrict__ y_cost, const float* __restrict__ y_grad,
                             float* __restrict__ gsq_part) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  if (!st.not_done[p]) return;  // the scan body is the identity once the condition is false (apg.py:451-464)
  const int P = prm.P, E = prm.H * prm.n_opt;
  const int e0 = blockIdx.y * APG_CH, e1 = min(e0 + APG_CH, E);
  const bool improv = st.opt_cost[p] > y_cost[p];
  float gsq = 0.0f;
  for (int e = e0; e < e1; ++e) {
    const float g = y_grad[(size_t)e * P + p], yy = y[(size_t)e * P + p];
    gsq = fmaf(g, g, gsq);
    st.grad_yk[(size_t)e * P + p] = g;
    st.yk[(size_t)e * P + p] = yy;
    st.xk[(size_t)e * P + p] = xv[(size_t)e * 2 * P + p];
    if (improv) st.x_opt[(size_t)e * P + p] = yy;
  }
  gsq_part[(size_t)blockIdx.y * P + p] = gsq;
}

__global__ void k_apg_update_fin(const __grid_constant__ sde4_apg_params prm, const __grid_constant__ sde4_apg_state st,
                                 const float* y_cost, const int32_t* pick, const float* sel_step, const int32_t* sel_ls,
                                 const float* __restrict__ gsq_part, int chunks) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  if (!st.not_done[p]) return;
  const int ns = st.num_steps[p];
  const float cost_y = y_cost[p], opt_cost = st.opt_cost[p], mom = st.momentum[p];
  float moment_next;
  if (!prm.use_moment_scale) {
    moment_next = (float)ns / ((float)ns + 3.0f);  // apg.py:229
  } else {
    moment_next = pick[p] ? prm.moment_scale * mom : fminf(mom / prm.moment_scale, 1.0f);
  }
  bool stop_not_sat = fabsf(cost_y - opt_cost) > prm.atol + prm.rtol * fabsf(opt_cost);  // apg.py:238
  const bool improv = opt_cost > cost_y;
  const int noimp = improv ? 0 : st.num_iter_no_improv[p] + 1;
  stop_not_sat = stop_not_sat && (noimp < prm.max_no_improvement_iter);
  float gsq = 0.0f;
  for (int c = 0; c < chunks; ++c) gsq += gsq_part[(size_t)c * prm.P + p];
  const float tot = (float)(ns + 1), s = sel_step[p];
  st.avg_linesearch[p] = ((float)ns * st.avg_linesearch[p] + (float)sel_ls[p]) / tot;
  st.avg_stepsize[p] = ((float)ns * st.avg_stepsize[p] + s) / tot;
  st.avg_momentum[p] = ((float)ns * st.avg_momentum[p] + moment_next) / tot;
  st.num_steps[p] = ns + 1;
  st.num_iter_no_improv[p] = noimp;
  st.momentum[p] = moment_next;
  st.stepsize[p] = s;
  if (improv) st.opt_cost[p] = cost_y;
  st.curr_cost[p] = cost_y;
  st.grad_sqr[p] = gsq;
  st.not_done[p] = stop_not_sat ? 1 : 0;
}

}  // namespace sde4

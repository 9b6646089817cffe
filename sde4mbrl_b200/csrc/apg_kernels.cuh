// apg_kernels.cuh -- per-problem bookkeeping of the batched accelerated proximal gradient solver.
// Restates sde4mbrl/apg.py for P independent problems with masks instead of lax.cond / lax.scan:
//   init_apg :53-113, one_step_apg :164-274, armijo_line_search :332-379, update_search :312-321,
//   wolfe_cond_violated :324-329, _while_loop :451-464.
// One thread per problem; every [H][n_opt][*] array is problem-minor, so a warp touches full lines.
#pragma once
#include "common.cuh"

namespace sde4 {

__global__ void k_apg_init(const __grid_constant__ sde4_apg_params prm, const __grid_constant__ sde4_apg_state st,
                           const float* x0, const float* cost0, const float* grad0, const float* momentum,
                           const float* stepsize, float init_stepsize, float beta_init) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  const int P = prm.P, E = prm.H * prm.n_opt;
  float gsq = 0.0f;
  for (int e = 0; e < E; ++e) {
    const float x = x0[(size_t)e * P + p], g = grad0[(size_t)e * P + p];
    st.x_opt[(size_t)e * P + p] = x;
    st.xk[(size_t)e * P + p] = x;
    st.yk[(size_t)e * P + p] = x;
    st.grad_yk[(size_t)e * P + p] = g;
    gsq = fmaf(g, g, gsq);
  }
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
                                 const __grid_constant__ sde4_apg_state st, float* cand, float* svals) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  const int P = prm.P, E = prm.H * prm.n_opt, K = prm.maxls + 1;
  float s = apg_first_step(prm, st.stepsize[p]);
  for (int k = 0; k < K; ++k) {
    if (k > 0) s = fminf(s * prm.decrease_factor, prm.max_stepsize);  // update_search, apg.py:318
    svals[(size_t)k * P + p] = s;
    for (int e = 0; e < E; ++e) {
      const float y = st.yk[(size_t)e * P + p], g = st.grad_yk[(size_t)e * P + p];
      cand[(size_t)e * K * P + (size_t)k * P + p] = y - s * g;  // apg.py:319,358
    }
  }
}

__global__ void k_apg_select(const __grid_constant__ sde4_apg_params prm, const __grid_constant__ sde4_apg_state st,
                             const float* cand, const float* svals, const float* cand_cost, float* xv,
                             float* sel_step, int32_t* sel_ls) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  const int P = prm.P, E = prm.H * prm.n_opt, K = prm.maxls + 1;
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
  sel_step[p] = svals[(size_t)ks * P + p];
  sel_ls[p] = 1 + ks;  // num_iter of armijo_line_search
  const float mom = st.momentum[p];
  for (int e = 0; e < E; ++e) {
    float x = cand[(size_t)e * K * P + (size_t)ks * P + p];
    if (prm.has_prox) {
      const int j = e % prm.n_opt;
      x = fminf(fmaxf(x, prm.lb[j]), prm.ub[j]);  // constr_vect, nsde.py:1422
    }
    const float xp = st.xk[(size_t)e * P + p];
    xv[(size_t)e * 2 * P + p] = x;
    xv[(size_t)e * 2 * P + P + p] = xp + mom * (x - xp);  // vcurr, apg.py:215
  }
}

__global__ void k_apg_pick(const __grid_constant__ sde4_apg_params prm, const float* xv, const float* xv_cost,
                           float* y, float* y_cost, int32_t* pick) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  const int P = prm.P, E = prm.H * prm.n_opt;
  const float cx = xv_cost[p], cv = xv_cost[P + p];
  const int x_le_v = cx <= cv;  // apg.py:219
  pick[p] = x_le_v;
  y_cost[p] = x_le_v ? cx : cv;
  for (int e = 0; e < E; ++e) y[(size_t)e * P + p] = xv[(size_t)e * 2 * P + (x_le_v ? 0 : P) + p];
}

__global__ void k_apg_update(const __grid_constant__ sde4_apg_params prm, const __grid_constant__ sde4_apg_state st,
                             const float* xv, const float* y, const float* y_cost, const float* y_grad,
                             const int32_t* pick, const float* sel_step, const int32_t* sel_ls) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= prm.P) return;
  if (!st.not_done[p]) return;  // the scan body is the identity once the condition is false (apg.py:451-464)
  const int P = prm.P, E = prm.H * prm.n_opt;
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
  for (int e = 0; e < E; ++e) {
    const float g = y_grad[(size_t)e * P + p], yy = y[(size_t)e * P + p];
    gsq = fmaf(g, g, gsq);
    st.grad_yk[(size_t)e * P + p] = g;
    st.yk[(size_t)e * P + p] = yy;
    st.xk[(size_t)e * P + p] = xv[(size_t)e * 2 * P + p];
    if (improv) st.x_opt[(size_t)e * P + p] = yy;
  }
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

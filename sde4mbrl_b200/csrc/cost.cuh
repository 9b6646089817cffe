// cost.cuh -- MPC stage / terminal cost and its gradient, evaluated per particle per step.
//
//   one_step_cost_f        sde4mbrlExamples/rotor_uav/sde_mpc_design.py:73-132
//   constr_cost_noprox     sde4mbrl/nsde.py:1400-1410
//   constr_cost_withprox   sde4mbrl/nsde.py:1413-1419
//   (the generic QUADRATIC kind is ours: the reference ships no cartpole / MSD MPC cost)
#pragma once
#include "models.cuh"

namespace sde4 {

// value of the shaping  e -> sig ? sigmoid(e*w - 7)*wsig : e*w   and d/d(e)
template <bool PLAIN = false>
SDE4_DEV float shaped(float e, float w, float wsig, bool sig, float& dde) {
  const float v = e * w;
  if constexpr (PLAIN) {
    dde = w;
    return v;
  }
  if (__builtin_expect(sig, 0)) {  // warp-uniform (descriptor flag): a real branch, not a select
    const float s = sigmoidf_(v - 7.0f);
    dde = s * (1.0f - s) * wsig * w;
    return s * wsig;
  }
  dde = w;
  return v;
}

// Stage cost c(x, u, xref) and (optionally) its gradient scaled by `wt`:
//   gx += wt * dc/dx,  gu += wt * dc/du.
// PLAIN: the caller guarantees the unshaped cost that goes with the model's state dimension -- 13 states: kind ==
// ROTOR_TRACK without sigmoid / tanh shaping; 4 states: QUADRATIC on the cartpole features; otherwise: QUADRATIC on
// the state itself -- which makes the evaluation branch free (the lane-split kernels rely on straight-line loop
// bodies, see kernels.cuh).
template <int NX, int NU, bool GRAD, bool PLAIN = false>
SDE4_DEV float stage_cost(const sde4_cost_desc& c, const float (&x)[NX], const float (&u)[NU],
                          const float* __restrict__ xref, bool with_u, float wt, float (&gx)[NX], float (&gu)[NU]) {
  constexpr bool PLAIN_ROTOR = PLAIN && NX == 13, PLAIN_QUAD = PLAIN && NX != 13;
  float tot = 0.0f;
  float dx[NX], du[NU];  // d tot / d x, d tot / d u  (before the outer shaping)
#pragma unroll
  for (int i = 0; i < NX; ++i) dx[i] = 0.0f;
#pragma unroll
  for (int j = 0; j < NU; ++j) du[j] = 0.0f;

  if (PLAIN_ROTOR || (!PLAIN && c.kind == SDE4_COST_ROTOR_TRACK)) {
    if constexpr (NX == 13) {
      // p (0-2), v (3-5), w (10-12): squared error times weight
#pragma unroll
      for (int i = 0; i < 13; ++i) {
        if (i >= 6 && i < 10) continue;
        const int grp = i < 3 ? 0 : (i < 6 ? 1 : 3);
        const float diff = x[i] - xref[i];
        float dde;
        tot += shaped<PLAIN>(diff * diff, c.w_x[i], c.w_x_sig[i], (c.sig_mask >> grp) & 1u, dde);
        dx[i] = 2.0f * diff * dde;
      }
      // attitude: (q_ref (x) q^-1)[1:]^2   (sde_mpc_design.py:83)
      {
        const Quat qr{xref[6], xref[7], xref[8], xref[9]};
        const Quat qx{x[6], x[7], x[8], x[9]};
        const Quat r = qmul(qr, qconj(qx));
        float d1, d2, d3;
        tot += shaped<PLAIN>(r.x * r.x, c.w_x[7], c.w_x_sig[7], (c.sig_mask >> 2) & 1u, d1);
        tot += shaped<PLAIN>(r.y * r.y, c.w_x[8], c.w_x_sig[8], (c.sig_mask >> 2) & 1u, d2);
        tot += shaped<PLAIN>(r.z * r.z, c.w_x[9], c.w_x_sig[9], (c.sig_mask >> 2) & 1u, d3);
        if (GRAD) {
          const Quat gr{0.0f, 2.0f * r.x * d1, 2.0f * r.y * d2, 2.0f * r.z * d3};
          const Quat gq = qconj(qmul(qconj(qr), gr));  // r = qr (x) conj(q)
          dx[6] = gq.w;
          dx[7] = gq.x;
          dx[8] = gq.y;
          dx[9] = gq.z;
        }
      }
    }
    if (with_u) {
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        const float diff = u[j] - c.uref[j];
        float dde;
        tot += shaped<PLAIN>(diff * diff, c.w_u[j], c.w_u_sig[j], (c.sig_mask >> 4) & 1u, dde);
        du[j] = 2.0f * diff * dde;
      }
    }
  } else if (PLAIN_QUAD || c.kind == SDE4_COST_QUADRATIC) {
    bool done = false;
    if constexpr (NX == 4) {
      if (PLAIN_QUAD || c.feat == 1) {  // cartpole features [x, xdot, sin th, cos th, thdot]
        float s, co, sr, cr;
        sincosf(x[2], &s, &co);
        sincosf(xref[2], &sr, &cr);
        const float e0 = x[0] - xref[0], e1 = x[1] - xref[1], e2 = s - sr, e3 = co - cr, e4 = x[3] - xref[3];
        tot = c.w_x[0] * e0 * e0 + c.w_x[1] * e1 * e1 + c.w_x[2] * e2 * e2 + c.w_x[3] * e3 * e3 + c.w_x[4] * e4 * e4;
        dx[0] = 2.0f * c.w_x[0] * e0;
        dx[1] = 2.0f * c.w_x[1] * e1;
        dx[2] = 2.0f * (c.w_x[2] * e2 * co - c.w_x[3] * e3 * s);
        dx[3] = 2.0f * c.w_x[4] * e4;
        done = true;
      }
    }
    if (!done) {
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const float diff = x[i] - xref[i];
        tot = fmaf(c.w_x[i] * diff, diff, tot);
        dx[i] = 2.0f * c.w_x[i] * diff;
      }
    }
    if (with_u) {
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        const float diff = u[j] - c.uref[j];
        tot = fmaf(c.w_u[j] * diff, diff, tot);
        du[j] = 2.0f * c.w_u[j] * diff;
      }
    }
  } else {
    return 0.0f;
  }
  float val, outer;
  if (!PLAIN && c.use_res_sig) {  // tanh(tot - lo) * mult   (sde_mpc_design.py:128-130)
    const float th = tanhf_(tot - c.res_sig_lo);
    val = th * c.res_sig_mult;
    outer = (1.0f - th * th) * c.res_sig_mult;
  } else {
    val = tot * c.res_mult;
    outer = c.res_mult;
  }
  if (GRAD) {
    const float s = wt * outer;
#pragma unroll
    for (int i = 0; i < NX; ++i) gx[i] = fmaf(s, dx[i], gx[i]);
    if (with_u) {
#pragma unroll
      for (int j = 0; j < NU; ++j) gu[j] = fmaf(s, du[j], gu[j]);
    }
  }
  return val;
}

// State-bound penalty of one stage.  slack[i] is the slack value paired with state i at this
// stage (WITHPROX only).  Returns the penalty; with GRAD accumulates wt*d/dx into gx and
// writes wt*d/dslack into gs[i].
template <int NX, bool GRAD>
SDE4_DEV float constr_cost(const sde4_cost_desc& c, const float (&x)[NX], const float (&slack)[NX], float wt,
                           float (&gx)[NX], float (&gs)[NX]) {
  float tot = 0.0f;
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    if (GRAD) gs[i] = 0.0f;
    if (c.slack_col[i] < 0) continue;
    if (c.constr_mode == SDE4_CONSTR_WITHPROX) {
      const float dd = x[i] - slack[i] * c.slack_scale[i];
      tot = fmaf(c.pen[i] * dd, dd, tot);
      if (GRAD) {
        const float g = 2.0f * c.pen[i] * dd * c.constr_weight * wt;
        gx[i] += g;
        gs[i] = -g * c.slack_scale[i];
      }
    } else {
      const float hi = x[i] - c.ub[i], lo = c.lb[i] - x[i];
      float g = 0.0f;
      if (hi > 0.0f) {
        tot = fmaf(c.pen[i] * hi, hi, tot);
        g += 2.0f * c.pen[i] * hi;
      }
      if (lo > 0.0f) {
        tot = fmaf(c.pen[i] * lo, lo, tot);
        g -= 2.0f * c.pen[i] * lo;
      }
      if (GRAD) gx[i] += g * c.constr_weight * wt;
    }
  }
  return tot * c.constr_weight;
}

}  // namespace sde4

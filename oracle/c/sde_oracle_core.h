/* sde_oracle_core.h -- CPU restatement (plain C, GCC vector extensions) of the multirotor nSDE
 * hot path: Euler-Maruyama particle rollout + MPC cost + reverse-mode gradient w.r.t. the
 * control sequence.  TEST INFRASTRUCTURE / CPU BASELINE ONLY -- never linked into the product.
 *
 * It follows the reference line by line (paths relative to /root/reference):
 *   euler_maruyama ............ sde4mbrl/sde_solver.py:243-317
 *   diffusion ................. sde4mbrl/nsde.py:298-449
 *   SDERotorModel ............. sde4mbrlExamples/rotor_uav/sde_rotor_model.py:121-458
 *   motor_model ............... sde4mbrlExamples/rotor_uav/motor_models.py:21-95
 *   quaternions ............... sde4mbrlExamples/rotor_uav/utils.py:13-63
 *   compute_cost_function ..... sde4mbrl/nsde.py:1485-1527,1576-1586
 *   one_step_cost_f ........... sde4mbrlExamples/rotor_uav/sde_mpc_design.py:73-132
 *   value_and_grad ............ sde4mbrl/apg.py:82,224 (hand-derived adjoint)
 * PARITY UNPINNED against the reference itself (no jax here); pinned against the torch oracle
 * (float64 autograd) in tests/test_oracle_c.py.
 *
 * Included twice (VL = 8 for AVX2, VL = 16 for AVX-512): particles are processed VL at a time, one
 * SIMD lane per particle, OpenMP over chunks -- i.e. what a good vectorising compiler (XLA:CPU)
 * would make of the reference's vmap over particles.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/sde4b200.h"

#ifndef VL
#error "define VL"
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, CAT(_vl, VL))

typedef float vf __attribute__((vector_size(VL * 4)));
typedef int32_t vi __attribute__((vector_size(VL * 4)));

static inline vf vset(float a) { return (vf){0} + a; }
static inline vf vsel(vi m, vf a, vf b) { return (vf)(((vi)a & m) | ((vi)b & ~m)); }
static inline vf vmaxf(vf a, vf b) { return vsel(a > b, a, b); }
static inline vf vminf(vf a, vf b) { return vsel(a < b, a, b); }

/* expf: Cephes-style, 2 ulp */
static inline vf vexp(vf x) {
  x = vminf(vmaxf(x, vset(-87.0f)), vset(88.0f));
  vf fx = x * 1.44269504088896341f + 0.5f;
  vi n = __builtin_convertvector(fx, vi);
  vf fn = __builtin_convertvector(n, vf);
  vi adj = fn > fx;            /* floor */
  n += adj;                    /* adj is -1 where true */
  fn = __builtin_convertvector(n, vf);
  vf r = x - fn * 0.693359375f;
  r = r - fn * -2.12194440e-4f;
  vf p = vset(1.9875691500E-4f);
  p = p * r + 1.3981999507E-3f;
  p = p * r + 8.3334519073E-3f;
  p = p * r + 4.1665795894E-2f;
  p = p * r + 1.6666665459E-1f;
  p = p * r + 5.0000001201E-1f;
  p = p * r * r + r + 1.0f;
  vi e = (n + 127) << 23;
  return p * (vf)e;
}
static inline vf vsigmoid(vf a) { return 1.0f / (1.0f + vexp(-a)); }
/* the rational float tanh XLA:CPU / Eigen evaluate */
static inline vf vtanh(vf a) {
  vf x = vminf(vmaxf(a, vset(-7.99881172180175781f)), vset(7.99881172180175781f));
  vf x2 = x * x;
  vf p = vset(-2.76076847742355e-16f);
  p = p * x2 + 2.00018790482477e-13f;
  p = p * x2 + -8.60467152213735e-11f;
  p = p * x2 + 5.12229709037114e-08f;
  p = p * x2 + 1.48572235717979e-05f;
  p = p * x2 + 6.37261928875436e-04f;
  p = p * x2 + 4.89352455891786e-03f;
  p = p * x;
  vf q = vset(1.19825839466702e-06f);
  q = q * x2 + 1.18534705686654e-04f;
  q = q * x2 + 2.26843463243900e-03f;
  q = q * x2 + 4.89352518554385e-03f;
  return p / q;
}
static inline void vact(int act, vf a, vf* h, vf* d) {
  if (act == SDE4_ACT_TANH) {
    vf t = vtanh(a);
    *h = t;
    *d = 1.0f - t * t;
  } else if (act == SDE4_ACT_SWISH) {
    vf s = vsigmoid(a);
    *h = a * s;
    *d = s + a * s * (1.0f - s);
  } else {
    *h = a;
    *d = vset(1.0f);
  }
}

#define MAXW 32
static inline int up4i(int n) { return (n + 3) & ~3; }

/* hk.nets.MLP, two hidden layers; packed layout of include/sde4b200.h (w[in][out4], b[out4]) */
typedef struct {
  vf d1[MAXW], d2[MAXW];
} mlp_keep;

static inline void mlp_fwd(const float* W, const sde4_net_desc* nd, const vf* in, vf* out, mlp_keep* k) {
  const int h1p = up4i(nd->h1), h2p = up4i(nd->h2), op = up4i(nd->n_out);
  const float* w0 = W;
  const float* b0 = w0 + nd->n_in * h1p;
  const float* w1 = b0 + h1p;
  const float* b1 = w1 + nd->h1 * h2p;
  const float* w2 = b1 + h2p;            /* transposed: w2T[out][h2p] */
  const float* b2 = w2 + nd->n_out * h2p;
  (void)op;
  vf h1[MAXW], h2[MAXW];
  for (int j = 0; j < nd->h1; ++j) {
    vf a = vset(b0[j]);
    for (int i = 0; i < nd->n_in; ++i) a += in[i] * w0[i * h1p + j];
    vf d;
    vact(nd->act, a, &h1[j], &d);
    if (k) k->d1[j] = d;
  }
  for (int j = 0; j < nd->h2; ++j) {
    vf a = vset(b1[j]);
    for (int i = 0; i < nd->h1; ++i) a += h1[i] * w1[i * h2p + j];
    vf d;
    vact(nd->act, a, &h2[j], &d);
    if (k) k->d2[j] = d;
  }
  for (int j = 0; j < nd->n_out; ++j) {
    vf a = vset(b2[j]);
    for (int i = 0; i < nd->h2; ++i) a += h2[i] * w2[j * h2p + i];
    out[j] = a;
  }
}

static inline void mlp_bwd(const float* W, const sde4_net_desc* nd, const mlp_keep* k, const vf* gout, vf* gin) {
  const int h1p = up4i(nd->h1), h2p = up4i(nd->h2);
  const float* w0 = W;
  const float* w1 = w0 + nd->n_in * h1p + h1p;
  const float* w2 = w1 + nd->h1 * h2p + h2p;
  vf g2[MAXW], g1[MAXW];
  for (int i = 0; i < nd->h2; ++i) {
    vf a = vset(0.0f);
    for (int j = 0; j < nd->n_out; ++j) a += gout[j] * w2[j * h2p + i];
    g2[i] = a * k->d2[i];
  }
  for (int i = 0; i < nd->h1; ++i) {
    vf a = vset(0.0f);
    for (int j = 0; j < nd->h2; ++j) a += g2[j] * w1[i * h2p + j];
    g1[i] = a * k->d1[i];
  }
  for (int i = 0; i < nd->n_in; ++i) {
    vf a = vset(0.0f);
    for (int j = 0; j < nd->h1; ++j) a += g1[j] * w0[i * h1p + j];
    gin[i] = a;
  }
}

static inline int net_size(const sde4_net_desc* nd) {
  if (nd->n_in == 0) return 0;
  return nd->n_in * up4i(nd->h1) + up4i(nd->h1) + nd->h1 * up4i(nd->h2) + up4i(nd->h2) + nd->n_out * up4i(nd->h2) +
         up4i(nd->n_out);
}

typedef struct {
  vf w, x, y, z;
} quat;
static inline quat qmul(quat a, quat b) {
  quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}
static inline quat qconj(quat a) { return (quat){a.w, -a.x, -a.y, -a.z}; }
static inline quat qvec(vf x, vf y, vf z) { return (quat){vset(0.0f), x, y, z}; }
static inline void qacc(quat* a, quat b) {
  a->w += b.w;
  a->x += b.x;
  a->y += b.y;
  a->z += b.z;
}

static const float MIX6[4][6] = {{-1.0f, 1.0f, 0.5f, -0.5f, -0.5f, 0.5f},
                                 {0.0f, 0.0f, -0.866f, 0.866f, -0.866f, 0.866f},
                                 {1.0f, -1.0f, 1.0f, -1.0f, -1.0f, 1.0f},
                                 {1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f}};
static const float MIX4[4][4] = {{-0.707f, 0.707f, 0.707f, -0.707f},
                                 {-0.707f, 0.707f, -0.707f, 0.707f},
                                 {-1.0f, -1.0f, 1.0f, 1.0f},
                                 {1.0f, 1.0f, 1.0f, 1.0f}};
static inline float mixv(int nu, int r, int i) { return nu == 6 ? MIX6[r][i] : MIX4[r][i]; }

typedef struct {
  const sde4_model_desc* d;
  const float *S, *scaler, *Wd, *Wa, *Wb;
  float coeff[SDE4_MAX_NX], offs[SDE4_MAX_NX]; /* exp(scaler)*1e-7 per noise output (nsde.py:413) */
  int nout_idx[SDE4_MAX_NX], n_out, nin_idx[SDE4_MAX_NX], n_in, a_idx[6], b_idx[6];
} rotor_ctx;

static void FN(ctx_init)(rotor_ctx* c, const sde4_model_desc* d, const float* params) {
  c->d = d;
  int off = 0;
  c->S = params + off;
  off += 16;
  c->scaler = params + off;
  off += up4i(d->n_scaler);
  c->Wd = params + off;
  off += net_size(&d->density);
  c->Wa = params + off;
  off += net_size(&d->net_a);
  c->Wb = params + off;
  c->n_out = c->n_in = 0;
  for (int i = 0; i < 16; ++i) {
    if ((d->noise_out_mask >> i) & 1u) c->nout_idx[c->n_out++] = i;
    if ((d->noise_in_mask >> i) & 1u) c->nin_idx[c->n_in++] = i;
  }
  int ka = 0, kb = 0;
  for (int i = 0; i < 6; ++i) {
    if ((d->net_a_in_mask >> i) & 1u) c->a_idx[ka++] = i;
    if ((d->net_b_in_mask >> i) & 1u) c->b_idx[kb++] = i;
  }
  for (int k = 0; k < c->n_out; ++k) {
    c->coeff[k] = d->n_scaler ? expf(c->scaler[k]) * 1e-7f : 0.0f;
    c->offs[k] = d->n_scaler ? expf(c->scaler[c->n_out + k]) * 1e-7f : 0.0f;
  }
}

static inline vf motor_thrust(const float* S, vf u) { return ((S[7] * u + S[8]) * u + S[9]) * u + S[10]; }
static inline vf motor_thrust_du(const float* S, vf u) { return (3.0f * S[7] * u + 2.0f * S[8]) * u + S[9]; }

/* drift f(x,u) (sde_rotor_model.py:427-458); with `vf_ct != NULL` also the VJP: gx += J_x^T vf_ct, gu += J_u^T vf_ct */
static inline void rotor_drift(const rotor_ctx* c, const vf* x, const vf* u, vf* f, const vf* ct, vf* gx, vf* gu) {
  const sde4_model_desc* d = c->d;
  const float* S = c->S;
  const int nu = d->n_u;
  const int drag = d->flags & SDE4_ROTOR_AERO_DRAG;
  const quat q = {x[6], x[7], x[8], x[9]};
  const quat qc = qconj(q);
  const quat V = qvec(x[3], x[4], x[5]);
  const quat t1 = qmul(V, q);
  const quat r1 = qmul(qc, t1); /* v_b = q^-1 (v q) */
  const vf vb[3] = {r1.x, r1.y, r1.z};
  vf in6[6];
  in6[0] = vb[0] * d->inv_state_scaling[3];
  in6[1] = vb[1] * d->inv_state_scaling[4];
  in6[2] = vb[2] * d->inv_state_scaling[5];
  in6[3] = x[10] * d->inv_state_scaling[10];
  in6[4] = x[11] * d->inv_state_scaling[11];
  in6[5] = x[12] * d->inv_state_scaling[12];
  vf Fb[3] = {vset(0.0f), vset(0.0f), vset(0.0f)}, Mt[3] = {vset(0.0f), vset(0.0f), vset(0.0f)};
  mlp_keep ka, kb;
  vf ina[6], inb[6];
  if (d->net_a.n_in) {
    for (int k = 0; k < d->net_a.n_in; ++k) ina[k] = in6[c->a_idx[k]];
    mlp_fwd(c->Wa, &d->net_a, ina, Fb, ct ? &ka : NULL);
  }
  if (drag) {
    Fb[0] += -S[11] * vb[0];
    Fb[1] += -S[12] * vb[1];
    Fb[2] += -S[13] * vb[2] + S[14] * (vb[0] * vb[0] + vb[1] * vb[1]);
  }
  if (d->net_b.n_in) {
    for (int k = 0; k < d->net_b.n_in; ++k) inb[k] = in6[c->b_idx[k]];
    mlp_fwd(c->Wb, &d->net_b, inb, Mt, ct ? &kb : NULL);
  }
  vf T = vset(0.0f), Mm[3] = {vset(0.0f), vset(0.0f), vset(0.0f)};
  for (int i = 0; i < nu; ++i) {
    vf th = motor_thrust(S, u[i]);
    Mm[0] += (mixv(nu, 0, i) * S[4]) * th;
    Mm[1] += (mixv(nu, 1, i) * S[5]) * th;
    Mm[2] += (mixv(nu, 2, i) * S[6]) * th;
    T += mixv(nu, 3, i) * th;
  }
  Fb[2] += T;
  const quat Fq = qvec(Fb[0], Fb[1], Fb[2]);
  const quat t2 = qmul(q, Fq);
  const quat r2 = qmul(t2, qc);
  const float m = S[0], g = d->cfloat[0];
  const float I0 = S[1], I1 = S[2], I2 = S[3];
  const vf w0 = x[10], w1 = x[11], w2 = x[12];
  const vf Iw0 = I0 * w0, Iw1 = I1 * w1, Iw2 = I2 * w2;
  if (f) {
    f[0] = x[3];
    f[1] = x[4];
    f[2] = x[5];
    f[3] = r2.x / m;
    f[4] = r2.y / m;
    f[5] = r2.z / m - g;
    f[10] = (Mm[0] + Mt[0] - (w1 * Iw2 - w2 * Iw1)) / I0;
    f[11] = (Mm[1] + Mt[1] - (w2 * Iw0 - w0 * Iw2)) / I1;
    f[12] = (Mm[2] + Mt[2] - (w0 * Iw1 - w1 * Iw0)) / I2;
    const quat qd = qmul(q, qvec(w0, w1, w2));
    f[6] = 0.5f * qd.w;
    f[7] = 0.5f * qd.x;
    f[8] = 0.5f * qd.y;
    f[9] = 0.5f * qd.z;
  }
  if (!ct) return;
  /* ---------------- reverse mode ---------------- */
  quat gq = {vset(0.0f), vset(0.0f), vset(0.0f), vset(0.0f)};
  vf gw[3] = {vset(0.0f), vset(0.0f), vset(0.0f)};
  gx[3] += ct[0];
  gx[4] += ct[1];
  gx[5] += ct[2];
  const vf gM[3] = {ct[10] / I0, ct[11] / I1, ct[12] / I2};
  {
    const vf c0 = -gM[0], c1 = -gM[1], c2 = -gM[2];
    gw[0] += Iw1 * c2 - Iw2 * c1;
    gw[1] += Iw2 * c0 - Iw0 * c2;
    gw[2] += Iw0 * c1 - Iw1 * c0;
    gw[0] += I0 * (c1 * w2 - c2 * w1);
    gw[1] += I1 * (c2 * w0 - c0 * w2);
    gw[2] += I2 * (c0 * w1 - c1 * w0);
  }
  {
    const quat gqd = {0.5f * ct[6], 0.5f * ct[7], 0.5f * ct[8], 0.5f * ct[9]};
    qacc(&gq, qmul(gqd, qconj(qvec(w0, w1, w2))));
    const quat gOm = qmul(qc, gqd);
    gw[0] += gOm.x;
    gw[1] += gOm.y;
    gw[2] += gOm.z;
  }
  const quat R2 = qvec(ct[3] / m, ct[4] / m, ct[5] / m);
  const quat t2bar = qmul(R2, q);
  const quat gFq = qmul(qc, t2bar);
  const vf gFb[3] = {gFq.x, gFq.y, gFq.z};
  qacc(&gq, qconj(qmul(qconj(t2), R2)));
  qacc(&gq, qmul(t2bar, qconj(Fq)));
  for (int i = 0; i < nu; ++i) {
    vf gth = mixv(nu, 3, i) * gFb[2] + (mixv(nu, 0, i) * S[4]) * gM[0] + (mixv(nu, 1, i) * S[5]) * gM[1] +
             (mixv(nu, 2, i) * S[6]) * gM[2];
    gu[i] += gth * motor_thrust_du(S, u[i]);
  }
  vf gvb[3] = {vset(0.0f), vset(0.0f), vset(0.0f)};
  vf gin6[6] = {vset(0.0f), vset(0.0f), vset(0.0f), vset(0.0f), vset(0.0f), vset(0.0f)};
  if (d->net_a.n_in) {
    vf gi[6];
    mlp_bwd(c->Wa, &d->net_a, &ka, gFb, gi);
    for (int k = 0; k < d->net_a.n_in; ++k) gin6[c->a_idx[k]] += gi[k];
  }
  if (drag) {
    gvb[0] += -S[11] * gFb[0] + 2.0f * S[14] * vb[0] * gFb[2];
    gvb[1] += -S[12] * gFb[1] + 2.0f * S[14] * vb[1] * gFb[2];
    gvb[2] += -S[13] * gFb[2];
  }
  if (d->net_b.n_in) {
    vf gi[6];
    mlp_bwd(c->Wb, &d->net_b, &kb, gM, gi);
    for (int k = 0; k < d->net_b.n_in; ++k) gin6[c->b_idx[k]] += gi[k];
  }
  gvb[0] += gin6[0] * d->inv_state_scaling[3];
  gvb[1] += gin6[1] * d->inv_state_scaling[4];
  gvb[2] += gin6[2] * d->inv_state_scaling[5];
  gw[0] += gin6[3] * d->inv_state_scaling[10];
  gw[1] += gin6[4] * d->inv_state_scaling[11];
  gw[2] += gin6[5] * d->inv_state_scaling[12];
  {
    const quat R1 = qvec(gvb[0], gvb[1], gvb[2]);
    qacc(&gq, qconj(qmul(R1, qconj(t1))));
    const quat t1bar = qmul(q, R1);
    const quat gV = qmul(t1bar, qc);
    gx[3] += gV.x;
    gx[4] += gV.y;
    gx[5] += gV.z;
    qacc(&gq, qmul(qconj(V), t1bar));
  }
  gx[6] += gq.w;
  gx[7] += gq.x;
  gx[8] += gq.y;
  gx[9] += gq.z;
  gx[10] += gw[0];
  gx[11] += gw[1];
  gx[12] += gw[2];
}

/* diffusion sigma(x) (nsde.py:437-449); with ct the VJP gx += J^T ct */
static inline void rotor_diffusion(const rotor_ctx* c, const vf* x, vf* sig, const vf* ct, vf* gx) {
  const sde4_model_desc* d = c->d;
  if (sig)
    for (int i = 0; i < 13; ++i) sig[i] = vset(d->noise_prior[i]);
  if (!d->density.n_in) return;
  vf z[16], dens[1];
  for (int k = 0; k < c->n_in; ++k) z[k] = x[c->nin_idx[k]] * d->inv_red_scale;
  mlp_keep kd;
  mlp_fwd(c->Wd, &d->density, z, dens, ct ? &kd : NULL);
  vf gd[1] = {vset(0.0f)};
  for (int k = 0; k < c->n_out; ++k) {
    const int i = c->nout_idx[k];
    vf e = vsigmoid((d->aggr + c->coeff[k]) * dens[0] - 7.0f + c->offs[k]);
    if (sig) sig[i] *= e;
    if (ct) gd[0] += ct[i] * d->noise_prior[i] * (e * (1.0f - e) * (d->aggr + c->coeff[k]));
  }
  if (!ct) return;
  vf gz[16];
  mlp_bwd(c->Wd, &d->density, &kd, gd, gz);
  for (int k = 0; k < c->n_in; ++k) gx[c->nin_idx[k]] += gz[k] * d->inv_red_scale;
}

static inline vf shaped(vf e, float w, float wsig, int sig, vf* dde) {
  vf v = e * w;
  if (sig) {
    vf s = vsigmoid(v - 7.0f);
    *dde = s * (1.0f - s) * (wsig * w);
    return s * wsig;
  }
  *dde = vset(w);
  return v;
}

/* one_step_cost_f; with grad: gx += wt*dc/dx, gu += wt*dc/du */
static inline vf rotor_cost(const sde4_cost_desc* c, int nu, const vf* x, const vf* u, const float* xref, int with_u,
                            float wt, vf* gx, vf* gu) {
  vf tot = vset(0.0f), dx[13], du[SDE4_MAX_NU];
  for (int i = 0; i < 13; ++i) dx[i] = vset(0.0f);
  for (int i = 0; i < 13; ++i) {
    if (i >= 6 && i < 10) continue;
    const int grp = i < 3 ? 0 : (i < 6 ? 1 : 3);
    vf diff = x[i] - xref[i], dde;
    tot += shaped(diff * diff, c->w_x[i], c->w_x_sig[i], (c->sig_mask >> grp) & 1u, &dde);
    dx[i] = 2.0f * diff * dde;
  }
  {
    const quat qr = {vset(xref[6]), vset(xref[7]), vset(xref[8]), vset(xref[9])};
    const quat qx = {x[6], x[7], x[8], x[9]};
    const quat r = qmul(qr, qconj(qx));
    vf d1, d2, d3;
    const int sg = (c->sig_mask >> 2) & 1u;
    tot += shaped(r.x * r.x, c->w_x[7], c->w_x_sig[7], sg, &d1);
    tot += shaped(r.y * r.y, c->w_x[8], c->w_x_sig[8], sg, &d2);
    tot += shaped(r.z * r.z, c->w_x[9], c->w_x_sig[9], sg, &d3);
    const quat gr = {vset(0.0f), 2.0f * r.x * d1, 2.0f * r.y * d2, 2.0f * r.z * d3};
    const quat gq = qconj(qmul(qconj(qr), gr));
    dx[6] = gq.w;
    dx[7] = gq.x;
    dx[8] = gq.y;
    dx[9] = gq.z;
  }
  if (with_u)
    for (int j = 0; j < nu; ++j) {
      vf diff = u[j] - c->uref[j], dde;
      tot += shaped(diff * diff, c->w_u[j], c->w_u_sig[j], (c->sig_mask >> 4) & 1u, &dde);
      du[j] = 2.0f * diff * dde;
    }
  vf val, outer;
  if (c->use_res_sig) {
    vf th = vtanh(tot - c->res_sig_lo);
    val = th * c->res_sig_mult;
    outer = (1.0f - th * th) * c->res_sig_mult;
  } else {
    val = tot * c->res_mult;
    outer = vset(c->res_mult);
  }
  if (gx) {
    vf s = wt * outer;
    for (int i = 0; i < 13; ++i) gx[i] += s * dx[i];
    if (with_u)
      for (int j = 0; j < nu; ++j) gu[j] += s * du[j];
  }
  return val;
}

/* One chunk of VL particles: forward sweep (checkpoints), backward sweep.
 * noise: dw[H][13][VL] for this chunk.  xs: workspace [(H+1)*13] vf, aux: [H] vf.
 * Accumulates sum over lanes of cost into *cost_sum and of the gradient into grad_sum[H*nu]. */
static void FN(chunk_cost_grad)(const rotor_ctx* c, const sde4_cost_desc* cd, const float* y0, const float* opt,
                                const float* dt, const float* xref, int H, const vf* dw, int nvalid, int want_grad,
                                vf* xs, vf* aux, double* cost_sum, double* grad_sum, float* xtraj_out, int noisy) {
  const int nu = c->d->n_u;
  vf x[13], u[SDE4_MAX_NU], un[SDE4_MAX_NU];
  for (int i = 0; i < 13; ++i) {
    x[i] = vset(y0[i]);
    xs[i] = x[i];
  }
  for (int j = 0; j < nu; ++j) u[j] = vset(opt[j]);
  vf cost = vset(0.0f);
  vf c_prev = rotor_cost(cd, nu, x, u, xref, 1, 0.0f, NULL, NULL);
  for (int t = 0; t < H; ++t) {
    vf f[13], sig[13];
    rotor_drift(c, x, u, f, NULL, NULL, NULL);
    if (noisy) {
      rotor_diffusion(c, x, sig, NULL, NULL);
      for (int i = 0; i < 13; ++i) x[i] = (x[i] + f[i] * dt[t]) + sig[i] * dw[t * 13 + i]; /* sde_solver.py:301 */
    } else {
      for (int i = 0; i < 13; ++i) x[i] = x[i] + f[i] * dt[t];
    }
    {
      vf s = x[6] * x[6] + x[7] * x[7] + x[8] * x[8] + x[9] * x[9];
      vf r;
      for (int l = 0; l < VL; ++l) r[l] = 1.0f / sqrtf(s[l]);
      for (int i = 6; i < 10; ++i) x[i] *= r; /* projection_fn, sde_rotor_model.py:160-171 */
      aux[t] = r;
    }
    for (int i = 0; i < 13; ++i) xs[(t + 1) * 13 + i] = x[i];
    const int tn = t + 1 < H ? t + 1 : H - 1;
    for (int j = 0; j < nu; ++j) un[j] = vset(opt[tn * nu + j]);
    vf c_next = rotor_cost(cd, nu, x, un, xref + (t + 1) * 13, 1, 0.0f, NULL, NULL);
    cost += (0.5f * dt[t]) * (c_prev + c_next); /* nsde.py:1515 */
    c_prev = c_next;
    for (int j = 0; j < nu; ++j) u[j] = un[j];
  }
  for (int l = 0; l < nvalid; ++l) *cost_sum += cost[l];
  if (xtraj_out)
    for (int t = 0; t <= H; ++t)
      for (int i = 0; i < 13; ++i)
        for (int l = 0; l < nvalid; ++l) xtraj_out[((size_t)l * (H + 1) + t) * 13 + i] = xs[t * 13 + i][l];
  if (!want_grad) return;
  vf lam[13], gu_pend[SDE4_MAX_NU];
  for (int i = 0; i < 13; ++i) lam[i] = vset(0.0f);
  for (int j = 0; j < nu; ++j) gu_pend[j] = vset(0.0f);
  rotor_cost(cd, nu, x, u, xref + H * 13, 1, 0.5f * dt[H - 1], lam, gu_pend);
  vf xn[13];
  for (int i = 0; i < 13; ++i) xn[i] = x[i];
  for (int t = H - 1; t >= 0; --t) {
    vf gu[SDE4_MAX_NU], gx[13], ctf[13], cts[13];
    for (int i = 0; i < 13; ++i) x[i] = xs[t * 13 + i];
    for (int j = 0; j < nu; ++j) {
      u[j] = vset(opt[t * nu + j]);
      gu[j] = gu_pend[j];
      gu_pend[j] = vset(0.0f);
    }
    {
      vf dot = xn[6] * lam[6] + xn[7] * lam[7] + xn[8] * lam[8] + xn[9] * lam[9];
      for (int i = 6; i < 10; ++i) lam[i] = (lam[i] - xn[i] * dot) * aux[t];
    }
    for (int i = 0; i < 13; ++i) {
      gx[i] = lam[i];
      ctf[i] = dt[t] * lam[i];
    }
    rotor_drift(c, x, u, NULL, ctf, gx, gu);
    if (noisy) {
      for (int i = 0; i < 13; ++i) cts[i] = dw[t * 13 + i] * lam[i];
      rotor_diffusion(c, x, NULL, cts, gx);
    }
    const float wt = t > 0 ? 0.5f * (dt[t - 1] + dt[t]) : 0.5f * dt[t];
    rotor_cost(cd, nu, x, u, xref + t * 13, 1, wt, gx, gu);
    for (int j = 0; j < nu; ++j)
      for (int l = 0; l < nvalid; ++l) grad_sum[t * nu + j] += gu[j][l];
    for (int i = 0; i < 13; ++i) {
      lam[i] = gx[i];
      xn[i] = x[i];
    }
  }
}

/* sde4b200.h -- C-ABI of libsde4b200.so, the B200 (sm_100a) implementation of the
 * sde4mbrl hot path: batched Euler-Maruyama particle rollouts of the controlled
 * neural SDE and the fused cost + reverse-mode pass the APG-MPC solver needs.
 *
 * The reference (wuwushrek/sde4mbrl) is pure Python/JAX and has no FFI of its own;
 * each entry point below names the reference interface it replaces (paths relative
 * to the reference root).  This is the set of symbols an XLA custom-call /
 * jax.ffi shim (or ctypes, as in sde4mbrl_b200/_lib.py) binds -- see INTEGRATION.md.
 *
 * Conventions
 *   - plain C, POD structs, device pointers unless stated, caller owns every
 *     buffer; the library allocates nothing persistent and keeps no global
 *     mutable state except a thread-local error string;
 *   - all launchers are asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return 0 on success, a negative sde4_status otherwise; the message is
 *     available from sde4_last_error();
 *   - "samples" g = p*N + n enumerate (problem p, particle n); every per-sample
 *     array is SAMPLE-MINOR (g is the fastest index) so that a warp's accesses
 *     coalesce: xevol[H+1][x_rows][G], noise[H][n_x][G].
 *     The reference's logical [P, N, H+1, n_x] array is the strided view
 *     (stride_p=N, stride_n=1, stride_t=x_rows*G, stride_i=G) of that buffer.
 */
#ifndef SDE4B200_H_
#define SDE4B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDE4_ABI_VERSION 1
#define SDE4_MAX_NX 16
#define SDE4_MAX_NU 8
#define SDE4_MAX_NOPT 24
#define SDE4_LANES_WS 64
/* sde4_rollout_args.lanes = SDE4_LANES_TC: the tensor-core forward rollout (one thread per sample, the [32, 32] hidden
 * layers as tcgen05 3xTF32 products); picked automatically for >= 2 tiles of 128 samples per SM */
#define SDE4_LANES_TC 65

typedef enum sde4_status {
  SDE4_OK = 0,
  SDE4_ERR_INVALID = -1,      /* bad argument (null pointer, negative size ...)      */
  SDE4_ERR_UNSUPPORTED = -2,  /* model / architecture / option has no compiled kernel */
  SDE4_ERR_CUDA = -3,         /* CUDA runtime error (message has the cudaError)       */
  SDE4_ERR_WORKSPACE = -4     /* workspace too small                                   */
} sde4_status;

/* Which ControlledSDE subclass (sde4mbrl/nsde.py:134) the descriptor describes. */
typedef enum sde4_model_id {
  SDE4_MODEL_MSD = 1,       /* sde4mbrlExamples/mass_spring_damper/mass_spring_model.py:21-123 */
  SDE4_MODEL_CARTPOLE = 2,  /* sde4mbrlExamples/cartpole/cartpole_sde.py:20-106               */
  SDE4_MODEL_ROTOR = 3      /* sde4mbrlExamples/rotor_uav/sde_rotor_model.py:121-458          */
} sde4_model_id;

typedef enum sde4_act { SDE4_ACT_NONE = 0, SDE4_ACT_TANH = 1, SDE4_ACT_SWISH = 2 } sde4_act;

/* jax.random stream layout (jax_threefry_partitionable False / True). */
typedef enum sde4_rng_mode { SDE4_RNG_ORIGINAL = 0, SDE4_RNG_PARTITIONABLE = 1 } sde4_rng_mode;

/* Where the Brownian increments dw[H][n_x] of each sample come from. */
typedef enum sde4_noise_mode {
  SDE4_NOISE_NONE = 0,      /* ODE: diffusion term skipped                                    */
  SDE4_NOISE_THREEFRY = 1,  /* in-kernel threefry2x32, bit-exact with jax.random (see rng_*)  */
  SDE4_NOISE_GIVEN = 2      /* caller-supplied standard normals, noise[H][n_x][G]             */
} sde4_noise_mode;

/* model flags */
#define SDE4_MSD_CONTROL       (1 << 0)  /* params['control']            mass_spring_model.py:54 */
#define SDE4_MSD_GROUND_TRUTH  (1 << 1)  /* params['ground_truth']       :58                     */
#define SDE4_MSD_SIDE_INFO     (1 << 2)  /* params['include_side_info']  :80                     */
#define SDE4_CART_SIDE_INFO    (1 << 0)  /* params['side_info']          cartpole_sde.py:70      */
#define SDE4_ROTOR_AERO_DRAG   (1 << 0)  /* params['aero_drag_effect']   sde_rotor_model.py:322  */

/* hk.nets.MLP with exactly two hidden layers (every shipped config):
 * n_in -> h1 -> h2 -> n_out, activation after the two hidden layers only.
 * n_in == 0 means "network absent". */
typedef struct sde4_net_desc {
  int32_t n_in, h1, h2, n_out, act;
} sde4_net_desc;

/* Static + numeric description of one nSDE model.  Everything that is NOT a
 * learnable parameter lives here (it comes from the reference's `params_model`
 * dict); learnable parameters come in the packed `params` block, laid out by
 * sde4_param_layout(). */
typedef struct sde4_model_desc {
  int32_t model_id;        /* sde4_model_id                                         */
  int32_t n_x, n_u;        /* solver state dim and control dim                      */
  int32_t flags;           /* SDE4_<MODEL>_* bits                                   */
  sde4_net_desc density;   /* diffusion_density_nn.density_nn   nsde.py:384-390     */
  sde4_net_desc net_a;     /* MSD: Fres; cartpole/rotor: res_forces                 */
  sde4_net_desc net_b;     /* cartpole: control_nn; rotor: res_moments              */
  uint32_t noise_in_mask;  /* bit i: reduced_state[i] feeds the density net :353,361*/
  uint32_t noise_out_mask; /* bit i: state i is scaled by eta (others by 1)  :356,363*/
  uint32_t net_a_in_mask;  /* rotor: residual_forces.active_indx over (v_b, w) :317 */
  uint32_t net_b_in_mask;  /* rotor: residual_moments.active_indx             :348 */
  int32_t n_scaler;        /* 0: no scaler_nn; else 2*popcount(noise_out_mask) :372 */
  float aggr;              /* density_nn.aggressiveness                      :398   */
  float noise_prior[SDE4_MAX_NX];       /* noise_prior_params                        */
  float inv_state_scaling[SDE4_MAX_NX]; /* 1/state_scaling (1 if absent)             */
  float inv_red_scale;     /* rotor: 1/max(state_scaling) sde_rotor_model.py:139-144;
                              cartpole: 1/state_scaling[-1] cartpole_sde.py:41      */
  float cfloat[8];         /* MSD: mass,k,b ; cartpole: nn_out_scale[2]; rotor: gravity */
} sde4_model_desc;

/* Stage / terminal cost evaluated inside the cost+gradient kernel. */
typedef enum sde4_cost_kind {
  SDE4_COST_NONE = 0,
  SDE4_COST_ROTOR_TRACK = 1, /* one_step_cost_f, rotor_uav/sde_mpc_design.py:73-132   */
  SDE4_COST_QUADRATIC = 2    /* sum w_i (phi(x)-phi(xref))_i^2 + sum r_j (u-uref)_j^2 */
} sde4_cost_kind;

typedef enum sde4_constr_mode {
  SDE4_CONSTR_NONE = 0,
  SDE4_CONSTR_NOPROX = 1,   /* constr_cost_noprox   nsde.py:1400-1410 */
  SDE4_CONSTR_WITHPROX = 2  /* constr_cost_withprox nsde.py:1413-1419 */
} sde4_constr_mode;

typedef struct sde4_cost_desc {
  int32_t kind;              /* sde4_cost_kind                                         */
  int32_t feat;              /* QUADRATIC: 0 identity, 1 cartpole [x,xd,sin,cos,thd]   */
  uint32_t sig_mask;         /* ROTOR: bit g (0 p,1 v,2 q,3 w,4 u): '<g>err_sig' given */
  int32_t use_res_sig;       /* ROTOR: 'res_sig' = (lo, mult) given                    */
  float res_mult;            /* 'res_mult' (default 1)                                 */
  float res_sig_lo, res_sig_mult;
  float w_x[SDE4_MAX_NX];    /* ROTOR: perr(0-2) verr(3-5) qerr(7-9) werr(10-12); QUADRATIC: feature weights */
  float w_x_sig[SDE4_MAX_NX];
  float w_u[SDE4_MAX_NU];    /* 'uerr' broadcast to n_u                                */
  float w_u_sig[SDE4_MAX_NU];
  float uref[SDE4_MAX_NU];
  /* state constraints (nsde.py:75-130, 1392-1470) */
  int32_t constr_mode;       /* sde4_constr_mode                                       */
  int32_t n_slack;           /* extra opt columns (WITHPROX), else 0                   */
  int32_t slack_col[SDE4_MAX_NX]; /* state i -> slack column k, or -1 if unconstrained */
  float pen[SDE4_MAX_NX], lb[SDE4_MAX_NX], ub[SDE4_MAX_NX], slack_scale[SDE4_MAX_NX];
  float constr_weight;
  /* particle-independent control slew term of augmented_cost (sde_mpc_design.py:141-148) */
  int32_t has_slew;
  float u_slew_coeff[SDE4_MAX_NU];
} sde4_cost_desc;

/* ---- rollout: create_sampling_fn(...).multi_sampling, sde4mbrl/nsde.py:1024-1028
 *      -> sample_sde (:544-556) -> euler_maruyama, sde4mbrl/sde_solver.py:243-317.
 *      One call = P problems x N particles x H steps. */
enum sde4_solver { SDE4_SOLVER_EULER_MARUYAMA = 0, SDE4_SOLVER_STRAT_HEUN = 1, SDE4_SOLVER_STRAT_MILSTEIN = 2 };

typedef struct sde4_rollout_args {
  int32_t struct_size;          /* sizeof(sde4_rollout_args), ABI check               */
  int32_t P, N, H;
  const float* params;          /* packed learnable block (sde4_param_layout), 16B aligned */
  const float* y0;              /* [P][n_x] initial observation (= state, identity encoder) */
  const float* u;               /* controls; element (p,t,j) at u[p*u_stride_p + t*u_stride_t + j*u_stride_j] */
  int64_t u_stride_p, u_stride_t, u_stride_j;
  const float* dt;              /* [H] time steps (compute_timesteps, nsde.py:15-40)  */
  const uint32_t* key;          /* jax-layout keys uint32[2]; problem p uses key + p*key_stride_p */
  int64_t key_stride_p;         /* 0: one key shared by all problems, 2: one per problem */
  int32_t rng_mode;             /* sde4_rng_mode                                      */
  int32_t noise_mode;           /* sde4_noise_mode                                    */
  const float* noise;           /* NOISE_GIVEN: [H][n_x][G]                           */
  float* xevol;                 /* out [H+1][x_rows][G] (sde4_sim_rollout: may be NULL)  */
  int32_t x_rows;               /* >= n_x (rows beyond n_x are left untouched)        */
  int32_t block_threads;        /* 0 = library heuristic                              */
  /* particle sharding across GPUs: this call holds particles [particle_offset,
   * particle_offset+N) of particle_total per problem (keys are split(rng, particle_total)).
   * particle_total == 0 means (offset 0, total N). */
  int32_t particle_offset, particle_total;
  /* lanes per sample: 0 = library heuristic, 1 = one thread per sample, L > 1 = lane-split kernel
   * (only if compiled for this architecture, else an error).  sde4_cost_args only: SDE4_LANES_WS = the
   * warp-specialised value + gradient kernel (the two halves of a step on different warps; an error when the
   * architecture has none or the call is not value + gradient with threefry noise and no state constraints). */
  int32_t lanes;
  /* fixed-step scheme, sde4_solver.  EULER_MARUYAMA is the reference's only registered solver
   * (sde4mbrl/sde_solver.py:243-317, :321-325).  The two Stratonovich schemes exist in the reference as commented-out
   * code only (sde_solver.py:9-57 Heun, :114-177 derivative-free Milstein); they are offered here as an EXTENSION
   * of the rollout kernel (no reference to check against; unlike the live Euler-Maruyama step they scale the noise
   * by sqrt(dt), as the comments do).  The cost / gradient kernels are Euler-Maruyama only. */
  int32_t solver;
} sde4_rollout_args;

int sde4_rollout(const sde4_model_desc* model, const sde4_rollout_args* args, void* stream);

/* ---- the nSDE as an MBRL simulator: cartpole_sde_gym._my_pred_fn (sde4mbrlExamples/cartpole/cartpole_sde.py:214-223:
 *      mean / population std over the particles of the LAST state) batched over P start states, and the per-step
 *      reward / termination of mbrl-lib's ModelEnv (mbrlLibUtils/reward_functions.py:3-9, termination_functions.py:3-14)
 *      accumulated inside the rollout.  The same rollout as sde4_rollout, but nothing of the trajectory is written unless
 *      args->xevol is given: 2^20 start states x 30 steps leave 16 MB + 4 MB instead of 520 MB.
 *        last[n_x][G]        last state of every sample (required: it is also the reduction's input)
 *        racc[G]             per sample: sum_t r(u_t, obs(x_{t+1})) over the steps before (and including) its
 *                            termination; required when reward_kind != 0
 *        mean / std [P][n_x] over the N particles of each problem (NULL to skip; std is the population std, jnp.std)
 *        reward_mean[P]      mean over the particles of racc (ModelEnv.evaluate_action_sequences)
 *        reward_of_mean[P], done_of_mean[P]   reward / termination of the MEAN next state with the LAST control
 *                            (ModelEnv.step with the nSDE as the dynamics model); NULL to skip */
typedef enum sde4_reward_kind {
  SDE4_REWARD_NONE = 0,
  SDE4_REWARD_CARTPOLE_SWINGUP = 1 /* r = cos(obs[3]) - 0.1 |obs[0]| on obs = [x, xd, sin th, cos th, thd] (sic: the
                                      reference takes the cosine of cos th); done when |x| >= 120 */
} sde4_reward_kind;
typedef struct sde4_sim_out {
  int32_t reward_kind;
  float* last;
  float* racc;
  float* mean;
  float* std;
  float* reward_mean;
  float* reward_of_mean;
  int32_t* done_of_mean;
} sde4_sim_out;
int sde4_sim_rollout(const sde4_model_desc* model, const sde4_rollout_args* args, const sde4_sim_out* out, void* stream);

/* ---- cost + gradient: create_online_cost_sampling_fn(...).multi_sampling,
 *      sde4mbrl/nsde.py:1485-1527,1576-1586 and jax.value_and_grad of it as used by
 *      sde4mbrl/apg.py:82,224; stage cost = one_step_cost_f, control slew =
 *      augmented_cost (rotor_uav/sde_mpc_design.py:73-167).
 *      cost[p] = mean_n sum_t 0.5*dt_t*(c_t + c_{t+1}) (+ terminal) (+ slew(opt_p));
 *      grad[p] = d cost[p] / d opt[p]   (same strides as opt). */
/* Cross-GPU sum of [cost | grad] folded into the reduction kernel (K3), for ONE problem's particles sharded over up to 8
 * GPUs of a box (jnp.mean over particles, sde4mbrl/nsde.py:1586, when the particles span devices).  Every rank's K3 writes
 * its partial [cost[P] | grad[P][H][n_opt]] into ITS buffer of the epoch's parity; the last CTA of the grid then publishes
 * the epoch in every peer's flag array (st.release.sys over NVLink), waits until every peer has published it
 * (ld.acquire.sys), sums the world's buffers with peer loads in rank order -- the same bits on every rank -- and writes
 * `out`.  Two launches per sharded evaluation (K2, K3) and no library collective.
 * All pointers are device pointers valid in THIS process: buf[r] / flag[r] are rank r's allocations mapped here (CUDA IPC,
 * torch symmetric memory, or plain peer access within one process).  The caller alternates buf between two buffers per
 * rank by epoch parity (a rank may be one evaluation ahead of a peer, never two: it cannot pass the flag wait of epoch
 * e + 1 before every peer has finished reading epoch e), starts `epoch` at 1 and uses the same sequence on every rank;
 * flags and `counter` start at zero. */
typedef struct sde4_xgpu_reduce {
  int32_t world, rank;
  int32_t n;                /* floats of [cost | grad] = P * (1 + H * n_opt)                          */
  uint32_t epoch;
  float* buf[8];            /* rank r's partial buffer of this epoch's parity, n floats               */
  uint32_t* flag[8];        /* rank r's flag array, uint32[world]: slot s = last epoch rank s published */
  float* out;               /* [n] summed result (local memory)                                       */
  uint32_t* counter;        /* local uint32 (CTA ticket of K3)                                        */
} sde4_xgpu_reduce;

typedef struct sde4_cost_args {
  int32_t struct_size;
  int32_t P, N, H;
  const float* params;
  const float* y0;              /* [P][n_x]                                           */
  const float* opt;             /* optimisation variables, n_opt = n_u + n_slack columns */
  int64_t opt_stride_p, opt_stride_t, opt_stride_j;
  const float* dt;              /* [H]                                                */
  const float* xref;            /* reference (p,t,i) at xref[p*xref_stride_p + t*n_x + i], t in [0,H] */
  int64_t xref_stride_p;        /* 0: shared by all problems                          */
  const uint32_t* key;
  int64_t key_stride_p;
  int32_t rng_mode, noise_mode;
  const float* noise;           /* NOISE_GIVEN: [H][n_x][G]                           */
  const sde4_cost_desc* stage;  /* HOST pointer, copied at launch                     */
  const sde4_cost_desc* terminal; /* HOST pointer or NULL                             */
  float* cost;                  /* out [P]                                            */
  float* grad;                  /* out, same indexing as opt; NULL = value only       */
  float* xtraj;                 /* out [H+1][n_x+1][G] (row n_x = stage cost) or NULL */
  void* workspace;              /* >= sde4_cost_workspace_bytes(...)                   */
  size_t workspace_bytes;
  int32_t block_threads;
  /* particle sharding (see sde4_rollout_args): cost/grad are scaled by 1/particle_total so that a
   * SUM all-reduce over the shards gives the mean over all particles.  The slew term is added by
   * the shard with particle_offset == 0 only. */
  int32_t particle_offset, particle_total;
  int32_t lanes;                /* see sde4_rollout_args */
  /* stacked evaluations: problem p reads y0 / xref / key of problem (p mod data_mod) when data_mod > 0
   * (several candidate control sequences of the same problem, e.g. the APG line search) */
  int32_t data_mod;
  /* optional device int32[P]: problems with mask 0 are skipped -- their cost / grad entries are left untouched (the
   * batched APG solver evaluates line-search candidates one step size at a time and drops the problems that have
   * already accepted one; whole warps of skipped problems exit at once).  NULL: every problem is evaluated. */
  const int32_t* problem_mask;
  /* sde4_loss_grad only (ignored by sde4_cost_grad): optional device float[H][P*N], weight of the data term of
   * sample g = p*N + n at time t+1 in step_weight[t*P*N + g] -- the 0/1 indicator of the time indexes
   * `num_sample2consider` keeps (sde4mbrl/nsde.py:753-760), see sde4_rng_choice_mask.  NULL: every step counts. */
  const float* step_weight;
  /* optional HOST pointer: fold the cross-GPU sum of [cost | grad] into K3 (see sde4_xgpu_reduce).  `cost` / `grad` are
   * then ignored -- the partials go to xgpu->buf[rank] and the summed result to xgpu->out as [cost[P] | grad[P][H][n_opt]];
   * needs value + gradient.  NULL: single-device evaluation. */
  const sde4_xgpu_reduce* xgpu;
} sde4_cost_args;

/* Upper bound over the lane / noise modes the launcher may pick. */
size_t sde4_cost_workspace_bytes(const sde4_model_desc* model, int32_t P, int32_t N, int32_t H,
                                 int32_t n_slack, int32_t want_grad, int32_t have_xtraj);
int sde4_cost_grad(const sde4_model_desc* model, const sde4_cost_args* args, void* stream);

/* Host-buffer variant of sde4_cost_grad: the call a jitted `multi_sampling` + `jax.value_and_grad`
 * (sde4mbrl/nsde.py:1576-1586, apg.py:82,224) makes when its arguments live in host memory.
 * In `args`, y0 / opt / xref / key / cost / grad are HOST pointers (contiguous [P][n_x], [P][H][n_opt],
 * [P or 1][H+1][n_x], [P or 1][2]); params, dt, noise, xtraj and workspace stay device pointers.
 * The inputs are packed into `staging_pinned` (page-locked host memory), moved with ONE async copy into
 * `staging_device`, evaluated, and [cost | grad] comes back with ONE copy (or, when the pinned buffer is mapped
 * into the device's address space -- unified addressing -- stored straight into it by the reduce kernel); the call returns after the
 * stream has been synchronised.  Both staging buffers are caller-owned and at least
 * sde4_cost_grad_host_staging_bytes() long.
 * With `args->xgpu` (particles of the problem sharded over GPUs, see sde4_xgpu_reduce) `xgpu->out` is ignored: the last
 * CTA of the reduce kernel writes the cross-GPU TOTAL into the staging buffer and `cost` / `grad` receive that total. */
size_t sde4_cost_grad_host_staging_bytes(const sde4_model_desc* model, int32_t P, int32_t H, int32_t n_slack);
int sde4_cost_grad_host(const sde4_model_desc* model, const sde4_cost_args* args, void* staging_pinned,
                        void* staging_device, size_t staging_bytes, void* stream);

/* ---- RNG self-test: jax.random.{split, bits, normal} (call sites
 *      sde4mbrl/sde_solver.py:276,277,283; sde4mbrl/nsde.py:1026,1581).
 *      out[i] = element (offset+i) of random_bits(key, (total,)), uint32, bit-exact. */
int sde4_rng_bits(const uint32_t* key_host, uint64_t total, uint64_t offset, uint64_t count,
                  int32_t rng_mode, uint32_t* out, void* stream);
/* out[i] = element (offset+i) of jax.random.normal(key, (total,)), float32 */
int sde4_rng_normal(const uint32_t* key_host, uint64_t total, uint64_t offset, uint64_t count,
                    int32_t rng_mode, float* out, void* stream);
/* out[num][2] = jax.random.split(key, num); key_dev is a DEVICE uint32[2] */
int sde4_rng_split(const uint32_t* key_dev, int32_t num, int32_t rng_mode, uint32_t* out, void* stream);
/* batched forms for the key chains of the training loss (sde4mbrl/nsde.py:615,709,781): keys_dev[K][2] (device);
 * out[K][num][2] = split(keys[k], num);  out[K][n] = normal(keys[k], (n,)) */
int sde4_rng_split_many(const uint32_t* keys_dev, uint64_t K, int32_t num, int32_t rng_mode, uint32_t* out, void* stream);
int sde4_rng_normal_many(const uint32_t* keys_dev, uint64_t K, uint64_t n, int32_t rng_mode, float* out, void* stream);
/* out[t*K + k] = 1.0f if t is among jax.random.choice(keys[k], n, shape=(m,), replace=False), else 0.0f
 * (sde4mbrl/nsde.py:755; jax: permutation(key, n)[:m] = one round of a stable sort of arange(n) by
 * random_bits(split(key)[1], 32, (n,)), valid for n <= 1625 where jax needs a single round). */
int sde4_rng_choice_mask(const uint32_t* keys_dev, uint64_t K, int32_t n, int32_t m, int32_t rng_mode, float* out, void* stream);
/* out[k*n + j] = jax.random.randint(keys[k], (n,), minval, maxval)[j] (int32; the per-step index the 'random'
 * u_sampling_strategy draws, sde4mbrl/nsde.py:65 and :1308).  jax: k1, k2 = split(key); offset =
 * ((bits(k1) % span) * (2^32 % span) + bits(k2) % span) % span in uint32, span = maxval - minval (1 if not positive). */
int sde4_rng_randint_many(const uint32_t* keys_dev, uint64_t K, uint64_t n, int32_t minval, int32_t maxval, int32_t rng_mode,
                          int32_t* out, void* stream);

/* ---- training loss: create_model_loss_fn(...).loss_fn data term and jax.grad of it w.r.t. the model
 *      parameters, sde4mbrl/nsde.py:683-804 (sample_for_loss_computation), :1185-1203 (multi_sampling,
 *      loss_fn), as used by sde4mbrl/train_sde.py:248-262.
 *      Problems = minibatch elements b (P = B), each with N particles:
 *        loss[b] = mean_n sum_{t=1..H} sum_i w_i (phi(ymeas[b,t]) - phi(x[b,n,t]))_i^2
 *      with the QUADRATIC cost descriptor (w_x = 1/obs_weights^2, feat = state_transform_for_loss) in
 *      args->stage, args->xref = ymeas[b][H+1][n_x] (rows = solver steps), args->opt = the controls
 *      u[b][H][n_u], args->key = split(rng, B) (one key per element; the kernel derives
 *      split(k_b, N)[n] -> split(., 3)[0] -> split(., 2)[0] as nsde.py:1187,709 / sde_solver.py:276).
 *      grad_params (device, sde4_param_layout() floats, same layout as the parameter block) receives
 *      d (mean_b loss[b]) / d params; args->grad (may be NULL) the gradient w.r.t. the controls.
 *      args->step_weight restricts the sum over t to the steps `num_sample2consider` draws per sample.
 *      The density_loss / mu_coeff terms are i.i.d. work at the data points and live above this call
 *      (sde4mbrl_b200/density_loss.py). */
size_t sde4_loss_workspace_bytes(const sde4_model_desc* model, int32_t B, int32_t N, int32_t H);
int sde4_loss_grad(const sde4_model_desc* model, const sde4_cost_args* args, float* grad_params, void* stream);

/* ---- distance-aware regulariser of the training loss, value and parameter gradient on the device:
 *      ControlledSDE.density_loss (sde4mbrl/nsde.py:597-654), mu_coeff_nn (:657-680), their means in
 *      sample_for_loss_computation (:771-804) and weighting in create_model_loss_fn.loss_fn (:1205-1250).
 *      Per data point (b, t) of the minibatch (no rollout), z = noise_aug_state_ctrl(y[b][t], u[b][t]):
 *        den = sigmoid(density_net(z) - 7), g = d den / d z, mu = mu_coeff_nn(z);
 *        for particle n, ball sample s: delta = normal(rng_ball(b, n, t), (ns, dim))[s] * ball_radius,
 *          cond = den(z + delta) - den - g.delta - 0.5 mu |delta|^2, hinge^2 = min(cond, 0)^2
 *      with rng_ball = split(split(split(split(split(key, B)[b], N)[n], 3)[2], Hd)[t])[1].
 *      out[0] = w_grad * gradDensity + w_scvex * densitySCVEX + w_mu * lossMuCoeff, out[1] = gradDensity = mean |g|^2,
 *      out[2] = densitySCVEX = mean_{b,n} sum_t sum_s hinge^2, out[3] = muCoeff = mean mu, out[4] = lossMuCoeff =
 *      1/mu^2 | 1/mu | exp(-temp mu) of that mean (pen_mu_type).  grad_params (sde4_param_layout() floats) receives
 *      d out[0] / d density network (zero elsewhere), grad_mu the gradient w.r.t. the mu parameters (dnn: packed like a
 *      model network, w0[in][up4(h1)] b0 w1[h1][up4(h2)] b1 w2T[1][up4(h2)] b2[4]; global: 1 float). */
typedef enum sde4_mu_type { SDE4_MU_CONSTANT = 0, SDE4_MU_GLOBAL = 1, SDE4_MU_DNN = 2 } sde4_mu_type;
typedef enum sde4_pen_mu_type { SDE4_PEN_MU_QUAD_INV = 0, SDE4_PEN_MU_LIN_INV = 1, SDE4_PEN_MU_EXP_INV = 2 } sde4_pen_mu_type;
#define SDE4_REG_MAXD 24
typedef struct sde4_density_reg_args {
  uint32_t struct_size;        /* sizeof(sde4_density_reg_args)                                  */
  int32_t B, N, Hd;            /* minibatch, particles, data horizon                             */
  const float* params;         /* device, packed model block                                     */
  const float* y;              /* device float[B][Hd+1][n_x] (observations = states)             */
  const float* u;              /* device float[B][Hd][n_u]                                       */
  const uint32_t* key;         /* device uint32[2]: the rng of loss_fn                           */
  int32_t rng_mode;            /* sde4_rng_mode                                                  */
  int32_t ball_nsamples;
  float ball_radius[SDE4_REG_MAXD]; /* per density input                                         */
  int32_t mu_type;             /* sde4_mu_type                                                   */
  sde4_net_desc mu_net;        /* SDE4_MU_DNN: hidden sizes / activation of the mu network       */
  const float* mu_params;      /* device: packed mu network | float[1] (global) | NULL           */
  float mu_coeff;              /* params['density_loss']['mu_coeff']                             */
  float w_grad, w_scvex, w_mu; /* weights in out[0], pen_scvex_mult folded in                    */
  int32_t pen_mu_type;         /* sde4_pen_mu_type                                               */
  float pen_mu_temp;
  float* out;                  /* device float[5]                                                */
  float* grad_params;          /* device float[sde4_param_layout()]                              */
  float* grad_mu;              /* device: same shape as mu_params (may be NULL for constant)     */
  void* workspace;             /* device, 256-byte aligned                                       */
  size_t workspace_bytes;      /* >= sde4_density_reg_workspace_bytes(...)                       */
} sde4_density_reg_args;
size_t sde4_density_reg_workspace_bytes(const sde4_model_desc* model, int32_t B, int32_t N, int32_t Hd, int32_t ball_nsamples);
int sde4_density_reg(const sde4_model_desc* model, const sde4_density_reg_args* args, void* stream);

/* ---- batched APG-MPC: sde4mbrl/apg.py (init_apg :53-113, one_step_apg :164-274,
 *      armijo_line_search :332-379, _while_loop :451-464) for P independent problems at once.
 *      The control flow of the reference (lax.cond / bounded scans) becomes per-problem masks:
 *      a problem whose `not_done` flag is 0 keeps its state, exactly like the no-op iterations
 *      of the reference's scan.  All [H][n_opt][*] arrays are PROBLEM-MINOR (problem index fastest).
 *      One iteration = sde4_apg_candidates -> sde4_cost_grad on (maxls+1)*P problems (value only)
 *      -> sde4_apg_select -> sde4_cost_grad on 2P problems (value only) -> sde4_apg_pick -> sde4_cost_grad on P
 *      problems (value + gradient) -> sde4_apg_update; or, for latency-bound sizes, value + gradient on the 2P
 *      problems and sde4_apg_pick copying the gradient of the better of x / v (one sequential rollout less). */
typedef struct sde4_apg_params {
  int32_t P, H, n_opt;
  int32_t maxls;               /* linesearch.maxls                                            */
  int32_t reset_increase;      /* linesearch.reset_option == 'increase'                       */
  int32_t has_prox;            /* box projection on [lb, ub] (nsde.py:1435-1460)              */
  int32_t use_moment_scale;    /* optimizer_params['moment_scale'] is not None                */
  int32_t max_no_improvement_iter;
  float coef, decrease_factor, increase_factor, max_stepsize;
  float moment_scale, atol, rtol;
  float lb[SDE4_MAX_NOPT], ub[SDE4_MAX_NOPT];
} sde4_apg_params;

/* device pointers; scalars are [P]; vectors are [H][n_opt][P] */
typedef struct sde4_apg_state {
  int32_t *num_steps, *num_iter_no_improv, *not_done;
  float *momentum, *avg_momentum, *avg_linesearch, *avg_stepsize, *stepsize;
  float *opt_cost, *curr_cost, *init_cost, *grad_sqr;
  float *x_opt, *xk, *yk, *grad_yk;
} sde4_apg_state;

/* floats of the `scratch` buffer sde4_apg_init / sde4_apg_update need (|grad|^2 partials); the candidate
 * buffer `cand` is dead at both points and large enough, so callers may pass it. */
int sde4_apg_scratch_floats(const sde4_apg_params* prm);
/* init_apg after the value+gradient evaluation at x0: cost0[P], grad0 -> state (apg.py:82-112). */
int sde4_apg_init(const sde4_apg_params* prm, const sde4_apg_state* st, const float* x0, const float* cost0,
                  const float* grad0, const float* momentum_or_null, const float* stepsize_or_null,
                  float init_stepsize, float beta_init, float* scratch, void* stream);
/* candidates cand[H][n_opt][(maxls+1)*P] = yk - s_k*grad_yk, svals[(maxls+1)][P] (apg.py:179-184,312-321,358). */
/* xmask[P] (optional): cleared here, raised by sde4_apg_select for the problems whose x = prox(accepted candidate)
 * differs from the candidate (or is the never-evaluated fall-back); for the others select copies the candidate's cost
 * into xv_cost[p], and the x / v evaluation may skip x (sde4_cost_args.problem_mask = [xmask | ones]). */
int sde4_apg_candidates(const sde4_apg_params* prm, const sde4_apg_state* st, float* cand, float* svals,
                        int32_t* xmask_or_null, void* stream);
/* Armijo selection (apg.py:324-329,362-379), prox, momentum point: xv[H][n_opt][2P] = [xcurr | vcurr];
 * sel_step[P], sel_ls[P] (stepsize, number of line-search steps). */
int sde4_apg_select(const sde4_apg_params* prm, const sde4_apg_state* st, const float* cand, const float* svals,
                    const float* cand_cost, float* xv, float* sel_step, int32_t* sel_ls, int32_t* xmask_or_null,
                    float* xv_cost_or_null, void* stream);
/* ynext = where(cost_x <= cost_v, xcurr, vcurr) (apg.py:219-221): y[H][n_opt][P], pick[P].  If the x / v evaluation
 * was a value+gradient launch, xv_grad[H][n_opt][2P] is given and the gradient of the chosen point is copied into
 * y_grad[H][n_opt][P] (apg.py:224 then needs no further launch); both NULL otherwise. */
int sde4_apg_pick(const sde4_apg_params* prm, const float* xv, const float* xv_cost, float* y, float* y_cost,
                  int32_t* pick, const float* xv_grad_or_null, float* y_grad_or_null, void* stream);
/* Speculative iteration for latency-bound sizes: with Q = 2(maxls+1) + (has_prox ? maxls : 0),
 * spec[H][n_opt][Q*P] = [ x_k = prox(cand_k), k = 0..maxls | v_k = xk + momentum (x_k - xk) | (has_prox) cand_k, k < maxls ]
 * is evaluated by ONE value+gradient sde4_cost_grad launch over Q*P problems (data_mod = P); sde4_apg_select then
 * takes cand_cost = the cost block of the un-projected candidates (the x block when there is no prox), and
 * sde4_apg_spec_gather copies cost and gradient of the accepted step's x / v into xv_cost[2P] / xv_grad[H][n_opt][2P]
 * for sde4_apg_pick -- one sequential rollout per iteration instead of two (apg.py:164-224 unchanged in effect). */
int sde4_apg_spec_points(const sde4_apg_params* prm, const sde4_apg_state* st, const float* cand, float* spec, void* stream);
int sde4_apg_spec_gather(const sde4_apg_params* prm, const int32_t* sel_ls, const float* spec_cost, const float* spec_grad,
                         float* xv_cost, float* xv_grad, void* stream);
/* progressive line search: after the costs of candidate k are known, active[p] &= wolfe_cond_violated(k)
 * (apg.py:324-329, 362-379); problems that accepted candidate k are skipped by the next candidate's launch.
 * k == 0 first sets active[p] = not_done[p]. */
int sde4_apg_ls_step(const sde4_apg_params* prm, const sde4_apg_state* st, const float* svals, const float* cand_cost,
                     int32_t k, int32_t* active, void* stream);
/* state update with grad(ynext) (apg.py:224-273); problems with not_done == 0 are left untouched */
int sde4_apg_update(const sde4_apg_params* prm, const sde4_apg_state* st, const float* xv, const float* y,
                    const float* y_cost, const float* y_grad, const int32_t* pick, const float* sel_step,
                    const int32_t* sel_ls, float* scratch, void* stream);

/* ---- layout of the packed parameter block -------------------------------------
 * Returns the number of floats of the block for `model` (<0 on error).  The
 * order is: model scalars, scaler vector, density net, net_a, net_b; every
 * tensor starts on a 4-float boundary; a network is
 *   w0[in][h1_4] b0[h1_4] w1[h1][h2_4] b1[h2_4] w2T[out][h2_4] b2[out_4]
 * (w0, w1 row-major [in][out] as haiku stores them, `out` padded up to a multiple of 4 with
 * zeros; the last layer is stored transposed, one contiguous row per output).  If `offsets` is non-NULL it receives, in floats:
 *   [0] scalars  [1] scaler  [2] density  [3] net_a  [4] net_b  [5] total.
 * Rotor scalars (16): mass,Ixx,Iyy,Izz,Lmx,Lmy,CoeffDrag, c3,c2,c1,c0 (motor
 * polynomial, highest power first), aero_kdx,kdy,kdz,kh, pad. */
int sde4_param_layout(const sde4_model_desc* model, int32_t* offsets /*[6] or NULL*/);

/* 1 if a compiled specialisation exists for this architecture, else 0. */
int sde4_model_supported(const sde4_model_desc* model);
/* Human-readable list of compiled architectures (static string). */
const char* sde4_supported_archs(void);

const char* sde4_last_error(void);
int sde4_version(void);
/* sizeof() of the four POD structs, in the order model_desc, cost_desc, rollout_args, cost_args
 * (lets a foreign-language binding verify its struct layout at load time). */
void sde4_struct_sizes(int32_t out[4]);
void sde4_apg_struct_sizes(int32_t out[2]); /* sizeof sde4_apg_params, sde4_apg_state */
void sde4_aux_struct_sizes(int32_t out[3]); /* sizeof sde4_sim_out, sde4_xgpu_reduce, sde4_density_reg_args */

/* ---- FP32 FMA micro-benchmark (roofline denominator) ---------------
 * Runs `iters` dependent-chain-free FFMA (variant 0) / packed fma.rn.f32x2
 * (variant 1) per thread on `blocks` x `threads`; writes nothing but a checksum.
 * Returns the flop count through *flops; time it with events on `stream`. */
int sde4_fma_peak(int32_t variant, int32_t blocks, int32_t threads, int32_t iters, float* sink,
                  double* flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDE4B200_H_ */

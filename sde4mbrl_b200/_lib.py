"""ctypes binding of ``libsde4b200.so`` (the C-ABI declared in ``include/sde4b200.h``).

This is the binding used in this image (no jax / XLA FFI headers available); the
XLA-FFI stub a reference maintainer would add is shown in ``INTEGRATION.md``.
The product path FAILS LOUDLY when the CUDA library is missing -- there is no
CPU fallback anywhere in this package.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsde4b200.so")

MAX_NX, MAX_NU, MAX_NOPT = 16, 8, 24

MODEL_MSD, MODEL_CARTPOLE, MODEL_ROTOR = 1, 2, 3
ACT_NONE, ACT_TANH, ACT_SWISH = 0, 1, 2
RNG_ORIGINAL, RNG_PARTITIONABLE = 0, 1
NOISE_NONE, NOISE_THREEFRY, NOISE_GIVEN = 0, 1, 2
LANES_WS = 64  # sde4_cost_args.lanes: the warp-specialised value + gradient kernel (SDE4_LANES_WS)
LANES_TC = 65  # sde4_rollout_args.lanes: the tensor-core forward rollout (SDE4_LANES_TC)
MSD_CONTROL, MSD_GROUND_TRUTH, MSD_SIDE_INFO = 1, 2, 4
CART_SIDE_INFO = 1
ROTOR_AERO_DRAG = 1
COST_NONE, COST_ROTOR_TRACK, COST_QUADRATIC = 0, 1, 2
CONSTR_NONE, CONSTR_NOPROX, CONSTR_WITHPROX = 0, 1, 2


class NetDesc(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("h1", C.c_int32), ("h2", C.c_int32), ("n_out", C.c_int32), ("act", C.c_int32)]


class ModelDesc(C.Structure):
    _fields_ = [("model_id", C.c_int32), ("n_x", C.c_int32), ("n_u", C.c_int32), ("flags", C.c_int32),
                ("density", NetDesc), ("net_a", NetDesc), ("net_b", NetDesc),
                ("noise_in_mask", C.c_uint32), ("noise_out_mask", C.c_uint32),
                ("net_a_in_mask", C.c_uint32), ("net_b_in_mask", C.c_uint32),
                ("n_scaler", C.c_int32), ("aggr", C.c_float),
                ("noise_prior", C.c_float * MAX_NX), ("inv_state_scaling", C.c_float * MAX_NX),
                ("inv_red_scale", C.c_float), ("cfloat", C.c_float * 8)]


class CostDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("feat", C.c_int32), ("sig_mask", C.c_uint32), ("use_res_sig", C.c_int32),
                ("res_mult", C.c_float), ("res_sig_lo", C.c_float), ("res_sig_mult", C.c_float),
                ("w_x", C.c_float * MAX_NX), ("w_x_sig", C.c_float * MAX_NX),
                ("w_u", C.c_float * MAX_NU), ("w_u_sig", C.c_float * MAX_NU), ("uref", C.c_float * MAX_NU),
                ("constr_mode", C.c_int32), ("n_slack", C.c_int32), ("slack_col", C.c_int32 * MAX_NX),
                ("pen", C.c_float * MAX_NX), ("lb", C.c_float * MAX_NX), ("ub", C.c_float * MAX_NX),
                ("slack_scale", C.c_float * MAX_NX), ("constr_weight", C.c_float),
                ("has_slew", C.c_int32), ("u_slew_coeff", C.c_float * MAX_NU)]

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        for i in range(MAX_NX):
            self.slack_col[i] = -1
        self.res_mult = 1.0
        self.constr_weight = 1.0


class RolloutArgs(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("P", C.c_int32), ("N", C.c_int32), ("H", C.c_int32),
                ("params", C.c_void_p), ("y0", C.c_void_p), ("u", C.c_void_p),
                ("u_stride_p", C.c_int64), ("u_stride_t", C.c_int64), ("u_stride_j", C.c_int64),
                ("dt", C.c_void_p), ("key", C.c_void_p), ("key_stride_p", C.c_int64),
                ("rng_mode", C.c_int32), ("noise_mode", C.c_int32), ("noise", C.c_void_p),
                ("xevol", C.c_void_p), ("x_rows", C.c_int32), ("block_threads", C.c_int32),
                ("particle_offset", C.c_int32), ("particle_total", C.c_int32), ("lanes", C.c_int32),
                ("solver", C.c_int32)]


SOLVER_EULER_MARUYAMA, SOLVER_STRAT_HEUN, SOLVER_STRAT_MILSTEIN = 0, 1, 2
SOLVERS = {"euler_maruyama": SOLVER_EULER_MARUYAMA, "stratonovich_heun": SOLVER_STRAT_HEUN,
           "stratonovich_milstein": SOLVER_STRAT_MILSTEIN}


class CostArgs(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("P", C.c_int32), ("N", C.c_int32), ("H", C.c_int32),
                ("params", C.c_void_p), ("y0", C.c_void_p), ("opt", C.c_void_p),
                ("opt_stride_p", C.c_int64), ("opt_stride_t", C.c_int64), ("opt_stride_j", C.c_int64),
                ("dt", C.c_void_p), ("xref", C.c_void_p), ("xref_stride_p", C.c_int64),
                ("key", C.c_void_p), ("key_stride_p", C.c_int64),
                ("rng_mode", C.c_int32), ("noise_mode", C.c_int32), ("noise", C.c_void_p),
                ("stage", C.POINTER(CostDesc)), ("terminal", C.POINTER(CostDesc)),
                ("cost", C.c_void_p), ("grad", C.c_void_p), ("xtraj", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("block_threads", C.c_int32),
                ("particle_offset", C.c_int32), ("particle_total", C.c_int32), ("lanes", C.c_int32),
                ("data_mod", C.c_int32), ("problem_mask", C.c_void_p), ("step_weight", C.c_void_p),
                ("xgpu", C.c_void_p)]


REWARD_NONE, REWARD_CARTPOLE_SWINGUP = 0, 1


class SimOut(C.Structure):  # sde4_sim_out
    _fields_ = [("reward_kind", C.c_int32), ("last", C.c_void_p), ("racc", C.c_void_p), ("mean", C.c_void_p),
                ("std", C.c_void_p), ("reward_mean", C.c_void_p), ("reward_of_mean", C.c_void_p),
                ("done_of_mean", C.c_void_p)]


MU_CONSTANT, MU_GLOBAL, MU_DNN = 0, 1, 2
PEN_MU = {"quad_inv": 0, "lin_inv": 1, "exp_inv": 2}
REG_MAXD = 24


class DensityRegArgs(C.Structure):  # sde4_density_reg_args
    _fields_ = [("struct_size", C.c_uint32), ("B", C.c_int32), ("N", C.c_int32), ("Hd", C.c_int32),
                ("params", C.c_void_p), ("y", C.c_void_p), ("u", C.c_void_p), ("key", C.c_void_p),
                ("rng_mode", C.c_int32), ("ball_nsamples", C.c_int32), ("ball_radius", C.c_float * REG_MAXD),
                ("mu_type", C.c_int32), ("mu_net", NetDesc), ("mu_params", C.c_void_p), ("mu_coeff", C.c_float),
                ("w_grad", C.c_float), ("w_scvex", C.c_float), ("w_mu", C.c_float), ("pen_mu_type", C.c_int32),
                ("pen_mu_temp", C.c_float), ("out", C.c_void_p), ("grad_params", C.c_void_p), ("grad_mu", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class XgpuReduce(C.Structure):  # sde4_xgpu_reduce
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("n", C.c_int32), ("epoch", C.c_uint32),
                ("buf", C.c_void_p * 8), ("flag", C.c_void_p * 8), ("out", C.c_void_p), ("counter", C.c_void_p)]


class ApgParams(C.Structure):
    _fields_ = [("P", C.c_int32), ("H", C.c_int32), ("n_opt", C.c_int32), ("maxls", C.c_int32),
                ("reset_increase", C.c_int32), ("has_prox", C.c_int32), ("use_moment_scale", C.c_int32),
                ("max_no_improvement_iter", C.c_int32),
                ("coef", C.c_float), ("decrease_factor", C.c_float), ("increase_factor", C.c_float),
                ("max_stepsize", C.c_float), ("moment_scale", C.c_float), ("atol", C.c_float), ("rtol", C.c_float),
                ("lb", C.c_float * MAX_NOPT), ("ub", C.c_float * MAX_NOPT)]


APG_STATE_INT = ("num_steps", "num_iter_no_improv", "not_done")
APG_STATE_F32 = ("momentum", "avg_momentum", "avg_linesearch", "avg_stepsize", "stepsize", "opt_cost", "curr_cost",
                 "init_cost", "grad_sqr")
APG_STATE_VEC = ("x_opt", "xk", "yk", "grad_yk")


class ApgState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in APG_STATE_INT + APG_STATE_F32 + APG_STATE_VEC]


EXPORTED_SYMBOLS = (
    "sde4_rollout", "sde4_cost_workspace_bytes", "sde4_cost_grad", "sde4_loss_workspace_bytes", "sde4_loss_grad", "sde4_rng_bits", "sde4_rng_normal",
    "sde4_rng_split", "sde4_param_layout", "sde4_model_supported", "sde4_supported_archs",
    "sde4_last_error", "sde4_version", "sde4_struct_sizes", "sde4_apg_struct_sizes", "sde4_fma_peak",
    "sde4_apg_init", "sde4_apg_candidates", "sde4_apg_select", "sde4_apg_pick", "sde4_apg_update",
    "sde4_apg_scratch_floats", "sde4_cost_grad_host", "sde4_cost_grad_host_staging_bytes",
    "sde4_sim_rollout", "sde4_aux_struct_sizes", "sde4_density_reg", "sde4_density_reg_workspace_bytes", "sde4_rng_split_many", "sde4_rng_normal_many", "sde4_rng_choice_mask", "sde4_rng_randint_many", "sde4_apg_ls_step", "sde4_apg_spec_points", "sde4_apg_spec_gather",
)

_lib = None


class Sde4Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Sde4Error(
            "libsde4b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C sde4mbrl_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.sde4_rollout.argtypes = [C.POINTER(ModelDesc), C.POINTER(RolloutArgs), C.c_void_p]
    lib.sde4_rollout.restype = C.c_int
    lib.sde4_sim_rollout.argtypes = [C.POINTER(ModelDesc), C.POINTER(RolloutArgs), C.POINTER(SimOut), C.c_void_p]
    lib.sde4_sim_rollout.restype = C.c_int
    lib.sde4_density_reg.argtypes = [C.POINTER(ModelDesc), C.POINTER(DensityRegArgs), C.c_void_p]
    lib.sde4_density_reg.restype = C.c_int
    lib.sde4_density_reg_workspace_bytes.argtypes = [C.POINTER(ModelDesc), C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.sde4_density_reg_workspace_bytes.restype = C.c_size_t
    lib.sde4_cost_workspace_bytes.argtypes = [C.POINTER(ModelDesc), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                              C.c_int32, C.c_int32]
    lib.sde4_cost_workspace_bytes.restype = C.c_size_t
    lib.sde4_cost_grad.argtypes = [C.POINTER(ModelDesc), C.POINTER(CostArgs), C.c_void_p]
    lib.sde4_cost_grad.restype = C.c_int
    lib.sde4_cost_grad_host.argtypes = [C.POINTER(ModelDesc), C.POINTER(CostArgs), C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_void_p]
    lib.sde4_cost_grad_host.restype = C.c_int
    lib.sde4_cost_grad_host_staging_bytes.argtypes = [C.POINTER(ModelDesc), C.c_int32, C.c_int32, C.c_int32]
    lib.sde4_cost_grad_host_staging_bytes.restype = C.c_size_t
    lib.sde4_loss_workspace_bytes.argtypes = [C.POINTER(ModelDesc), C.c_int32, C.c_int32, C.c_int32]
    lib.sde4_loss_workspace_bytes.restype = C.c_size_t
    lib.sde4_loss_grad.argtypes = [C.POINTER(ModelDesc), C.POINTER(CostArgs), C.c_void_p, C.c_void_p]
    lib.sde4_loss_grad.restype = C.c_int
    for name in ("sde4_rng_bits", "sde4_rng_normal"):
        fn = getattr(lib, name)
        fn.argtypes = [C.POINTER(C.c_uint32), C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.c_void_p, C.c_void_p]
        fn.restype = C.c_int
    lib.sde4_rng_split.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.sde4_rng_split.restype = C.c_int
    lib.sde4_param_layout.argtypes = [C.POINTER(ModelDesc), C.POINTER(C.c_int32)]
    lib.sde4_param_layout.restype = C.c_int
    lib.sde4_model_supported.argtypes = [C.POINTER(ModelDesc)]
    lib.sde4_model_supported.restype = C.c_int
    lib.sde4_supported_archs.restype = C.c_char_p
    lib.sde4_last_error.restype = C.c_char_p
    lib.sde4_version.restype = C.c_int
    lib.sde4_rng_split_many.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.sde4_rng_split_many.restype = C.c_int
    lib.sde4_rng_normal_many.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32, C.c_void_p, C.c_void_p]
    lib.sde4_rng_normal_many.restype = C.c_int
    lib.sde4_rng_choice_mask.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.sde4_rng_choice_mask.restype = C.c_int
    lib.sde4_rng_randint_many.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.sde4_rng_randint_many.restype = C.c_int
    lib.sde4_fma_peak.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_double),
                                  C.c_void_p]
    lib.sde4_fma_peak.restype = C.c_int
    for name in ("sde4_apg_init", "sde4_apg_candidates", "sde4_apg_select", "sde4_apg_pick", "sde4_apg_update"):
        getattr(lib, name).restype = C.c_int
    lib.sde4_apg_init.argtypes = [C.POINTER(ApgParams), C.POINTER(ApgState), C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    lib.sde4_apg_ls_step.argtypes = [C.POINTER(ApgParams), C.POINTER(ApgState), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                     C.c_void_p]
    lib.sde4_apg_ls_step.restype = C.c_int
    lib.sde4_apg_spec_points.argtypes = [C.POINTER(ApgParams), C.POINTER(ApgState), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sde4_apg_spec_points.restype = C.c_int
    lib.sde4_apg_spec_gather.argtypes = [C.POINTER(ApgParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]
    lib.sde4_apg_spec_gather.restype = C.c_int
    lib.sde4_apg_scratch_floats.argtypes = [C.POINTER(ApgParams)]
    lib.sde4_apg_scratch_floats.restype = C.c_int
    lib.sde4_apg_candidates.argtypes = [C.POINTER(ApgParams), C.POINTER(ApgState), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
    lib.sde4_apg_select.argtypes = [C.POINTER(ApgParams), C.POINTER(ApgState)] + [C.c_void_p] * 9
    lib.sde4_apg_pick.argtypes = [C.POINTER(ApgParams)] + [C.c_void_p] * 8
    lib.sde4_apg_update.argtypes = [C.POINTER(ApgParams), C.POINTER(ApgState)] + [C.c_void_p] * 9
    asz = (C.c_int32 * 2)()
    lib.sde4_apg_struct_sizes(asz)
    if list(asz) != [C.sizeof(ApgParams), C.sizeof(ApgState)]:
        raise Sde4Error("ABI mismatch (apg structs): %s" % (list(asz),))
    sizes = (C.c_int32 * 4)()
    lib.sde4_struct_sizes(sizes)
    mine = [C.sizeof(ModelDesc), C.sizeof(CostDesc), C.sizeof(RolloutArgs), C.sizeof(CostArgs)]
    if list(sizes) != mine:
        raise Sde4Error("ABI mismatch between _lib.py and libsde4b200.so: %s vs %s" % (mine, list(sizes)))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise Sde4Error("libsde4b200: error %d: %s" % (rc, load().sde4_last_error().decode()))


def last_error():
    return load().sde4_last_error().decode()

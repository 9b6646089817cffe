"""Callers of the hot path for the multirotor: model / MPC loaders and the
receding-horizon MPC step, with the reference's names, arguments and return values.

  load_predictor_function   sde4mbrlExamples/rotor_uav/sde_rotor_model.py:551-624
  load_mpc_solver           sde4mbrlExamples/rotor_uav/sde_rotor_model.py:627-692
  parse_trajectory, ned2enu, extract_targets, reset_apg, create_full_cost_function, apg_mpc,
  load_mpc_from_cfgfile     sde4mbrlExamples/rotor_uav/sde_mpc_design.py:26-432
  frame conversions         sde4mbrlExamples/rotor_uav/utils.py:13-63,143-221

Everything here is host logic on small arrays (numpy); rollouts, costs and gradients go through
``nsde.create_sampling_fn`` / ``nsde.create_online_cost_sampling_fn`` (CUDA).  Stage and terminal costs are
``costs.StageCost`` descriptors evaluated inside the kernels instead of Python lambdas; the learned value /
policy networks of the reference were never used (``contain_valuepol = False``, sde_mpc_design.py:284-285).
"""
import copy
import os
import pickle

import numpy as np
import torch

from . import costs as _costs
from . import random as _random
from ._lib import Sde4Error
from .apg import apg, init_apg, init_opt_state
from .engine import is_host
from .io import update_params
from .nsde import compute_timesteps, create_online_cost_sampling_fn, create_sampling_fn
from .sde_models import SDERotorModel

state_names = ["x", "y", "z", "vx", "vy", "vz", "qw", "qx", "qy", "qz", "wx", "wy", "wz"]
state_name2axis = {name: i for i, name in enumerate(state_names)}


# ---------------------------------------------------------------------------------------------
# frames (utils.py): NED <-> ENU is the half turn about (1,1,0)/sqrt(2), FRD <-> FLU the half turn about x
# ---------------------------------------------------------------------------------------------
_Q_ENU_NED = np.array([0.0, np.sqrt(0.5), np.sqrt(0.5), 0.0])
_Q_FLU_FRD = np.array([0.0, 1.0, 0.0, 0.0])


def quatmult(a, b):
    """Hamilton product, [w, x, y, z] (utils.py:13-28)."""
    aw, av = a[0], np.asarray(a[1:4])
    bw, bv = b[0], np.asarray(b[1:4])
    return np.concatenate(([aw * bw - av @ bv], aw * bv + bw * av + np.cross(av, bv)))


def quatinv(a):
    return np.array([a[0], -a[1], -a[2], -a[3]])


def quat_rotatevector(q, v):
    """q (0, v) q^-1 (utils.py:42-52)."""
    return quatmult(quatmult(q, np.concatenate(([0.0], v))), quatinv(q))[1:]


def quat_from_euler(roll, pitch, yaw):
    """utils.py:65-84."""
    cy, sy = np.cos(yaw * 0.5), np.sin(yaw * 0.5)
    cr, sr = np.cos(roll * 0.5), np.sin(roll * 0.5)
    cp, sp = np.cos(pitch * 0.5), np.sin(pitch * 0.5)
    return np.array([cy * cr * cp + sy * sr * sp, cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp,
                     sy * cr * cp - cy * sr * sp])


def ned2enu(x):
    """Full state NED/FRD -> ENU/FLU (sde_mpc_design.py:37-46); entries beyond 13 pass through.
    The map is an involution, so it is also enu2ned."""
    x = np.asarray(x.detach().cpu() if isinstance(x, torch.Tensor) else x, dtype=np.float64)
    q = quatmult(quatmult(_Q_ENU_NED, x[6:10]), quatinv(_Q_FLU_FRD))
    out = np.concatenate((quat_rotatevector(_Q_ENU_NED, x[0:3]), quat_rotatevector(_Q_ENU_NED, x[3:6]), q,
                          quat_rotatevector(_Q_FLU_FRD, x[10:13]), x[13:]))
    return out.astype(np.float32)


enu2ned = ned2enu


# ---------------------------------------------------------------------------------------------
# reference trajectories
# ---------------------------------------------------------------------------------------------
def load_trajectory(filename):
    """csv -> {column: array} (utils.py:240-263)."""
    import pandas as pd
    df = pd.read_csv(os.path.expanduser(filename))
    return {k: np.asarray(v) for k, v in df.to_dict("list").items()}


def parse_trajectory(_traj):
    """-> (t[T], states[T, 13]) in ``state_names`` order (sde_mpc_design.py:26-35)."""
    return np.asarray(_traj["t"], np.float64), np.stack([np.asarray(_traj[s], np.float64) for s in state_names], axis=1)


def extract_targets(curr_t, _time_evol, _state_evol, time_steps):
    """Piecewise-linear look-up (extrapolating at both ends) of the reference at ``curr_t`` and at
    ``curr_t + time_steps`` (sde_mpc_design.py:48-70).  Returns ``(x(curr_t), [x(curr_t); x(curr_t + time_steps)])``."""
    t = np.asarray(_time_evol, np.float64)
    xs = np.asarray(_state_evol, np.float64)

    def look(tq):
        tq = np.atleast_1d(np.asarray(tq, np.float64))
        i = np.clip(np.searchsorted(t, tq, side="left") - 1, 0, t.shape[0] - 2)
        w = (tq - t[i]) / (t[i + 1] - t[i])
        return xs[i] + (xs[i + 1] - xs[i]) * w[:, None]

    curr = look(curr_t)[0]
    nxt = look(curr_t + np.asarray(time_steps, np.float64))
    return curr.astype(np.float32), np.concatenate((curr[None], nxt)).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# loaders
# ---------------------------------------------------------------------------------------------
def load_yaml(path):
    import yaml
    with open(os.path.expanduser(path)) as f:
        return yaml.load(f.read(), yaml.SafeLoader)


def _load_pickle(path):
    with open(os.path.expanduser(path), "rb") as f:
        d = pickle.load(f)
    return d["nominal"], d["sde"]


def _strip_learned(params_model, drift_too):
    params_model.pop("diffusion_density_nn", None)
    if drift_too:
        params_model.pop("residual_forces", None)
        params_model.pop("residual_moments", None)
        params_model.pop("aero_drag_effect", None)


def load_predictor_function(learned_params_dir, prior_dist=False, nonoise=False, modified_params={},
                            return_control=False, return_time_steps=False):
    """sde_rotor_model.py:551-624.  ``fn(x, u, rng) -> yevol[N, H+1, n_y]`` (and ``uevol`` when
    ``return_control``); with ``return_time_steps`` also the time grid ``[H+1]`` starting at 0."""
    _model_params, _sde_learned = _load_pickle(learned_params_dir)
    params_model = update_params(_model_params, modified_params)
    if params_model.get("control_augmented_state", False):
        raise Sde4Error("control_augmented_state has no compiled kernel")
    if prior_dist:
        _strip_learned(params_model, drift_too=True)
    if nonoise or prior_dist:
        params_model["noise_prior_params"] = [0] * len(params_model["noise_prior_params"])
    time_steps = compute_timesteps(params_model)
    time_evol = np.concatenate(([0.0], np.cumsum(time_steps)))
    _prior_params, m_sampling = create_sampling_fn(params_model, sde_constr=SDERotorModel)
    if prior_dist:
        _sde_learned = _prior_params

    def res_fn(*x):
        out = m_sampling(_sde_learned, *x)
        return out[1:] if return_control else out[1]

    return (res_fn, time_evol) if return_time_steps else res_fn


def load_mpc_solver(mpc_config_dir, modified_params={}, nominal_model=False):
    """sde_rotor_model.py:627-692.  ``mpc_config_dir`` is the yaml path (or an already loaded dict).
    Returns ``((sde_learned, mpc_params), multi_cost_sampling, vmapped_prox, construct_opt_params)``."""
    mpc_params = load_yaml(mpc_config_dir) if isinstance(mpc_config_dir, (str, os.PathLike)) else mpc_config_dir
    _model_params, _sde_learned = _load_pickle(mpc_params["learned_model_params"])
    params_model = update_params(_model_params, modified_params)
    if mpc_params.get("use_rate_control", False):
        raise Sde4Error("use_rate_control (control_augmented_state) has no compiled kernel")
    if "model_update" in mpc_params:
        params_model = update_params(params_model, mpc_params["model_update"])
    sys_id = mpc_params.get("use_sysId_model", False)
    if sys_id:
        print("Using the system identification model")
        _strip_learned(params_model, drift_too=True)
        params_model["noise_prior_params"] = [0] * len(params_model["noise_prior_params"])
    if nominal_model:
        _strip_learned(params_model, drift_too=False)
        params_model["noise_prior_params"] = [0] * len(params_model["noise_prior_params"])
    _prior, multi_cost, vmapped_prox, _, construct_opt_params = create_online_cost_sampling_fn(
        params_model, mpc_params, sde_constr=SDERotorModel)
    if sys_id:
        _sde_learned = _prior
    mpc_params["model"] = params_model
    return (_sde_learned, mpc_params), multi_cost, vmapped_prox, construct_opt_params


# ---------------------------------------------------------------------------------------------
# one receding-horizon step
# ---------------------------------------------------------------------------------------------
def reset_apg(x, apg_mpc_params, _constructor_uslack, rng=None, uref=None, _pol_fn=None, convert_to_enu=True,
              sample_sde=None):
    """sde_mpc_design.py:169-196 with ``_pol_fn = None`` (the only case the reference exercises)."""
    if _pol_fn is not None:
        raise Sde4Error("policy warm starts were never enabled in the reference (contain_valuepol = False)")
    opt_init = _constructor_uslack(u=uref, x=x)
    return init_opt_state(opt_init, apg_mpc_params)


def create_full_cost_function(xt, rng, target_x, multi_cost_sampling, cfg_dict, _sde_learned, _terminal_cost=None,
                              extra_dyn_args=None, convert_x=False, mod_cost_params={}):
    """sde_mpc_design.py:198-219.  Returns ``(cost_xt, red_cost_f)``: ``cost_xt(u_params)`` is the mean stage cost
    over the particles plus the control-slew term of ``augmented_cost`` (:134-167, folded into the reduction
    kernel); ``red_cost_f`` is the stage-cost descriptor."""
    if convert_x:
        xt = ned2enu(xt)
    cfg_dict["cost_params"].update(mod_cost_params)
    n_u = cfg_dict["model"]["n_u"]
    red_cost_f = _costs.with_slew(_costs.one_step_cost_f(cfg_dict["cost_params"], n_u), cfg_dict["cost_params"], n_u)

    def cost_xt(u_params):
        return multi_cost_sampling(_sde_learned, xt, u_params, rng, red_cost_f, _terminal_cost=_terminal_cost,
                                   extra_dyn_args=extra_dyn_args, extra_cost_args=target_x,
                                   extra_cost_terminal_args=target_x[-1])[0]

    return cost_xt, red_cost_f


# The APG solve of ``apg_mpc`` runs on the device-side solver (apg_batched.BatchedAPG with P = 1: line search, prox,
# momentum and stop test in CUDA kernels, one graph replay per iteration, stop flag polled every few iterations) when
# the pieces are the ones this module hands out; set to False to force the host ``apg.py`` loop.
USE_DEVICE_APG = True
_POLL_EVERY = 4


def _device_solver(multi_cost_sampling, cfg_dict, _sde_learned, stage, term, box):
    """The cached ``apg_batched.BatchedAPG`` of this cost (``nsde`` ``multi_sampling.device_solver``)."""
    return multi_cost_sampling.device_solver(_sde_learned, stage, term, cfg_dict["apg_mpc"], box)


def apg_mpc(xt, rng, past_solution, target_x, multi_cost_sampling, proximal_fn, cfg_dict, _sde_learned,
            _terminal_cost=None, extra_dyn_args=None, convert_to_enu=True, mod_cost_params={}):
    """sde_mpc_design.py:221-262.  Returns ``(uopt[H, n_u], new_solution, rng_next, vehicle_states)``."""
    host_rng = is_host(rng)
    ks = _random.split(rng)  # device kernel (bit-exact jax.random.split)
    if host_rng:  # keep host keys on the host so that the cost calls stay on the host-buffer path
        ks = ks.cpu().numpy().view(np.uint32)
    rng_next, rng = ks[0], ks[1]
    if convert_to_enu:
        xt = ned2enu(xt)
    cost_xt, red_cost_f = create_full_cost_function(xt, rng, target_x, multi_cost_sampling, cfg_dict, _sde_learned,
                                                    _terminal_cost, extra_dyn_args, convert_x=False,
                                                    mod_cost_params=mod_cost_params)
    box = getattr(proximal_fn, "box", None)
    on_device = (USE_DEVICE_APG and extra_dyn_args is None and "linesearch" in cfg_dict["apg_mpc"]
                 and hasattr(multi_cost_sampling, "final_stage") and (proximal_fn is None or box is not None))
    if on_device:
        solver = _device_solver(multi_cost_sampling, cfg_dict, _sde_learned, multi_cost_sampling.final_stage(red_cost_f),
                                _terminal_cost, box)
        solver.init(past_solution.x_opt, np.asarray(xt, np.float32)[None], target_x, rng,
                    momentum=np.float32(past_solution.momentum), stepsize=np.float32(past_solution.avg_stepsize))
        solver.run(check_every=_POLL_EVERY)
        st = solver.state_of(0)
        new_opt_state = st._replace(**{k: getattr(st, k).clone() for k in ("x_opt", "xk", "yk", "grad_yk")})
        if not isinstance(past_solution.x_opt, torch.Tensor) or not past_solution.x_opt.is_cuda:
            new_opt_state = new_opt_state._replace(**{k: getattr(new_opt_state, k).cpu() for k in ("x_opt", "xk", "yk", "grad_yk")})
    else:
        opt_state = init_apg(past_solution.x_opt, cost_xt, cfg_dict["apg_mpc"], momentum=past_solution.momentum,
                             stepsize=past_solution.avg_stepsize)
        new_opt_state = apg(opt_state, cost_xt, cfg_dict["apg_mpc"], proximal_fn=proximal_fn)
    with torch.no_grad():
        _, vehicle_states = multi_cost_sampling(_sde_learned, xt, new_opt_state.x_opt, rng, red_cost_f,
                                                _terminal_cost=_terminal_cost, extra_dyn_args=extra_dyn_args,
                                                extra_cost_args=target_x, extra_cost_terminal_args=target_x[-1])
    uopt = new_opt_state.x_opt[:, :cfg_dict["model"]["n_u"]]
    if convert_to_enu:
        return uopt, new_opt_state, rng_next, vehicle_states[0, :, :len(xt)]
    return uopt, new_opt_state, rng_next, vehicle_states


def load_mpc_from_cfgfile(path_config, convert_to_enu=True, nominal_model=False, one_step_params={},
                          return_full_cost=False):
    """sde_mpc_design.py:264-424.  Returns ``(cfg_dict, (m_reset, m_mpc), get_state_at_t, one_step_prediction)``."""
    (_sde_learned, cfg_dict), multi_cost_sampling, vmapped_prox, construct_opt_params = \
        load_mpc_solver(path_config, modified_params={}, nominal_model=nominal_model)
    if cfg_dict["model"].get("control_augmented_state", False):
        raise Sde4Error("control_augmented_state has no compiled kernel")
    cfg_dict["cost_params"].pop("urate_err", None)
    proximal_fn = None
    if vmapped_prox is not None:
        proximal_fn = lambda x, stepsize=None: vmapped_prox(x)
        proximal_fn.box = vmapped_prox  # lets apg_mpc hand the box to the device-side solver
    cfg_dict["_time_steps"] = np.array(compute_timesteps(cfg_dict["model"]), dtype=np.float32)
    n_u = cfg_dict["model"]["n_u"]
    horizon_keys = {k: cfg_dict["model"][k] for k in ("horizon", "num_short_dt", "short_step_dt", "long_step_dt")}
    _n_step_function = load_predictor_function(cfg_dict["learned_model_params"],
                                               modified_params={"num_particles": 1, **horizon_keys},
                                               prior_dist=nominal_model, return_control=True)
    default_uref = np.asarray(cfg_dict["cost_params"]["uref"], np.float32)

    def m_reset(x=None, xdes=None, uref=default_uref, rng=None, _pol_params=None):
        return reset_apg(x, cfg_dict["apg_mpc"], construct_opt_params, rng=rng, uref=uref, _pol_fn=None,
                         convert_to_enu=convert_to_enu, sample_sde=_n_step_function)

    if "trajectory_path" in cfg_dict:
        print("TRAJECTORY TRACKING")
        time_evol_traj, state_evol_traj = parse_trajectory(load_trajectory(cfg_dict["trajectory_path"]))
        cum_time_steps = np.cumsum(np.array(cfg_dict["_time_steps"]))

        def m_mpc(xt, rng, past_solution, curr_t=0.0, xdes=None, _val_params=None, return_ref=False, mod_cost_params={}):
            _, xref = extract_targets(curr_t, time_evol_traj, state_evol_traj, cum_time_steps)
            res = apg_mpc(xt, rng, past_solution, xref, multi_cost_sampling, proximal_fn, cfg_dict, _sde_learned,
                          _terminal_cost=None, extra_dyn_args=None, convert_to_enu=convert_to_enu,
                          mod_cost_params=mod_cost_params)
            return (res, xref) if return_ref else res

        def get_state_at_t(curr_t, _cum_time_steps=None, _return_evol=False):
            xcurr, _xevol = extract_targets(curr_t, time_evol_traj, state_evol_traj,
                                            cum_time_steps if _cum_time_steps is None else _cum_time_steps)
            return (xcurr, _xevol) if _return_evol else xcurr

        cfg_dict["state_traj"] = state_evol_traj
        cfg_dict["time_traj"] = time_evol_traj
    else:
        print("SETPOINT-BASED TRACKING")
        term = None
        if "cost_params_end" in cfg_dict:  # user cost-to-go: the stage cost form without control terms
            assert "uerr" not in cfg_dict["cost_params_end"], "uerr not supported in cost_params_end"
            assert "uref" not in cfg_dict["cost_params_end"], "uref not supported in cost_params_end"
            term = _costs.one_step_cost_f(cfg_dict["cost_params_end"], n_u)

        def m_mpc(xt, rng, past_solution, curr_t=0.0, xdes=None, _val_params=None, return_ref=False, mod_cost_params={}):
            xtarget = np.tile(np.asarray(xdes, np.float32), (cfg_dict["_time_steps"].shape[0] + 1, 1))
            res = apg_mpc(xt, rng, past_solution, xtarget, multi_cost_sampling, proximal_fn, cfg_dict, _sde_learned,
                          _terminal_cost=term, extra_dyn_args=None, convert_to_enu=convert_to_enu,
                          mod_cost_params=mod_cost_params)
            return (res, xtarget) if return_ref else res

        get_state_at_t = None

    one_step = dict(cfg_dict.get("one_step_params", {}))
    _one_step_function = load_predictor_function(cfg_dict["learned_model_params"],
                                                 modified_params={"horizon": 1, "num_particles": 1, **one_step},
                                                 prior_dist=nominal_model)

    def one_step_prediction(xt, ut, rng):
        return _one_step_function(xt, ut, rng)[0, 1]

    if return_full_cost:
        def full_cost_fn(y, u, rng, terminal_cost_fn, extra_dyn_args=None, extra_cost_args=None):
            return create_full_cost_function(y, rng, extra_cost_args, multi_cost_sampling, cfg_dict, _sde_learned,
                                             terminal_cost_fn, extra_dyn_args, convert_to_enu)[0](u)
        return (cfg_dict, (m_reset, m_mpc), get_state_at_t, one_step_prediction), full_cost_fn
    return cfg_dict, (m_reset, m_mpc), get_state_at_t, one_step_prediction


def extract_policy(opt_state):
    """sde_mpc_design.py:427-432."""
    return {"avg_linesearch": opt_state.avg_linesearch, "M1": opt_state.yk[0, 0], "M2": opt_state.yk[0, 1],
            "num_steps": opt_state.num_steps, "grad_norm": opt_state.grad_sqr, "avg_stepsize": opt_state.avg_stepsize,
            "cost0": opt_state.init_cost, "costF": opt_state.opt_cost}

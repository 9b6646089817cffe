"""The model / MPC dictionaries of the BASELINE.json configs, as Python dicts.

They restate the *values* of the reference's YAML files (network sizes, scalings, priors, APG
parameters) so that tests and ``bench.py`` run without ``/root/reference`` (which does not exist
on the GPU box):
  * mass-spring-damper ..... sde4mbrlExamples/mass_spring_damper/mass_spring_damper.yaml:9-75
  * cartpole ............... sde4mbrlExamples/cartpole/cartpole_sde.yaml:12-75
  * hexacopter ............. sde4mbrlExamples/rotor_uav/hexa_ahg/optimizer_sde.yaml:12-90 with the
                             learned physical scalars of my_models/hexa_ahg_v4_sde.pkl
                             
  * hexacopter MPC ......... sde4mbrlExamples/rotor_uav/hexa_ahg/mpc_test_cfg.yaml
"""
import copy

import numpy as np


def msd_model(horizon=100, num_particles=200):
    return {
        "n_x": 2, "n_y": 2, "n_u": 1,
        "init_params": {"mass": 1.0, "kq_coeff": 1.0, "bqdot_coeff": 0.5},
        "noise_prior_params": [0.001, 0.01],
        "control": False, "control_dependent_noise": False,
        "residual_forces": {"type": "dnn", "hidden_layers": [4, 16], "activation_fn": "swish", "init_value": 0.001},
        "diffusion_density_nn": {
            "indx_noise_in": [0, 1], "indx_noise_out": [1],
            "scaler_nn": {"type": "scaler", "init_value": 0.01},
            "density_nn": {"init_value": 0.1, "activation_fn": "swish", "hidden_layers": [32, 32]}},
        "sde_solver": "euler_maruyama", "stepsize": 0.01, "num_particles": num_particles, "horizon": horizon,
    }


def cartpole_model(horizon=50, num_particles=1024, side_info=True):
    m = {
        "n_x": 4, "n_y": 4, "n_u": 1,
        "residual_forces": {"init_value": 0.01, "hidden_layers": [8, 24], "activation_fn": "tanh"},
        "noise_prior_params": [0.005, 0.05, 0.004, 0.01],
        "side_info": side_info, "control_dependent_noise": False,
        "state_scaling": [9, 5, 1, 10],
        # learned output scaling stored in my_models/cartpole_bb_learned_si_sde.pkl
        "nn_out_scale": [40.37018739297579, 77.33879519500572],
        "diffusion_density_nn": {
            "indx_noise_in": [1, 2, 3], "indx_noise_out": [1, 3],
            "scaler_nn": {"type": "scaler", "init_value": 0.01},
            "density_nn": {"init_value": 0.1, "activation_fn": "swish", "hidden_layers": [32, 32]}},
        "stepsize": 0.02, "num_particles": num_particles, "horizon": horizon,
    }
    if side_info:
        m["control_nn"] = {"init_value": 0.01, "hidden_layers": [6, 8], "activation_fn": "tanh"}
    return m


HEXA_LEARNED_SCALARS = {  # my_models/hexa_ahg_v4_sde.pkl
    "mass": 2.328, "Ixx": 0.0344, "Iyy": 0.0302, "Izz": 0.0811, "Lmx": 0.320, "Lmy": 0.149,
    "CoeffDrag": 0.0339, "coeffMot0": 9.169, "coeffMot1": 0.786,
    "aero_kdx": 1.160, "aero_kdy": 1.483, "aero_kdz": 0.892, "aero_kh": 0.0,
}


def hexa_model(horizon=100, num_particles=4096, stepsize=0.01, density=(32, 32), density_act="swish"):
    return {
        "n_x": 13, "n_y": 13, "n_u": 6,
        "noise_prior_params": [0.001, 0.001, 0.001, 0.01, 0.01, 0.01, 0.0, 0.0, 0.0, 0.0, 0.05, 0.05, 0.01],
        "init_params": {"mass": 2.295, "gravity": 9.81, "Ixx": 0.008, "Iyy": 0.008, "Izz": 0.011,
                        "fm_model": {"order": 1, "use_constant_term": True, "assume_sym": False}},
        "fixed_init_params": False, "control_dependent_noise": False,
        "state_scaling": [2.2, 2.1, 2.7, 2.4, 2.3, 1.5, 1, 0.3, 0.3, 0.6, 2.7, 2.4, 1.5],
        "diffusion_density_nn": {
            "indx_noise_in": [3, 4, 5, 10, 11, 12], "indx_noise_out": [3, 4, 5, 10, 11, 12],
            "scaler_nn": {"type": "scaler", "init_value": 0.01},
            "density_nn": {"init_value": 0.001, "activation_fn": density_act, "hidden_layers": list(density)}},
        "ground_effect": False, "aero_drag_effect": True,
        "residual_forces": {"active_indx": [0, 1, 2], "hidden_layers": [8, 16], "activation_fn": "tanh"},
        "residual_moments": {"hidden_layers": [8, 16], "activation_fn": "tanh"},
        "sde_solver": "euler_maruyama", "stepsize": stepsize, "num_particles": num_particles, "horizon": horizon,
    }


def hexa_mpc(horizon=20, dt=0.05, num_particles=1, max_iter=20):
    """hexa_ahg/mpc_test_cfg.yaml (max_iter shortened to the BASELINE's fixed 20 iterations)."""
    return {
        "input_constr": {"input_id": [0, 1, 2, 3, 4, 5], "input_bound": [[1.0e-4, 1.0]] * 6},
        "cost_params": {"uref": [0.33] * 6, "uerr": 0.01, "perr": [20.0, 20.0, 40.0], "verr": [3.0, 3.0, 3.0],
                        "qerr": [0.1, 0.1, 20], "werr": [3.0, 3.0, 3.0], "res_mult": 0.01, "u_slew_coeff": 0.1},
        "horizon": horizon, "num_short_dt": horizon, "short_step_dt": dt, "long_step_dt": dt,
        "num_particles": num_particles,
        "apg_mpc": {"stepsize": 1.0, "max_iter": max_iter, "max_no_improvement_iter": max_iter, "moment_scale": None,
                    "beta_init": 0.25, "atol": 0.0, "rtol": 0.0,
                    "linesearch": {"init_stepsize": 0.01, "max_stepsize": 1.0, "coef": 0.01, "decrease_factor": 0.7,
                                   "increase_factor": 1.3, "reset_option": "increase", "maxls": 4}},
    }


def cartpole_mpc(horizon=50, dt=0.02, num_particles=1024, max_iter=20):
    """BASELINE C2.  The reference has no cartpole MPC script; the swing-up tracking cost is ours
    (SURVEY 8d): quadratic in [x, xdot, sin th, cos th, thdot] towards the upright equilibrium."""
    m = hexa_mpc(horizon, dt, num_particles, max_iter)
    m["input_constr"] = {"input_id": [0], "input_bound": [[-1.0, 1.0]]}
    m["cost_params"] = {"feat": "cartpole", "xerr": [1.0, 0.1, 5.0, 5.0, 0.1], "uerr": [0.01], "uref": [0.0],
                        "res_mult": 1.0}
    return m


def random_params(model, seed=0, scale="fan_in"):
    """Random-init parameter tree for benchmarking.  ``scale='yaml'`` uses the YAML init ranges
    (weights ~1e-3: numerically trivial nets); ``'fan_in'`` uses U(-1/sqrt(in), 1/sqrt(in)) so the
    networks are numerically non-trivial."""
    tree = model.init_params(seed)
    rng = np.random.default_rng(seed + 1)
    if scale == "fan_in":
        for k, v in tree.items():
            if "w" in v and "b" in v:
                fan = v["w"].shape[0]
                lim = 1.0 / np.sqrt(fan)
                v["w"] = rng.uniform(-lim, lim, v["w"].shape).astype(np.float32)
                v["b"] = rng.uniform(-0.1, 0.1, v["b"].shape).astype(np.float32)
                if k.endswith("density/~/linear_2"):
                    # centre eta = sigmoid(aggr*d - 7) around 0.5 so the diffusion path is exercised
                    v["b"] = v["b"] + np.float32(7.0)
    top = tree.get(model.module_name, None)
    if model.module_name == "sde_rotor_model" and top is not None:
        for k, v in HEXA_LEARNED_SCALARS.items():
            if k in top:
                top[k] = np.float32(v)
        # keep the residual nets small relative to gravity so 100-step rollouts stay bounded
        if scale == "fan_in":
            for k, v in tree.items():
                if k.endswith("linear_2") and "init_residual_networks" in k:
                    v["w"] *= np.float32(0.05)
                    v["b"] *= np.float32(0.05)
    if model.module_name == "cart_pole_sde" and scale == "fan_in":
        for k, v in tree.items():
            if k.endswith("linear_2") and "init_residual_networks" in k:
                v["w"] *= np.float32(0.02)
                v["b"] *= np.float32(0.02)
    return tree


def hover_state():
    x = np.zeros(13, np.float32)
    x[2] = 1.0
    x[6] = 1.0
    return x


def clone(d):
    return copy.deepcopy(d)

"""Shared test helpers: build (product model, oracle model, params) triples for the BASELINE
configs at reduced size and convert between the reference's logical layouts and the kernels'
sample-minor layouts."""
import numpy as np
import torch

from oracle import nsde_oracle as orc
from sde4mbrl_b200 import configs, sde_models

KINDS = {
    "msd": (configs.msd_model, sde_models.MassSpringDamper, "mass_spring_damper"),
    "cartpole": (configs.cartpole_model, sde_models.CartPoleSDE, "cart_pole_sde"),
    "cartpole_bb": (lambda **k: configs.cartpole_model(side_info=False, **k), sde_models.CartPoleSDE, "cart_pole_sde"),
    "hexa": (configs.hexa_model, sde_models.SDERotorModel, "sde_rotor_model"),
    "hexa16": (lambda **k: configs.hexa_model(density=(16, 16), density_act="tanh", **k), sde_models.SDERotorModel,
               "sde_rotor_model"),
}


def build(kind, horizon, num_particles, seed=0, scale="fan_in", **kw):
    mk, cls, oname = KINDS[kind]
    pm = mk(horizon=horizon, num_particles=num_particles, **kw)
    model = cls(pm)
    nn = configs.random_params(model, seed=seed, scale=scale)
    return pm, model, nn


def oracle_model(kind, pm, nn, dtype=torch.float32):
    return orc.make_model(KINDS[kind][2], pm, nn, dtype)


def start_state(kind, rng):
    if kind == "msd":
        return rng.uniform(-0.1, 0.1, 2).astype(np.float32)
    if kind.startswith("cartpole"):
        return (np.array([0.0, 0.0, np.pi, 0.0]) + rng.normal(0, 0.1, 4)).astype(np.float32)
    x = configs.hover_state()
    x[3:6] += rng.normal(0, 0.1, 3).astype(np.float32)
    x[10:13] += rng.normal(0, 0.1, 3).astype(np.float32)
    q = np.array([1.0, 0.05, -0.03, 0.02])
    x[6:10] = (q / np.linalg.norm(q)).astype(np.float32)
    return x


def controls(kind, H, rng):
    if kind == "msd":
        return np.zeros((H, 1), np.float32)
    if kind.startswith("cartpole"):
        return rng.uniform(-1, 1, (H, 1)).astype(np.float32)
    return rng.uniform(0.30, 0.36, (H, 6)).astype(np.float32)


def to_native_noise(dw):
    """[N, H, n_x] (or [P, N, H, n_x]) -> [H, n_x, G] sample-minor."""
    dw = np.asarray(dw, np.float32)
    if dw.ndim == 3:
        return np.ascontiguousarray(dw.transpose(1, 2, 0))
    P, N, H, nx = dw.shape
    return np.ascontiguousarray(dw.reshape(P * N, H, nx).transpose(1, 2, 0))


def from_native_x(xbuf, n_x):
    """[H+1, rows, G] -> numpy [G, H+1, n_x]."""
    return xbuf[:, :n_x, :].permute(2, 0, 1).cpu().numpy()


def stage_cost_fn(kind, cp):
    if kind.startswith("hexa"):
        return lambda x, u, xref: orc.rotor_one_step_cost(x, u, xref, cp)
    return lambda x, u, xref: orc.quadratic_one_step_cost(x, u, xref, cp)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))

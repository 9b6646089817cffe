"""CUDA path vs the committed golden fixtures (tests/golden/, made by make_golden.py) and vs the
oracle on the reference's SHIPPED learned weights."""
import glob
import os

import numpy as np
import pytest
import torch

import helpers as hp
from oracle import nsde_oracle as orc
from oracle import threefry as tf
from sde4mbrl_b200 import configs, costs, io, sde_models

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cost_for(kind, H):
    if kind.startswith("hexa"):
        cp = configs.hexa_mpc()["cost_params"]
        return costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    if kind.startswith("cartpole"):
        cp = configs.cartpole_mpc()["cost_params"]
        return costs.QuadraticCost(cp["xerr"], cp["uerr"], cp["uref"], feat="cartpole", res_mult=cp["res_mult"])
    return costs.QuadraticCost([2.0, 0.5], [0.1], [0.0])


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "golden_*.npz"))))
def test_cuda_matches_golden(cuda, name):
    from sde4mbrl_b200.engine import Engine
    g = np.load(os.path.join(GOLD, name))
    kind = name.split("_")[1]
    H, N, seed, mode = int(g["H"]), int(g["N"]), int(g["seed"]), int(g["rng_mode"])
    pm, model, nn = hp.build(kind, H, N, seed=seed)
    eng = Engine(model, rng_mode=mode)
    ts = orc.compute_timesteps(pm)
    xbuf = eng.rollout(nn, g["y0"][None], g["u"], ts, key=g["key"], N=N)  # in-kernel threefry noise
    assert hp.rel_err(hp.from_native_x(xbuf, model.n_x), g["xevol"]) < (1e-4 if H > 50 else 2e-5)
    cost, grad, _ = eng.cost_grad(nn, g["y0"][None], g["u"], ts, g["xref"], _cost_for(kind, H).desc, key=g["key"], N=N)
    assert abs(float(cost[0]) - float(g["cost"])) <= 3e-5 * abs(float(g["cost"])) + 1e-7
    assert hp.rel_err(grad.cpu().numpy(), g["grad"]) < 1e-3


@pytest.mark.parametrize("fixture,cls,kind", [
    ("ref_weights_hexa_ahg_v4.npz", sde_models.SDERotorModel, "sde_rotor_model"),
    ("ref_weights_iris_sitl.npz", sde_models.SDERotorModel, "sde_rotor_model"),
    ("ref_weights_cartpole_bb_learned_si.npz", sde_models.CartPoleSDE, "cart_pole_sde")])
def test_shipped_weights_rollout_parity(cuda, fixture, cls, kind):
    """The reference's own learned models (hexa_ahg_v4, iris_sitl, cartpole side-info), their own
    horizon / stepsize, 32 particles."""
    from sde4mbrl_b200.engine import Engine
    nn, nominal = io.load_npz_fixture(os.path.join(GOLD, fixture))
    N = 32
    model = cls(nominal)
    om = orc.make_model(kind, nominal, nn)
    rng = np.random.default_rng(4)
    H = nominal["horizon"]
    if kind == "sde_rotor_model":
        y0 = hp.start_state("hexa", rng)
        u = rng.uniform(0.25, 0.45, (H, nominal["n_u"])).astype(np.float32)
    else:
        y0 = hp.start_state("cartpole", rng)
        u = rng.uniform(-1, 1, (H, 1)).astype(np.float32)
    key = tf.PRNGKey(11)
    want, _ = orc.sample_rollouts(om, y0, u, key, N)
    xbuf = Engine(model).rollout(nn, y0[None], u, orc.compute_timesteps(nominal), key=key, N=N)
    assert hp.rel_err(hp.from_native_x(xbuf, model.n_x), want.numpy()) < 2e-5

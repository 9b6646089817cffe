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


@pytest.mark.parametrize("lanes", [0, 1])
@pytest.mark.parametrize("fixture,cls,kind", [
    ("ref_weights_hexa_ahg_v4.npz", sde_models.SDERotorModel, "sde_rotor_model"),
    ("ref_weights_iris_sitl.npz", sde_models.SDERotorModel, "sde_rotor_model"),
    ("ref_weights_cartpole_bb_learned_si.npz", sde_models.CartPoleSDE, "cart_pole_sde")])
def test_shipped_weights_cost_grad_parity(cuda, fixture, cls, kind, lanes):
    """MPC cost and its gradient (not only rollouts) on the reference's own learned models: float32 value against the
    float32 oracle, gradient against float64 autograd of the oracle on the kernel's own threefry noise."""
    from sde4mbrl_b200.engine import Engine
    nn, nominal = io.load_npz_fixture(os.path.join(GOLD, fixture))
    N = 24
    model = cls(nominal)
    rng = np.random.default_rng(6)
    H = nominal["horizon"]
    n_u = nominal["n_u"]
    if kind == "sde_rotor_model":
        y0 = hp.start_state("hexa", rng)
        u = rng.uniform(0.25, 0.45, (H, n_u)).astype(np.float32)
        cp = dict(configs.hexa_mpc()["cost_params"])
        for k_ in ("uref", "uerr", "u_slew_coeff"):  # per-rotor lists of the hexacopter config, cut to this vehicle
            if k_ in cp and np.ndim(cp[k_]) > 0:
                cp[k_] = list(cp[k_])[:n_u]
        stage = costs.with_slew(costs.one_step_cost_f(cp, n_u), cp, n_u)
        cost_fn = lambda x, uu, xr: orc.rotor_one_step_cost(x, uu, xr, cp)
        xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
        xref[:, :3] = [0.0, 0.5, 1.4]
    else:
        y0 = hp.start_state("cartpole", rng)
        u = rng.uniform(-1, 1, (H, 1)).astype(np.float32)
        cp = configs.cartpole_mpc()["cost_params"]
        stage = costs.QuadraticCost(cp["xerr"], cp["uerr"], cp["uref"], feat="cartpole", res_mult=cp["res_mult"])
        cost_fn = lambda x, uu, xr: orc.quadratic_one_step_cost(x, uu, xr, cp)
        xref = np.zeros((H + 1, 4), np.float32)
    key = tf.PRNGKey(21)
    ts = orc.compute_timesteps(nominal)
    c, g, _ = Engine(model).cost_grad(nn, y0[None], u, ts, xref, stage.desc, key=key, N=N, lanes=lanes)
    _, dw = orc.sample_rollouts(orc.make_model(kind, nominal, nn), y0, u, key, N)
    om = orc.make_model(kind, nominal, nn, torch.float64)
    o = torch.tensor(u, dtype=torch.float64, requires_grad=True)
    c64, _ = orc.compute_cost(om, torch.tensor(y0, dtype=torch.float64), o, dw.double(), cost_fn,
                              torch.tensor(xref, dtype=torch.float64))
    if kind == "sde_rotor_model":
        c64 = c64 + orc.u_slew_cost(o, ts, cp, n_u)
    (g64,) = torch.autograd.grad(c64, o)
    assert abs(float(c[0]) - c64.item()) <= 2e-5 * abs(c64.item())
    assert hp.rel_err(g.cpu().numpy(), g64.numpy()) < 1e-3

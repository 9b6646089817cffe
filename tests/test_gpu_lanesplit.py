"""Lane-split kernels (8 adjacent lanes per sample, nets.cuh EvL): same results as the
one-thread-per-sample kernels and as the oracle."""
import numpy as np
import pytest
import torch

import helpers as hp
from oracle import nsde_oracle as orc
from oracle import threefry as tf
from sde4mbrl_b200 import configs, costs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,H,N,P", [("hexa", 100, 64, 1), ("hexa16", 20, 40, 1), ("hexa", 20, 3, 5), ("hexa", 20, 4, 3)])
def test_lane_split_rollout(cuda, kind, H, N, P):
    from sde4mbrl_b200.engine import Engine
    rng = np.random.default_rng(1)
    pm, model, nn = hp.build(kind, H, N, seed=3)
    om = hp.oracle_model(kind, pm, nn)
    y0 = np.stack([hp.start_state(kind, rng) for _ in range(P)])
    u = np.stack([hp.controls(kind, H, rng) for _ in range(P)])
    keys = tf.split(tf.PRNGKey(8), P)
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    x8 = hp.from_native_x(eng.rollout(nn, y0, u, ts, key=keys, N=N, lanes=8), 13).reshape(P, N, H + 1, 13)
    x1 = hp.from_native_x(eng.rollout(nn, y0, u, ts, key=keys, N=N, lanes=1), 13).reshape(P, N, H + 1, 13)
    tol = 1e-4 if H > 50 else 2e-5
    assert hp.rel_err(x8, x1) < tol
    for p in range(P):
        want, _ = orc.sample_rollouts(om, y0[p], u[p], keys[p], N)
        assert hp.rel_err(x8[p], want.numpy()) < tol


@pytest.mark.parametrize("kind,H,N", [("hexa", 100, 32), ("hexa", 20, 4), ("hexa", 20, 5), ("hexa16", 20, 33)])
def test_lane_split_cost_grad(cuda, kind, H, N):
    from sde4mbrl_b200.engine import Engine
    rng = np.random.default_rng(5)
    pm, model, nn = hp.build(kind, H, N, seed=11)
    y0, opt = hp.start_state(kind, rng), hp.controls(kind, H, rng)
    dw = rng.standard_normal((N, H, 13)).astype(np.float32)
    cp = configs.hexa_mpc()["cost_params"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    xref[:, :3] = [0.0, 0.5, 1.4]
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    nat = hp.to_native_noise(dw)
    c8, g8, xt8 = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, N=N, noise=nat, want_xtraj=True, lanes=8)
    c1, g1, xt1 = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, N=N, noise=nat, want_xtraj=True, lanes=1)
    assert abs(float(c8[0]) - float(c1[0])) <= 2e-5 * abs(float(c1[0]))
    assert hp.rel_err(g8.cpu().numpy(), g1.cpu().numpy()) < 2e-4
    assert hp.rel_err(xt8.cpu().numpy(), xt1.cpu().numpy()) < (1e-4 if H > 50 else 2e-5)
    om = hp.oracle_model(kind, pm, nn, torch.float64)
    o = torch.tensor(opt, dtype=torch.float64, requires_grad=True)
    c64, _ = orc.compute_cost(om, torch.tensor(y0, dtype=torch.float64), o, torch.tensor(dw, dtype=torch.float64),
                              hp.stage_cost_fn(kind, cp), torch.tensor(xref, dtype=torch.float64))
    c64 = c64 + orc.u_slew_cost(o, ts, cp, 6)
    (g64,) = torch.autograd.grad(c64, o)
    assert abs(float(c8[0]) - c64.item()) <= 2e-5 * abs(c64.item())
    assert hp.rel_err(g8.cpu().numpy(), g64.numpy()) < 1e-3
    # threefry noise path + value-only launch
    key = tf.PRNGKey(3)
    ca, ga, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, N=N, key=key, lanes=8)
    cb, gb, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, N=N, key=key, lanes=1)
    cc, _, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, N=N, key=key, lanes=8, want_grad=False)
    assert abs(float(ca[0]) - float(cb[0])) <= 2e-5 * abs(float(cb[0])) \
        and abs(float(cc[0]) - float(ca[0])) <= 2e-6 * abs(float(ca[0]))
    assert hp.rel_err(ga.cpu().numpy(), gb.cpu().numpy()) < 2e-4


def test_requesting_missing_lane_kernel_is_an_error(cuda):
    from sde4mbrl_b200.engine import Engine
    from sde4mbrl_b200._lib import Sde4Error
    pm, model, nn = hp.build("cartpole", 5, 2, seed=1)  # compiled with 1 and 8 lanes per sample, not 4
    with pytest.raises(Sde4Error):
        Engine(model).rollout(nn, np.zeros((1, 4), np.float32), np.zeros((5, 1), np.float32),
                              orc.compute_timesteps(pm), key=tf.PRNGKey(0), N=2, lanes=4)


@pytest.mark.parametrize("kind,lanes,H,N", [("cartpole", 8, 50, 37), ("cartpole_bb", 8, 30, 21)])
def test_lane_split_cartpole(cuda, kind, lanes, H, N):
    """Cartpole lane-split kernels (8 lanes; the 2->6->8 control net of the side-info model runs PADDED: 6 real + 2 ghost
    units) vs the one-thread-per-sample ones."""
    from sde4mbrl_b200.engine import Engine
    rng = np.random.default_rng(7)
    pm, model, nn = hp.build(kind, H, N, seed=5)
    y0, opt = hp.start_state(kind, rng), hp.controls(kind, H, rng)
    cp = configs.cartpole_mpc()["cost_params"]
    stage = costs.QuadraticCost(cp["xerr"], cp["uerr"], cp["uref"], feat="cartpole", res_mult=cp["res_mult"])
    xref = np.zeros((H + 1, 4), np.float32)
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    key = tf.PRNGKey(4)
    xl = eng.rollout(nn, y0[None], opt, ts, key=key, N=N, lanes=lanes)
    x1 = eng.rollout(nn, y0[None], opt, ts, key=key, N=N, lanes=1)
    assert hp.rel_err(xl.cpu().numpy(), x1.cpu().numpy()) < 2e-5
    cl, gl, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, N=N, key=key, lanes=lanes)
    c1, g1, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, N=N, key=key, lanes=1)
    assert abs(float(cl[0]) - float(c1[0])) <= 2e-5 * abs(float(c1[0]))
    assert hp.rel_err(gl.cpu().numpy(), g1.cpu().numpy()) < 2e-4
    cv, _, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, N=N, key=key, lanes=lanes, want_grad=False)
    assert abs(float(cv[0]) - float(c1[0])) <= 2e-5 * abs(float(c1[0]))

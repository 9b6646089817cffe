"""``control_dependent_noise: True`` (SURVEY 8a row a3, sde4mbrl/nsde.py:362): the density network of the distance-aware
diffusion reads ``concat(reduced_state[indx_noise_in], u)``, so the controls also act through sigma(x, u) dw and the
MPC gradient gains a second path.  Rollout, cost + gradient and the training loss vs the oracle.

Tolerances as in test_gpu_costgrad.py / test_gpu_loss.py (float32 kernels against float32 values and float64 autograd)."""
import numpy as np
import pytest
import torch

import helpers as hp
from oracle import nsde_oracle as orc
from oracle import threefry as tf
from sde4mbrl_b200 import configs, costs
from sde4mbrl_b200._lib import Sde4Error

pytestmark = pytest.mark.gpu


def _build(kind, H, N, seed):
    mk, cls, oname = hp.KINDS[kind]
    pm = mk(horizon=H, num_particles=N)
    pm["control_dependent_noise"] = True
    model = cls(pm)
    nn = configs.random_params(model, seed=seed)
    return pm, model, nn, oname


def _problem(kind, H, model, rng):
    y0 = hp.start_state(kind, rng)
    opt = hp.controls(kind, H, rng)
    if kind == "msd":
        opt = rng.uniform(-0.5, 0.5, (H, 1)).astype(np.float32)
    if kind.startswith("hexa"):
        cp = configs.hexa_mpc()["cost_params"]
        stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
        xref = np.tile(configs.hover_state(), (H + 1, 1))
        xref[:, :3] = [0.0, 0.5, 1.4]
    elif kind.startswith("cartpole"):
        cp = configs.cartpole_mpc()["cost_params"]
        stage = costs.QuadraticCost(cp["xerr"], cp["uerr"], cp["uref"], feat="cartpole", res_mult=cp["res_mult"])
        xref = np.zeros((H + 1, 4), np.float32)
    else:
        cp = {"feat": "identity", "xerr": [2.0, 0.5], "uerr": [0.1], "uref": [0.0], "res_mult": 1.0}
        stage = costs.QuadraticCost(cp["xerr"], cp["uerr"], cp["uref"])
        xref = np.tile(np.array([0.05, 0.0], np.float32), (H + 1, 1))
    return y0, opt, cp, stage, xref.astype(np.float32)


def _oracle_cost(kind, oname, pm, nn, y0, opt, dw, cp, xref, dtype):
    om = orc.make_model(oname, pm, nn, dtype)
    o = torch.tensor(opt, dtype=dtype, requires_grad=True)
    c, xtraj = orc.compute_cost(om, torch.tensor(y0, dtype=dtype), o, torch.tensor(dw, dtype=dtype),
                                hp.stage_cost_fn(kind, cp), torch.tensor(xref, dtype=dtype))
    if kind.startswith("hexa"):
        c = c + orc.u_slew_cost(o, orc.compute_timesteps(pm), cp, om.n_u)
    (g,) = torch.autograd.grad(c, o)
    return float(c), g.numpy(), xtraj.detach().numpy()


@pytest.mark.parametrize("kind,H,N", [("msd", 40, 33), ("cartpole", 30, 64), ("hexa", 20, 16)])
def test_cost_and_grad_with_control_dependent_noise(cuda, kind, H, N):
    from sde4mbrl_b200.engine import Engine
    pm, model, nn, oname = _build(kind, H, N, seed=11)
    desc = model.describe()
    assert desc.density.n_in == bin(desc.noise_in_mask).count("1") + model.n_u
    rng = np.random.default_rng(3)
    y0, opt, cp, stage, xref = _problem(kind, H, model, rng)
    dw = rng.standard_normal((N, H, model.n_x)).astype(np.float32)
    c64, g64, _ = _oracle_cost(kind, oname, pm, nn, y0, opt, dw, cp, xref, torch.float64)
    _, _, xt32 = _oracle_cost(kind, oname, pm, nn, y0, opt, dw, cp, xref, torch.float32)
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    cost, grad, xtraj = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, N=N, noise=hp.to_native_noise(dw),
                                      want_grad=True, want_xtraj=True)
    assert abs(float(cost[0]) - c64) <= 2e-5 * abs(c64) + 1e-7
    got = xtraj.permute(2, 0, 1).cpu().numpy()
    assert hp.rel_err(got[..., :model.n_x], xt32[..., :model.n_x]) < 1e-4
    assert hp.rel_err(grad.cpu().numpy(), g64) < 1e-3
    # the control path through the diffusion is really there: the same weights with the control rows of the density
    # net's first layer zeroed give a different gradient (for the MSD, whose drift ignores u, a zero one)
    nn0 = {m: {k: np.array(v) for k, v in d.items()} for m, d in nn.items()}
    key0 = [k for k in nn0 if k.endswith("density/~/linear_0")][0]
    nn0[key0]["w"][-model.n_u:] = 0.0
    _, grad0, _ = eng.cost_grad(nn0, y0[None], opt, ts, xref, stage.desc, N=N, noise=hp.to_native_noise(dw), want_grad=True)
    _, g64_0, _ = _oracle_cost(kind, oname, pm, nn0, y0, opt, dw, cp, xref, torch.float64)
    assert hp.rel_err(grad0.cpu().numpy(), g64_0) < 1e-3
    assert np.max(np.abs(g64 - g64_0)) > 0.0
    if kind == "msd":  # control: False -- the drift ignores u, the whole gradient beyond the cost's own u-term is the diffusion path
        assert np.max(np.abs(g64 - g64_0)) > 1e-4 * np.max(np.abs(g64))
    # value-only launch
    cost2, g2, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, N=N, noise=hp.to_native_noise(dw), want_grad=False)
    assert g2 is None and abs(float(cost2[0]) - float(cost[0])) <= 2e-6 * abs(float(cost[0]))


@pytest.mark.parametrize("kind,H,N", [("msd", 25, 9), ("cartpole", 20, 17), ("hexa", 15, 5)])
def test_rollout_with_control_dependent_noise_and_threefry_keys(cuda, kind, H, N):
    from sde4mbrl_b200 import nsde, random as R
    mk, cls, oname = hp.KINDS[kind]
    pm = mk(horizon=H, num_particles=N)
    pm["control_dependent_noise"] = True
    nn = configs.random_params(cls(pm), seed=2)
    _, multi_sampling = nsde.create_sampling_fn(pm, sde_constr=cls, seed=0)
    rng = np.random.default_rng(5)
    y0 = hp.start_state(kind, rng)
    u = hp.controls(kind, H, rng) if kind != "msd" else rng.uniform(-0.5, 0.5, (H, 1)).astype(np.float32)
    xevol, yevol, uevol = multi_sampling(nn, y0, u, R.PRNGKey(7))
    om = orc.make_model(oname, pm, {m: {k: torch.tensor(np.asarray(v)) for k, v in d.items()} for m, d in nn.items()})
    want, _ = orc.sample_rollouts(om, y0, u, tf.PRNGKey(7), N)
    assert tuple(xevol.shape) == (N, H + 1, pm["n_x"])
    assert hp.rel_err(xevol.cpu().numpy(), want.numpy()) < 5e-5


def test_training_loss_with_control_dependent_noise(cuda):
    """Parameter gradients, including the control rows of the density net's first layer."""
    from sde4mbrl_b200 import nsde, random as R
    mk, cls, oname = hp.KINDS["cartpole"]
    B, N, H = 7, 2, 12
    pm = mk(horizon=1, num_particles=1)
    pm["control_dependent_noise"] = True
    ow = [9, 5, 1, 1.0, 10]
    lp = {"seed": 1, "stepsize": pm["stepsize"], "data_stepsize": pm["stepsize"], "num_particles": N, "num_particles_test": 4,
          "horizon": H, "obs_weights": ow, "pen_data": 1.0, "pen_weights": 0.0}
    _, loss_fn, _, _ = nsde.create_model_loss_fn(pm, lp, sde_constr=cls, verbose=False)
    nn = configs.random_params(cls(pm), seed=4)
    rng = np.random.default_rng(0)
    y = (rng.normal(0, 0.5, (B, H + 1, 4)) + np.array([0, 0, 3.0, 0])).astype(np.float32)
    u = rng.uniform(-1, 1, (B, H, 1)).astype(np.float32)
    total, info, grads = loss_fn.value_and_grad(nn, y, u, R.PRNGKey(3))
    tree = {m: {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=True) for k, v in d.items()} for m, d in nn.items()}
    om = orc.make_model(oname, pm, tree, torch.float64)
    dw = torch.tensor(orc.loss_noise(tf.PRNGKey(3), B, N, H, 4), dtype=torch.float64)
    Ld, _ = orc.data_loss(om, oname, torch.tensor(y, dtype=torch.float64), torch.tensor(u, dtype=torch.float64), dw, ow)
    Ld.backward()
    assert abs(info["dataLoss"] - float(Ld.detach())) < 3e-5 * abs(float(Ld.detach()))
    gmax = max(float(v.grad.abs().max()) for d in tree.values() for v in d.values())
    for m, d in tree.items():
        for k, v in d.items():
            want = v.grad.numpy()
            assert grads[m][k].shape == want.shape
            assert np.max(np.abs(grads[m][k] - want)) <= 2e-3 * np.max(np.abs(want)) + 1e-5 * gmax, (m, k)
    w0 = [k for k in tree if k.endswith("density/~/linear_0")][0]
    assert tree[w0]["w"].shape[0] == 4 and float(tree[w0]["w"].grad[-1].abs().max()) > 0.0


def test_architectures_without_a_control_dependent_instantiation_are_rejected(cuda):
    from sde4mbrl_b200.engine import Engine
    mk, cls, _ = hp.KINDS["cartpole_bb"]
    pm = mk(horizon=5, num_particles=2)
    pm["control_dependent_noise"] = True
    with pytest.raises(Sde4Error):
        Engine(cls(pm)).packed(configs.random_params(cls(pm), seed=0))


@pytest.mark.parametrize("kind,lanes,H,N", [("cartpole", 8, 40, 37), ("msd", 4, 30, 21)])
def test_lane_split_kernels_with_control_dependent_noise(cuda, kind, lanes, H, N):
    """Lane-split instantiations (in-kernel threefry noise, straight-line loop bodies; for the gradient the
    software-pipelined diffusion adjoint that evaluates the density net of step t-1 -- with u_{t-1} -- while step t is
    applied) vs the one-thread-per-sample kernels and the oracle on the same key."""
    from sde4mbrl_b200.engine import Engine
    pm, model, nn, oname = _build(kind, H, N, seed=5)
    rng = np.random.default_rng(7)
    y0, opt, cp, stage, xref = _problem(kind, H, model, rng)
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
    # oracle with the noise the key chain produces (nsde.py:1581 -> sde_solver.py:276,283)
    dw = tf.rollout_noise(key, N, H, model.n_x)
    c64, g64, _ = _oracle_cost(kind, oname, pm, nn, y0, opt, dw, cp, xref, torch.float64)
    assert abs(float(cl[0]) - c64) <= 2e-5 * abs(c64) + 1e-7
    assert hp.rel_err(gl.cpu().numpy(), g64) < 1e-3
    # shaped (sigmoid-residual) cost -> the MODE-1 instantiation; given noise -> the generic lane-split kernel
    nat = hp.to_native_noise(dw)
    cg, gg, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, N=N, noise=nat, lanes=lanes)
    assert abs(float(cg[0]) - c64) <= 2e-5 * abs(c64) + 1e-7
    assert hp.rel_err(gg.cpu().numpy(), g64) < 1e-3

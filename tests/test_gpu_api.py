"""The reference-facing Python API (nsde.py factories, apg.py solver) on the GPU: shapes, views,
differentiability through torch autograd, and APG iterate parity against the oracle APG driving
the oracle cost."""
import contextlib
import io as _io

import numpy as np
import pytest
import torch

import helpers as hp
from oracle import apg_oracle
from oracle import nsde_oracle as orc
from oracle import threefry as tf
from sde4mbrl_b200 import configs, costs, sde_models

pytestmark = pytest.mark.gpu


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(_io.StringIO()):
        return fn(*a, **k)


def test_create_sampling_fn_shapes_and_values(cuda):
    from sde4mbrl_b200 import nsde, random as R
    pm = configs.hexa_model(horizon=15, num_particles=9)
    nn0, multi_sampling = nsde.create_sampling_fn(pm, sde_constr=sde_models.SDERotorModel)
    assert "sde_rotor_model" in nn0
    nn = configs.random_params(sde_models.SDERotorModel(pm), seed=2)
    rng = np.random.default_rng(0)
    y0, u = hp.start_state("hexa", rng), hp.controls("hexa", 15, rng)
    key = R.PRNGKey(5)
    x, y, uev = multi_sampling(nn, y0, u, key)
    assert tuple(x.shape) == (9, 16, 13) and tuple(y.shape) == (9, 16, 13) and tuple(uev.shape) == (9, 15, 6)
    want, _ = orc.sample_rollouts(hp.oracle_model("hexa", pm, nn), y0, u, tf.PRNGKey(5), 9)
    assert hp.rel_err(x.cpu().numpy(), want.numpy()) < 2e-5
    np.testing.assert_array_equal(uev[3].cpu().numpy(), u)
    with pytest.raises(Exception):
        multi_sampling(nn, y0, u, key, extra_args=(1,))
    with pytest.raises(AssertionError):
        multi_sampling(nn, y0, u[:-1], key)


def test_one_step_sampling_squeezes(cuda):
    from sde4mbrl_b200 import nsde, random as R
    pm = configs.cartpole_model(horizon=7, num_particles=1)
    one = nsde.create_one_step_sampling(pm, sde_constr=sde_models.CartPoleSDE)
    nn = configs.random_params(sde_models.CartPoleSDE(pm), seed=1)
    x, _, uev = one(nn, np.array([0, 0, 3.0, 0], np.float32), np.array([0.3], np.float32), R.PRNGKey(1))
    assert tuple(x.shape) == (2, 4) and tuple(uev.shape) == (1, 1)
    many = nsde.create_one_step_sampling(pm, sde_constr=sde_models.CartPoleSDE, num_samples=10)
    x, _, _ = many(nn, np.array([0, 0, 3.0, 0], np.float32), np.array([0.3], np.float32), R.PRNGKey(1))
    assert tuple(x.shape) == (10, 2, 4)


def _mpc_setup(N, H=20, max_iter=6):
    from sde4mbrl_b200 import nsde
    pm = configs.hexa_model(horizon=H, num_particles=N, stepsize=0.05)
    mpc = configs.hexa_mpc(horizon=H, dt=0.05, num_particles=N, max_iter=max_iter)
    out = _quiet(nsde.create_online_cost_sampling_fn, pm, mpc, sde_constr=sde_models.SDERotorModel)
    nn = configs.random_params(sde_models.SDERotorModel(pm), seed=3)
    return pm, mpc, nn, out


def test_online_cost_sampling_value_grad_and_xtraj(cuda):
    from sde4mbrl_b200 import random as R
    N, H = 16, 20
    pm, mpc, nn, (_, multi_cost, prox, constr_cost, construct_opt) = _mpc_setup(N, H)
    assert constr_cost is None and prox is not None
    cp = mpc["cost_params"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    rng = np.random.default_rng(1)
    y0 = hp.start_state("hexa", rng)
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    opt = construct_opt(u=np.full(6, 0.33, np.float32))
    assert tuple(opt.shape) == (H, 6) and abs(float(opt[0, 0]) - 0.3302) < 1e-6  # the doubled +1e-4 (nsde.py:1545,1563)
    opt = opt.clone().requires_grad_(True)
    c, xtraj = multi_cost(nn, y0, opt, R.PRNGKey(9), stage, extra_cost_args=xref)
    (g,) = torch.autograd.grad(c, opt)
    assert tuple(xtraj.shape) == (N, H + 1, 14)
    dw = tf.rollout_noise(tf.PRNGKey(9), N, H, 13)
    om = hp.oracle_model("hexa", pm, nn, torch.float64)
    o = opt.detach().cpu().double().requires_grad_(True)
    c64, xt = orc.compute_cost(om, torch.tensor(y0, dtype=torch.float64), o, torch.tensor(dw, dtype=torch.float64),
                               hp.stage_cost_fn("hexa", cp), torch.tensor(xref, dtype=torch.float64))
    c64 = c64 + orc.u_slew_cost(o, orc.compute_timesteps(pm), cp, 6)
    (g64,) = torch.autograd.grad(c64, o)
    assert abs(float(c) - float(c64)) < 2e-5 * abs(float(c64))
    assert hp.rel_err(g.cpu().numpy(), g64.numpy()) < 1e-3
    assert hp.rel_err(xtraj.cpu().numpy(), xt.detach().numpy()) < 1e-4
    lo = float(prox(torch.full((H, 6), -3.0, device="cuda")).min())
    assert abs(lo - 1e-4) < 1e-9  # box clip to [1e-4, 1] (mpc_test_cfg.yaml:10-13)


def test_host_buffer_path_matches_device_path(cuda):
    """numpy / CPU-tensor arguments go through sde4_cost_grad_host (one packed copy each way) and must give
    the device path's numbers bit for bit; results come back on the host."""
    from sde4mbrl_b200 import random as R
    N, H = 16, 20
    pm, mpc, nn, (_, multi_cost, _, _, construct_opt) = _mpc_setup(N, H)
    cp = mpc["cost_params"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    rng = np.random.default_rng(4)
    y0 = hp.start_state("hexa", rng)
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    opt_np = hp.controls("hexa", H, rng)
    key_np = tf.PRNGKey(9)
    o_dev = torch.as_tensor(opt_np, device="cuda").requires_grad_(True)
    c_dev, xt_dev = multi_cost(nn, y0, o_dev, R.PRNGKey(9), stage, extra_cost_args=xref)
    (g_dev,) = torch.autograd.grad(c_dev, o_dev)
    o_host = torch.as_tensor(opt_np).clone().requires_grad_(True)
    c_host, xt_host = multi_cost(nn, y0, o_host, key_np, stage, extra_cost_args=xref)
    (g_host,) = torch.autograd.grad(c_host, o_host)
    assert not c_host.is_cuda and not g_host.is_cuda and xt_host.is_cuda
    assert float(c_host) == float(c_dev)
    np.testing.assert_array_equal(g_host.numpy(), g_dev.cpu().numpy())
    np.testing.assert_array_equal(xt_host.cpu().numpy(), xt_dev.cpu().numpy())
    # value_and_grad = jax.value_and_grad(multi_sampling, argnums=2, has_aux=True): no autograd graph
    (c3, xt3), g3 = multi_cost.value_and_grad(nn, y0, opt_np, key_np, stage, extra_cost_args=xref)
    assert float(c3) == float(c_dev) and tuple(xt3.shape) == (N, H + 1, 14)
    np.testing.assert_array_equal(g3.numpy(), g_dev.cpu().numpy())
    np.testing.assert_array_equal(xt3.cpu().numpy(), xt_dev.cpu().numpy())  # (the lean host route builds its own view)
    (c4, _), g4 = multi_cost.value_and_grad(nn, y0, o_dev.detach(), R.PRNGKey(9), stage, extra_cost_args=xref)
    assert c4.is_cuda and float(c4) == float(c_dev) and torch.equal(g4, g_dev)
    # value only, numpy in (no autograd)
    c2, _ = multi_cost(nn, y0, opt_np, key_np, stage, extra_cost_args=xref)
    assert abs(float(c2) - float(c_dev)) <= 2e-6 * abs(float(c_dev))
    # batched problems [P, H, n_opt] with per-problem start states, references and keys
    P = 3
    y0b = np.stack([hp.start_state("hexa", rng) for _ in range(P)])
    optb = np.stack([hp.controls("hexa", H, rng) for _ in range(P)])
    keyb = np.stack([tf.PRNGKey(20 + i) for i in range(P)])
    eng = multi_cost.engine
    ch, gh, _ = eng.cost_grad_host(nn, y0b, optb, multi_cost.time_steps, xref, stage.desc, key=keyb, N=N)
    cd_, gd_, _ = eng.cost_grad(nn, y0b, optb, multi_cost.time_steps, xref, stage.desc, key=keyb, N=N)
    np.testing.assert_array_equal(ch, cd_.cpu().numpy())
    np.testing.assert_array_equal(gh, gd_.cpu().numpy())
    # the same batch through value_and_grad: host arrays (lean route) against device tensors (generic route)
    (cb_h, xtb_h), gb_h = multi_cost.value_and_grad(nn, y0b, optb, keyb, stage, extra_cost_args=xref)
    (cb_d, xtb_d), gb_d = multi_cost.value_and_grad(nn, torch.as_tensor(y0b, device="cuda"), torch.as_tensor(optb, device="cuda"),
                                                    torch.as_tensor(keyb.view(np.int32), device="cuda"), stage,
                                                    extra_cost_args=torch.as_tensor(xref, device="cuda"))
    assert tuple(xtb_h.shape) == tuple(xtb_d.shape) == (P, N, H + 1, 14) and not cb_h.is_cuda
    np.testing.assert_array_equal(cb_h.numpy(), cb_d.cpu().numpy())
    np.testing.assert_array_equal(gb_h.numpy(), gb_d.cpu().numpy())
    np.testing.assert_array_equal(xtb_h.cpu().numpy(), xtb_d.cpu().numpy())


def test_apg_mpc_trace_matches_oracle(cuda):
    """One MPC solve: host APG (CUDA cost/grad) vs oracle APG (oracle cost/grad), same noise."""
    from sde4mbrl_b200 import apg, random as R
    N, H, iters = 8, 20, 6
    pm, mpc, nn, (_, multi_cost, vprox, _, construct_opt) = _mpc_setup(N, H, iters)
    cp, op = mpc["cost_params"], mpc["apg_mpc"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    rng = np.random.default_rng(2)
    y0 = hp.start_state("hexa", rng)
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    xref[:, :3] = [0.0, 0.5, 1.4]
    key = R.PRNGKey(21)
    cost_gpu = lambda o: multi_cost(nn, y0, o, key, stage, extra_cost_args=xref)[0]
    prox_gpu = lambda x, stepsize=None: vprox(x)
    x0 = construct_opt(u=np.full(6, 0.33, np.float32))
    s = apg.apg(apg.init_apg(x0, cost_gpu, op), cost_gpu, op, proximal_fn=prox_gpu)
    # oracle side
    dw = torch.tensor(tf.rollout_noise(tf.PRNGKey(21), N, H, 13), dtype=torch.float64)
    om = hp.oracle_model("hexa", pm, nn, torch.float64)
    ts = orc.compute_timesteps(pm)
    y64, xr64 = torch.tensor(y0, dtype=torch.float64), torch.tensor(xref, dtype=torch.float64)
    cost_cpu = lambda o: orc.compute_cost(om, y64, o, dw, hp.stage_cost_fn("hexa", cp), xr64)[0] + orc.u_slew_cost(o, ts, cp, 6)
    prox_cpu = lambda x, stepsize=None: torch.clamp(x, 1e-4, 1.0)
    so = apg_oracle.apg(apg_oracle.init_apg(x0.cpu().double(), cost_cpu, op), cost_cpu, op, proximal_fn=prox_cpu)
    assert s.num_steps == so.num_steps == iters + 1
    assert float(s.opt_cost) < float(s.init_cost)
    np.testing.assert_allclose(float(s.opt_cost), float(so.opt_cost), rtol=2e-4)
    np.testing.assert_allclose(float(s.stepsize), float(so.stepsize), rtol=1e-5)
    assert hp.rel_err(s.x_opt.cpu().numpy(), so.x_opt.numpy()) < 2e-3


def test_apg_with_host_tensors_matches_device_tensors(cuda):
    """apg.py driven with CPU tensors / numpy keys (every cost call goes through sde4_cost_grad_host and the
    box prox runs on the CPU) follows the same iterates as with CUDA tensors."""
    from sde4mbrl_b200 import apg, random as R
    N, H, iters = 8, 20, 5
    pm, mpc, nn, (_, multi_cost, vprox, _, construct_opt) = _mpc_setup(N, H, iters)
    cp, op = mpc["cost_params"], mpc["apg_mpc"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    y0 = hp.start_state("hexa", np.random.default_rng(2))
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    xref[:, :3] = [0.0, 0.5, 1.4]
    prox = lambda x, stepsize=None: vprox(x)
    x0 = construct_opt(u=np.full(6, 0.33, np.float32))
    cost_dev = lambda o: multi_cost(nn, y0, o, R.PRNGKey(21), stage, extra_cost_args=xref)[0]
    s_dev = apg.apg(apg.init_apg(x0, cost_dev, op), cost_dev, op, proximal_fn=prox)
    cost_host = lambda o: multi_cost(nn, y0, o, tf.PRNGKey(21), stage, extra_cost_args=xref)[0]
    s_host = apg.apg(apg.init_apg(x0.cpu(), cost_host, op), cost_host, op, proximal_fn=prox)
    assert not s_host.x_opt.is_cuda and s_dev.x_opt.is_cuda
    assert int(s_host.num_steps) == int(s_dev.num_steps)
    assert abs(float(s_host.opt_cost) - float(s_dev.opt_cost)) <= 1e-5 * abs(float(s_dev.opt_cost))
    assert hp.rel_err(s_host.x_opt.numpy(), s_dev.x_opt.cpu().numpy()) < 1e-4


@pytest.mark.parametrize("with_prox,host_args", [(True, False), (False, False), (True, True)])
def test_apg_with_bound_cost_runs_on_the_device_solver(cuda, with_prox, host_args):
    """apg.apg handed ``multi_cost.bind(...)`` instead of a lambda: the fresh line-search solve runs on the device-side
    solver (apg_batched.BatchedAPG) and must return the APGState the host loop returns for the same closure; a state in
    the middle of a solve falls back to the host loop, with value + gradient from the fused launch."""
    from sde4mbrl_b200 import apg, random as R
    N, H, iters = 8, 20, 6
    pm, mpc, nn, (_, multi_cost, vprox, _, construct_opt) = _mpc_setup(N, H, iters)
    cp, op = mpc["cost_params"], mpc["apg_mpc"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    y0 = hp.start_state("hexa", np.random.default_rng(2))
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    xref[:, :3] = [0.0, 0.5, 1.4]
    key = tf.PRNGKey(21) if host_args else R.PRNGKey(21)
    prox = (lambda x, stepsize=None: vprox(x)) if with_prox else None
    if with_prox:
        prox.box = vprox  # (what rotor_mpc.load_mpc_from_cfgfile attaches; vprox itself is accepted too)
    x0 = construct_opt(u=np.full(6, 0.33, np.float32))
    if host_args:
        x0 = x0.cpu()
    cost_l = lambda o: multi_cost(nn, y0, o, key, stage, extra_cost_args=xref)[0]
    s_host = apg.apg(apg.init_apg(x0, cost_l, op), cost_l, op, proximal_fn=(lambda x, stepsize=None: vprox(x)) if with_prox else None)
    cost_b = multi_cost.bind(nn, y0, key, stage, extra_cost_args=xref)
    assert abs(float(cost_b(x0)) - float(cost_l(x0))) <= 1e-6 * abs(float(cost_l(x0)))
    st0 = apg.init_apg(x0, cost_b, op)
    assert abs(float(st0.grad_sqr) - float(apg.init_apg(x0, cost_l, op).grad_sqr)) <= 1e-5 * float(st0.grad_sqr)
    s_dev = apg.apg(st0, cost_b, op, proximal_fn=prox)
    assert getattr(multi_cost, "_device_solvers", None), "the solve did not reach the device-side solver"
    assert s_dev.x_opt.is_cuda == (not host_args)
    assert int(s_dev.num_steps) == int(s_host.num_steps) == iters + 1
    for f in ("opt_cost", "curr_cost", "init_cost", "stepsize", "avg_stepsize", "momentum", "avg_momentum", "avg_linesearch", "grad_sqr"):
        np.testing.assert_allclose(float(getattr(s_dev, f)), float(getattr(s_host, f)), rtol=2e-4, err_msg=f)
    assert bool(s_dev.not_done) == bool(s_host.not_done)
    for f in ("x_opt", "xk", "yk", "grad_yk"):
        assert hp.rel_err(getattr(s_dev, f).cpu().numpy(), getattr(s_host, f).cpu().numpy()) < 2e-3, f
    # a state in the middle of a solve: host loop (fused value + gradient), same continuation as the lambda
    op2 = dict(op, max_iter=2)
    mid_b = apg.apg(s_dev._replace(not_done=True), cost_b, op2, proximal_fn=prox)
    mid_l = apg.apg(s_dev._replace(not_done=True), cost_l, op2, proximal_fn=prox)
    assert int(mid_b.num_steps) == int(mid_l.num_steps) == iters + 3
    np.testing.assert_allclose(float(mid_b.opt_cost), float(mid_l.opt_cost), rtol=1e-5)


def test_rejects_python_cost_callables(cuda):
    from sde4mbrl_b200 import random as R
    from sde4mbrl_b200._lib import Sde4Error
    pm, mpc, nn, (_, multi_cost, _, _, construct_opt) = _mpc_setup(4, 20)
    with pytest.raises(Sde4Error):
        multi_cost(nn, hp.start_state("hexa", np.random.default_rng(0)), construct_opt(), R.PRNGKey(0),
                   lambda x, u, e: 0.0, extra_cost_args=np.zeros((21, 13), np.float32))


def test_simulator_adapter(cuda):
    """cartpole_sde_gym._my_pred_fn semantics (mean / population std over particles of the next state)
    and the batched predict() over many start states, on the reference's shipped cartpole weights."""
    import os
    from sde4mbrl_b200 import io, random as R
    from sde4mbrl_b200.simulator import SDESimulator
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    nn, nominal = io.load_npz_fixture(os.path.join(gold, "ref_weights_cartpole_bb_learned_si.npz"))
    sim = SDESimulator(sde_models.CartPoleSDE, nominal, nn, num_particles=10, tau=0.02, horizon=1)
    x0, u0 = np.array([0.0, 0.0, np.pi, 0.0], np.float32), np.array([0.3], np.float32)
    mean, std = sim.pred_fn(x0, u0, R.PRNGKey(10))
    pm = dict(nominal, horizon=1, num_particles=10, stepsize=0.02)
    om = orc.make_model("cart_pole_sde", pm, nn)
    want, _ = orc.sample_rollouts(om, x0, u0[None], tf.PRNGKey(10), 10)
    np.testing.assert_allclose(mean.cpu().numpy(), want[:, -1].mean(0).numpy(), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(std.cpu().numpy(), want[:, -1].std(0, unbiased=False).numpy(), rtol=2e-3, atol=1e-6)
    # batched: 1000 start states, key per state derived on the device from one key
    rng = np.random.default_rng(0)
    P = 1000
    obs = rng.uniform(-1, 1, (P, 4)).astype(np.float32) * np.array([2, 3, np.pi, 6], np.float32)
    act = rng.uniform(-1, 1, (P, 1)).astype(np.float32)
    nxt = sim.predict(obs, act, R.PRNGKey(3))
    assert tuple(nxt.shape) == (P, 4)
    keys = tf.split(tf.PRNGKey(3), P)
    for p in (0, 17, 999):
        w, _ = orc.sample_rollouts(om, obs[p], act[p][None], keys[p], 10)
        np.testing.assert_allclose(nxt[p].cpu().numpy(), w[:, -1].mean(0).numpy(), rtol=2e-5, atol=2e-6)


def test_model_env_adapter(cuda):
    """mbrl ModelEnv contract around the simulator: step() chains keys like CartPoleSDEEnv.step_dynamics
    (cartpole_sde.py:243-253), evaluate_action_sequences() equals stepping the same particles one by one."""
    import os
    from sde4mbrl_b200 import io, random as R
    from sde4mbrl_b200.simulator import (SDEModelEnv, SDESimulator, cartpole_obs, cartpole_swingup_reward,
                                         cartpole_swingup_termination)
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    nn, nominal = io.load_npz_fixture(os.path.join(gold, "ref_weights_cartpole_bb_learned_si.npz"))
    sim = SDESimulator(sde_models.CartPoleSDE, nominal, nn, num_particles=5, tau=0.02, horizon=1)
    env = SDEModelEnv(sim, cartpole_swingup_termination, cartpole_swingup_reward, obs_fn=cartpole_obs, seed=10)
    rng = np.random.default_rng(1)
    P = 64
    s0 = rng.uniform(-1, 1, (P, 4)).astype(np.float32) * np.array([2, 3, np.pi, 6], np.float32)
    act = rng.uniform(-1, 1, (P, 1)).astype(np.float32)
    ms = env.reset(s0)
    obs, rew, done, ms2 = env.step(act, ms)
    assert tuple(obs.shape) == (P, 5) and tuple(rew.shape) == (P, 1) and tuple(done.shape) == (P, 1)
    # the step used split(PRNGKey(10))[1], one key per state split from it, mean over 5 particles
    k = tf.split(tf.PRNGKey(10))[1]
    keys = tf.split(k, P)
    pm = dict(nominal, horizon=1, num_particles=5, stepsize=0.02)
    om = orc.make_model("cart_pole_sde", pm, nn)
    for p in (0, 31, 63):
        w, _ = orc.sample_rollouts(om, s0[p], act[p][None], keys[p], 5)
        nxt = w[:, -1].mean(0).numpy()
        np.testing.assert_allclose(ms2["state"][p].cpu().numpy(), nxt, rtol=2e-5, atol=2e-6)
        # (sic) the reference takes the cosine of observation 3, which already is cos(theta)
        np.testing.assert_allclose(float(rew[p]), np.cos(np.cos(nxt[2])) - 0.1 * abs(nxt[0]), rtol=1e-4, atol=1e-5)
    assert not bool(done.any())
    far = s0.copy()
    far[0, 0] = 500.0
    _, _, d2, _ = env.step(act, env.reset(far))
    assert bool(d2[0]) and not bool(d2[1:].any())
    # candidate action sequences: one launch (B x particles x H) == per-particle trajectories rewarded step by step
    B, Hs, Np = 16, 12, 3
    seqs = rng.uniform(-1, 1, (B, Hs, 1)).astype(np.float32)
    env2 = SDEModelEnv(sim, cartpole_swingup_termination, cartpole_swingup_reward, obs_fn=cartpole_obs, seed=4)
    tot = env2.evaluate_action_sequences(seqs, s0[0], Np)
    assert tuple(tot.shape) == (B,)
    k2 = tf.split(tf.PRNGKey(4))[1]
    kb = tf.split(k2, B)
    pm2 = dict(nominal, horizon=Hs, num_particles=Np, stepsize=0.02)
    for kk in ("num_short_dt", "short_step_dt", "long_step_dt"):
        pm2.pop(kk, None)
    om2 = orc.make_model("cart_pole_sde", pm2, nn)
    for b in (0, 7, 15):
        w, _ = orc.sample_rollouts(om2, s0[0], seqs[b], kb[b], Np)
        st = w[:, 1:].numpy()
        r = (np.cos(np.cos(st[..., 2])) - 0.1 * np.abs(st[..., 0])).sum(1).mean()
        np.testing.assert_allclose(float(tot[b]), r, rtol=2e-4)


def test_particle_sharded_value_and_grad_single_rank(cuda):
    """``multi_sampling.particle_reducer`` (particles of one problem sharded over the process group; here a group of
    one): the packed host-buffer route and the device route give the plain ``value_and_grad`` result."""
    from sde4mbrl_b200 import nsde, random as R, costs
    from sde4mbrl_b200.dist import CostGradReducer
    import helpers as hp
    H, N = 12, 9
    mpc = configs.hexa_mpc(horizon=H, dt=0.05, num_particles=N)
    pm = configs.hexa_model(horizon=H, num_particles=N, stepsize=0.05)
    _, ms, _, _, _ = nsde.create_online_cost_sampling_fn(pm, mpc, sde_constr=sde_models.SDERotorModel)
    nn = configs.random_params(sde_models.SDERotorModel(pm), seed=3)
    rng = np.random.default_rng(2)
    y0, opt = hp.start_state("hexa", rng), hp.controls("hexa", H, rng)
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    cp = mpc["cost_params"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    key = np.array([0, 77], np.uint32)
    (c0, xt0), g0 = ms.value_and_grad(nn, y0, opt, key, stage, extra_cost_args=xref)
    ms.particle_reducer = CostGradReducer(H * ms.n_opt + 1, torch.device("cuda:0"), grad_shape=(H, ms.n_opt))
    try:
        (c1, xt1), g1 = ms.value_and_grad(nn, y0, opt, key, stage, extra_cost_args=xref)  # host arrays
        assert not c1.is_cuda and not g1.is_cuda
        (c2, _), g2 = ms.value_and_grad(nn, torch.as_tensor(y0).cuda(), torch.as_tensor(opt).cuda(), R.PRNGKey(77), stage,
                                        extra_cost_args=torch.as_tensor(xref).cuda())
    finally:
        ms.particle_reducer = None
    for c, g in ((c1, g1), (c2, g2)):
        assert abs(float(c) - float(c0)) <= 1e-6 * abs(float(c0))
        np.testing.assert_allclose(g.cpu().numpy(), np.asarray(g0), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(xt1.cpu().numpy(), xt0.cpu().numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("kind", ["cartpole", "hexa"])
def test_controls_as_a_function_of_the_observation(cuda, kind):
    """``us`` callable (sde_solver.py:292): closed-loop rollout, one kernel launch per step, noise from the reference's
    key chain; checked against the oracle stepped with the same feedback law."""
    from sde4mbrl_b200 import nsde, random as R
    H, N = 15, 11
    mk, cls, oname = hp.KINDS[kind]
    pm = mk(horizon=H, num_particles=N)
    nn = configs.random_params(cls(pm), seed=2)
    _, ms = nsde.create_sampling_fn(pm, sde_constr=cls, seed=0)
    rng = np.random.default_rng(4)
    y0 = hp.start_state(kind, rng)
    n_x, n_u = pm["n_x"], pm["n_u"]
    K = rng.normal(0, 0.05, (n_x, n_u)).astype(np.float32)
    c = (np.full(n_u, 0.33) if kind == "hexa" else np.zeros(n_u)).astype(np.float32)
    Kt, ct = torch.as_tensor(K).cuda(), torch.as_tensor(c).cuda()
    policy = lambda yy: torch.clamp(yy @ Kt + ct, -1.0, 1.0)
    xevol, yevol, uevol = ms(nn, y0, policy, R.PRNGKey(9))
    assert tuple(xevol.shape) == (N, H + 1, n_x) and tuple(uevol.shape) == (N, H, n_u)
    om = orc.make_model(oname, pm, {m: {k: torch.tensor(np.asarray(v)) for k, v in d.items()} for m, d in nn.items()})
    dw = torch.tensor(tf.rollout_noise(tf.PRNGKey(9), N, H, n_x))
    ts = orc.compute_timesteps(pm)
    x = torch.tensor(y0)[None].expand(N, n_x)
    Kc, cc = torch.tensor(K), torch.tensor(c)
    xs = [x]
    for t in range(H):
        u = torch.clamp(x @ Kc + cc, -1.0, 1.0)
        x = orc.euler_maruyama(om, ts[t:t + 1], x, u[:, None, :], dw[:, t:t + 1])[:, 1]
        xs.append(x)
    want = torch.stack(xs, dim=1).numpy()
    assert hp.rel_err(xevol.cpu().numpy(), want) < 5e-5
    np.testing.assert_allclose(uevol.cpu().numpy()[:, 0], np.clip(want[:, 0] @ K + c, -1, 1), rtol=1e-5, atol=1e-6)


def test_fused_simulator_outputs(cuda):
    """sde4_sim_rollout (no trajectory write-out; particle mean / population std, per-step reward with termination and the
    reward / done of the mean state computed on the device) against the same quantities formed from the full
    trajectories of sde4_rollout -- incl. samples that terminate in the middle of the sequence."""
    import os
    from sde4mbrl_b200 import io, random as R
    from sde4mbrl_b200.simulator import (SDEModelEnv, SDESimulator, cartpole_obs, cartpole_swingup_reward,
                                         cartpole_swingup_termination)
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    nn, nominal = io.load_npz_fixture(os.path.join(gold, "ref_weights_cartpole_bb_learned_si.npz"))
    P, N, H = 37, 6, 15
    sim = SDESimulator(sde_models.CartPoleSDE, nominal, nn, num_particles=N, tau=0.02, horizon=H)
    rng = np.random.default_rng(5)
    obs = rng.uniform(-1, 1, (P, 4)).astype(np.float32) * np.array([2, 3, np.pi, 6], np.float32)
    obs[:8, 0] = 119.0 + 0.1 * np.arange(8)    # these cross |x| = 120 after a few steps (0.12 per step) ...
    obs[:8, 1] = 6.0
    obs[8, 0] = 500.0                          # ... this one is terminated from the first step on
    act = rng.uniform(-1, 1, (P, H, 1)).astype(np.float32)
    key = R.PRNGKey(9)
    traj = sim.predict(obs, act, key, return_particles=True)              # [P, N, H+1, 4] through sde4_rollout
    r = sim.step_summary(obs, act, key, reward="cartpole_swingup")
    last = traj[:, :, -1, :]
    np.testing.assert_allclose(r["mean"].cpu().numpy(), last.mean(1).cpu().numpy(), rtol=1e-6, atol=1e-6)
    r2 = sim.engine.sim_rollout(sim.pk, obs, act, sim.ts, R.split(key, P), N=N, want_std=True)
    np.testing.assert_allclose(r2["std"].cpu().numpy(), last.std(1, unbiased=False).cpu().numpy(), rtol=2e-4, atol=1e-6)
    assert r2["xevol"] is None  # nothing of the trajectory was written
    # per-particle reward summed up to (and including) the terminal step, averaged over the particles
    o = cartpole_obs(traj[:, :, 1:, :]).reshape(P * N * H, -1)
    a = torch.as_tensor(act, device="cuda")[:, None].expand(P, N, H, 1).reshape(P * N * H, 1)
    rew = cartpole_swingup_reward(a, o).view(P, N, H)
    done = cartpole_swingup_termination(a, o).view(P, N, H)
    alive = (torch.cumsum(done.int(), dim=2) - done.int()) == 0
    want = (rew * alive).sum(2).mean(1)
    np.testing.assert_allclose(r["reward_mean"].cpu().numpy(), want.cpu().numpy(), rtol=2e-5, atol=1e-5)
    assert bool(done[:8].any()) and not bool(done[:8, :, 0].all()) and bool(done[8].all())
    # reward / termination of the MEAN next state (ModelEnv.step)
    om = cartpole_obs(last.mean(1))
    np.testing.assert_allclose(r["reward_of_mean"].cpu().numpy(), cartpole_swingup_reward(a[:P], om).view(-1).cpu().numpy(),
                               rtol=2e-5, atol=1e-5)
    assert torch.equal(r["done_of_mean"].bool().view(-1, 1), cartpole_swingup_termination(a[:P], om))
    # the ModelEnv wrapper picks the fused path for this reward / termination / observation triple and an unfused one else
    env = SDEModelEnv(sim, cartpole_swingup_termination, cartpole_swingup_reward, obs_fn=cartpole_obs, seed=3)
    env_u = SDEModelEnv(sim, lambda a_, o_: cartpole_swingup_termination(a_, o_), cartpole_swingup_reward, obs_fn=cartpole_obs, seed=3)
    assert env._fused_kind == "cartpole_swingup" and env_u._fused_kind is None
    sim1 = SDESimulator(sde_models.CartPoleSDE, nominal, nn, num_particles=N, tau=0.02, horizon=1)
    e1 = SDEModelEnv(sim1, cartpole_swingup_termination, cartpole_swingup_reward, obs_fn=cartpole_obs, seed=3)
    e2 = SDEModelEnv(sim1, lambda a_, o_: cartpole_swingup_termination(a_, o_), cartpole_swingup_reward, obs_fn=cartpole_obs, seed=3)
    o1, rw1, d1, _ = e1.step(act[:, 0], e1.reset(obs))
    o2, rw2, d2, _ = e2.step(act[:, 0], e2.reset(obs))
    np.testing.assert_allclose(o1.cpu().numpy(), o2.cpu().numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(rw1.cpu().numpy(), rw2.cpu().numpy(), rtol=2e-5, atol=1e-5)
    assert torch.equal(d1, d2)

"""BASELINE.json's FULL sizes: parity against the C oracle where it finishes in seconds, and
size-independent properties elsewhere (determinism, shard additivity, value/trajectory consistency,
unit quaternions, finite-difference directional derivative with common random numbers)."""
import contextlib
import io as _io

import numpy as np
import pytest
import torch

import helpers as hp
from oracle import c_oracle, packer
from oracle import nsde_oracle as orc
from oracle import threefry as tf
from sde4mbrl_b200 import configs, costs, sde_models

pytestmark = pytest.mark.gpu


def _c3():
    N, H = 4096, 100
    pm = configs.hexa_model(horizon=H, num_particles=N, stepsize=0.01)
    model = sde_models.SDERotorModel(pm)
    nn = configs.random_params(model, seed=0)
    cp = configs.hexa_mpc(horizon=H, dt=0.01, num_particles=N)["cost_params"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    rng = np.random.default_rng(0)
    y0 = hp.start_state("hexa", rng)
    opt = hp.controls("hexa", H, rng)
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    xref[:, :3] = [0.0, 0.5, 1.4]
    return N, H, pm, model, nn, cp, stage, y0, opt, xref


def test_c3_full_size_against_c_oracle(cuda):
    """configs[2]: 4096 particles x 100 steps, in-kernel threefry noise, value + gradient, against
    oracle/c (itself pinned to the torch oracle by tests/test_oracle_models.py)."""
    from sde4mbrl_b200.engine import Engine
    N, H, pm, model, nn, cp, stage, y0, opt, xref = _c3()
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    key = tf.PRNGKey(2050)
    c, g, xt = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, key=key, N=N, want_xtraj=True)
    c, g, xt = float(c[0]), g.cpu().numpy(), xt.cpu().numpy()
    # descriptors and parameter block of the C oracle come from oracle/packer.py (straight from the params dict and the
    # haiku tree), NOT from the product's describe() / pack(): a packer error is not common mode here
    odesc = packer.rotor_desc(pm)
    cw, gw, xw = c_oracle.rotor_cost_grad(odesc, packer.rotor_track_cost_desc(cp, 6), packer.rotor_pack(pm, nn, odesc), y0,
                                          opt, ts, xref, N, key=key, want_xtraj=True)
    assert abs(c - cw) <= 2e-5 * abs(cw)
    assert hp.rel_err(g, gw) < 1e-3
    got = np.transpose(xt[:, :13, :], (2, 0, 1))  # [N, H+1, 13]
    assert hp.rel_err(got, xw) < 2e-4
    # properties: unit quaternions everywhere; the returned mean equals the trapezoid of the stage-cost column
    qn = np.linalg.norm(got[:, :, 6:10], axis=-1)
    assert np.abs(qn - 1.0).max() < 2e-6
    sc = xt[:, 13, :].astype(np.float64)  # [H+1, N]
    trap = (0.5 * ts.astype(np.float64)[:, None] * (sc[:-1] + sc[1:])).sum(0).mean()
    slew = float(orc.u_slew_cost(torch.tensor(opt, dtype=torch.float64), ts, cp, 6))
    assert abs((trap + slew) - c) <= 2e-6 * abs(c)


def test_c3_full_size_properties(cuda):
    from sde4mbrl_b200.engine import Engine
    N, H, pm, model, nn, cp, stage, y0, opt, xref = _c3()
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    key = tf.PRNGKey(7)
    run = lambda o, **k: eng.cost_grad(nn, y0[None], o, ts, xref, stage.desc, key=key, N=k.pop("N", N), **k)
    c1, g1, _ = run(opt)
    c1, g1 = c1.clone(), g1.clone()
    c2, g2, _ = run(opt)
    assert torch.equal(c1, c2) and torch.equal(g1, g2)  # deterministic reduction, bit for bit
    # shards of 1024 particles each (what 4 GPUs compute) add up to the whole
    cs, gs = torch.zeros_like(c1), torch.zeros_like(g1)
    for r in range(4):
        cr, gr, _ = run(opt, N=1024, particle_offset=1024 * r, particle_total=N)
        cs += cr
        gs += gr
    assert abs(float(cs[0]) - float(c1[0])) <= 1e-6 * abs(float(c1[0]))
    assert hp.rel_err(gs.cpu().numpy(), g1.cpu().numpy()) < 1e-5
    # one thread per sample and 8 lanes per sample agree
    c8, g8, _ = run(opt, lanes=8)
    cL1, gL1, _ = run(opt, lanes=1)
    assert abs(float(c8[0]) - float(cL1[0])) <= 2e-5 * abs(float(cL1[0]))
    assert hp.rel_err(g8.cpu().numpy(), gL1.cpu().numpy()) < 2e-4
    # directional derivative with common random numbers (value-only launches)
    d = np.random.default_rng(1).standard_normal(opt.shape).astype(np.float32)
    d /= np.linalg.norm(d)
    eps = 2e-2
    cp_, _, _ = run(opt + eps * d, want_grad=False)
    cm_, _, _ = run(opt - eps * d, want_grad=False)
    fd = (float(cp_[0]) - float(cm_[0])) / (2 * eps)
    an = float((g1.cpu().numpy() * d).sum())
    assert abs(fd - an) <= 2e-2 * max(abs(an), 1e-3)


def test_c1_msd_full_size_rollout(cuda):
    """configs[0]: mass-spring-damper, 200 particles x 100 steps, against the torch oracle."""
    from sde4mbrl_b200.engine import Engine
    pm, model, nn = hp.build("msd", 100, 200, seed=0)
    rng = np.random.default_rng(0)
    y0, u = hp.start_state("msd", rng), hp.controls("msd", 100, rng)
    key = tf.PRNGKey(0)
    want, _ = orc.sample_rollouts(hp.oracle_model("msd", pm, nn), y0, u, key, 200)
    x = Engine(model).rollout(nn, y0[None], u, orc.compute_timesteps(pm), key=key, N=200)
    assert hp.rel_err(hp.from_native_x(x, 2), want.numpy()) < 1e-4


def test_c5a_million_start_states(cuda):
    """configs[4]: 2^20 cartpole start states x horizon 30 (one particle each).  Deterministic, finite,
    and a random sample of the problems matches the oracle run on those problems alone."""
    from sde4mbrl_b200.engine import Engine
    P, H = 1 << 20, 30
    pm, model, nn = hp.build("cartpole", H, 1, seed=0)
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    g = torch.Generator(device="cuda").manual_seed(0)
    y0 = (torch.rand((P, 4), device="cuda", generator=g) * 2 - 1) * torch.tensor([2.0, 3.0, np.pi, 6.0], device="cuda")
    u = torch.rand((P, H, 1), device="cuda", generator=g) * 2 - 1
    keys = torch.stack([torch.zeros(P, dtype=torch.int32, device="cuda"),
                        torch.arange(P, dtype=torch.int32, device="cuda")], dim=1).contiguous()
    x = eng.rollout(nn, y0, u, ts, key=keys, N=1)
    assert tuple(x.shape) == (H + 1, 4, P) and bool(torch.isfinite(x).all())
    x2 = eng.rollout(nn, y0, u, ts, key=keys, N=1)
    assert torch.equal(x, x2)
    om = hp.oracle_model("cartpole", pm, nn)
    pick = np.random.default_rng(3).choice(P, 24, replace=False)
    for p in pick:
        want, _ = orc.sample_rollouts(om, y0[p].cpu().numpy(), u[p].cpu().numpy(), np.array([0, p], np.uint32), 1)
        assert hp.rel_err(x[:, :, p].cpu().numpy(), want.numpy()[0]) < 2e-5


def test_c4_batch_of_4096_solves(cuda):
    """configs[3]: 4096 independent hexacopter MPC solves (H = 20, N = 1, 20 APG iterations) in lock step;
    three of them are re-solved alone with the host APG and must reach the same optimum."""
    from sde4mbrl_b200 import apg, nsde
    from sde4mbrl_b200.apg_batched import BatchedAPG
    P, H, iters = 4096, 20, 20
    pm = configs.hexa_model(horizon=H, num_particles=1, stepsize=0.05)
    mpc = configs.hexa_mpc(horizon=H, dt=0.05, num_particles=1, max_iter=iters)
    with contextlib.redirect_stdout(_io.StringIO()):
        _, multi_cost, vprox, _, construct_opt = nsde.create_online_cost_sampling_fn(pm, mpc, sde_constr=sde_models.SDERotorModel)
    nn = configs.random_params(sde_models.SDERotorModel(pm), seed=0)
    cp, op = mpc["cost_params"], mpc["apg_mpc"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    rng = np.random.default_rng(5)
    y0 = np.tile(configs.hover_state(), (P, 1)).astype(np.float32)
    y0[:, 3:6] += rng.normal(0, 0.05, (P, 3)).astype(np.float32)
    y0[:, 10:13] += rng.normal(0, 0.05, (P, 3)).astype(np.float32)
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    xref[:, :3] = [0.0, 0.5, 1.4]
    keys = np.stack([np.zeros(P, np.uint32), np.arange(P, dtype=np.uint32)], axis=1)
    x0 = construct_opt(u=np.full(6, 0.33, np.float32))
    solver = BatchedAPG(multi_cost.engine, nn, stage, op, multi_cost.time_steps, 1, lb=1e-4, ub=1.0)
    solver.init(x0, y0, xref, keys).run(iters)
    opt_cost = solver.t["opt_cost"].cpu().numpy()
    init_cost = solver.t["init_cost"].cpu().numpy()
    assert np.isfinite(opt_cost).all() and (opt_cost <= init_cost + 1e-7).all() and (opt_cost < init_cost).mean() > 0.9
    xb = solver.x_opt().cpu().numpy()
    assert xb.min() >= 1e-4 - 1e-9 and xb.max() <= 1.0 + 1e-9  # box prox
    prox = lambda x, stepsize=None: vprox(x)
    for p in (0, 1777, 4095):
        cost_p = lambda o, p=p: multi_cost(nn, y0[p], o, keys[p], stage, extra_cost_args=xref)[0]
        s = apg.apg(apg.init_apg(x0.cpu(), cost_p, op), cost_p, op, proximal_fn=prox)
        assert abs(float(s.opt_cost) - float(opt_cost[p])) <= 2e-4 * abs(float(s.opt_cost))
        assert hp.rel_err(xb[p], s.x_opt.numpy()) < 5e-3


def test_c2_cartpole_mpc_solve_device_vs_host_apg(cuda):
    """configs[1]: one cartpole swing-up MPC problem, 1024 particles x horizon 50, 20 APG iterations: the device-side
    solver (one CUDA-graph replay per iteration) and the reference-facing host `apg.py` loop driving the same fused
    cost kernel reach the same optimum, and it improves on the initial cost."""
    from sde4mbrl_b200 import apg, nsde
    from sde4mbrl_b200.apg_batched import BatchedAPG
    H, N, iters = 50, 1024, 20
    pm = configs.cartpole_model(horizon=H, num_particles=N)
    mpc = configs.cartpole_mpc(horizon=H, dt=0.02, num_particles=N, max_iter=iters)
    mpc["apg_mpc"].update(atol=0.0, rtol=0.0, max_no_improvement_iter=iters)
    with contextlib.redirect_stdout(_io.StringIO()):
        _, mcost, vprox, _, _ = nsde.create_online_cost_sampling_fn(pm, mpc, sde_constr=sde_models.CartPoleSDE)
    nn = configs.random_params(mcost.engine.model, seed=0)
    for k_, v_ in nn.items():  # dynamics that react to the control (random_params damps the residual nets)
        if k_.endswith("linear_2") and "init_residual_networks" in k_:
            v_["w"], v_["b"] = v_["w"] * np.float32(25.0), v_["b"] * np.float32(25.0)
    cp = mpc["cost_params"]
    stage = costs.QuadraticCost(cp["xerr"], cp["uerr"], cp["uref"], feat="cartpole", res_mult=cp["res_mult"])
    y0 = np.array([0.0, 0.0, np.pi, 0.0], np.float32) + np.random.default_rng(3).normal(0, 0.01, 4).astype(np.float32)
    xref = np.zeros((H + 1, 4), np.float32)
    key = np.array([0, 7], np.uint32)
    x0 = np.full((H, 1), 1e-4, np.float32)
    solver = BatchedAPG(mcost.engine, nn, stage, mpc["apg_mpc"], mcost.time_steps, N, lb=-1.0, ub=1.0)
    solver.init(torch.as_tensor(x0, device="cuda"), y0[None], xref, key).run(iters)
    sb = solver.state_of(0)
    cost_h = lambda o: mcost(nn, y0, o, key, stage, extra_cost_args=xref)[0]
    sh = apg.apg(apg.init_apg(torch.as_tensor(x0), cost_h, mpc["apg_mpc"]), cost_h, mpc["apg_mpc"],
                 proximal_fn=lambda x, stepsize=None: vprox(x))
    assert sb.num_steps == sh.num_steps == iters + 1
    assert float(sb.opt_cost) < float(sb.init_cost)
    assert abs(float(sb.opt_cost) - float(sh.opt_cost)) <= 1e-4 * abs(float(sh.opt_cost))
    assert hp.rel_err(sb.x_opt.cpu().numpy(), sh.x_opt.cpu().numpy()) < 2e-3
    assert float(sb.x_opt.min()) >= -1.0 and float(sb.x_opt.max()) <= 1.0

"""K2/K3 parity: fused cost + reverse-mode kernel vs the oracle (float32 values, float64 autograd
gradients) on identical noise.

Tolerances: cost 2e-5 relative; gradient 1e-3 relative to its max-norm (the float32 adjoint is
compared with a float64 gradient of the same function, SURVEY section 7 hard part 6)."""
import numpy as np
import pytest
import torch

import helpers as hp
from oracle import nsde_oracle as orc
from sde4mbrl_b200 import configs, costs

pytestmark = pytest.mark.gpu


def _setup(kind, H, N, seed, scale="fan_in"):
    rng = np.random.default_rng(seed)
    pm, model, nn = hp.build(kind, H, N, seed=seed, scale=scale)
    y0 = hp.start_state(kind, rng)
    opt = hp.controls(kind, H, rng)
    if kind == "msd":
        opt = rng.uniform(-0.5, 0.5, (H, 1)).astype(np.float32)
    dw = rng.standard_normal((N, H, model.n_x)).astype(np.float32)
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
    return pm, model, nn, y0, opt, dw, cp, stage, xref.astype(np.float32)


def _oracle(kind, pm, nn, y0, opt, dw, cp, xref, dtype, slew):
    om = hp.oracle_model(kind, pm, nn, dtype)
    o = torch.tensor(opt, dtype=dtype, requires_grad=True)
    c, xtraj = orc.compute_cost(om, torch.tensor(y0, dtype=dtype), o, torch.tensor(dw, dtype=dtype),
                                hp.stage_cost_fn(kind, cp), torch.tensor(xref, dtype=dtype))
    if slew:
        c = c + orc.u_slew_cost(o, orc.compute_timesteps(pm), cp, om.n_u)
    (g,) = torch.autograd.grad(c, o)
    return float(c), g.numpy(), xtraj.detach().numpy()


# scale = "yaml": the reference's initialiser ranges (weights ~1e-3 .. 1e-1) -- where the MUFU tanh / sigmoid have their
# largest RELATIVE error in the hidden activations; same tolerances (cost 2e-5, gradient 1e-3 of its maximum)
@pytest.mark.parametrize("scale", ["fan_in", "yaml"])
@pytest.mark.parametrize("kind,H,N", [("msd", 100, 64), ("msd", 20, 5), ("cartpole", 50, 64), ("cartpole_bb", 30, 7),
                                      ("hexa", 100, 32), ("hexa", 20, 1), ("hexa16", 20, 33)])
def test_cost_and_grad_given_noise(cuda, kind, H, N, scale):
    from sde4mbrl_b200.engine import Engine
    pm, model, nn, y0, opt, dw, cp, stage, xref = _setup(kind, H, N, seed=11, scale=scale)
    slew = kind.startswith("hexa")
    c32, _, xt32 = _oracle(kind, pm, nn, y0, opt, dw, cp, xref, torch.float32, slew)
    c64, g64, _ = _oracle(kind, pm, nn, y0, opt, dw, cp, xref, torch.float64, slew)
    eng = Engine(model)
    cost, grad, xtraj = eng.cost_grad(nn, y0[None], opt, orc.compute_timesteps(pm), xref, stage.desc, N=N,
                                      noise=hp.to_native_noise(dw), want_grad=True, want_xtraj=True)
    assert abs(float(cost[0]) - c64) <= 2e-5 * abs(c64) + 1e-7
    got_traj = xtraj.permute(2, 0, 1).cpu().numpy()
    assert hp.rel_err(got_traj[..., :model.n_x], xt32[..., :model.n_x]) < (1e-4 if H > 50 else 2e-5)
    assert hp.rel_err(got_traj[..., -1], xt32[..., -1]) < 1e-4  # stage cost column (nsde.py:1525)
    assert hp.rel_err(grad.cpu().numpy(), g64) < 1e-3
    # value-only launch gives the same value
    cost2, g2, _ = eng.cost_grad(nn, y0[None], opt, orc.compute_timesteps(pm), xref, stage.desc, N=N,
                                 noise=hp.to_native_noise(dw), want_grad=False)
    # (the fused value+grad kernel accumulates the trapezoid in its backward sweep, the value-only one in
    # the reference's forward order: same terms, different rounding)
    assert g2 is None and abs(float(cost2[0]) - float(cost[0])) <= 2e-6 * abs(float(cost[0]))


def test_grad_matches_finite_differences(cuda):
    """Independent of the oracle: central differences of the kernel's own cost (threefry noise =
    common random numbers, so the cost is a deterministic function of opt)."""
    from sde4mbrl_b200.engine import Engine
    from oracle import threefry as tf
    pm, model, nn, y0, opt, dw, cp, stage, xref = _setup("hexa", 20, 32, seed=5)
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    key = tf.PRNGKey(17)
    _, grad, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, key=key, N=32)
    grad = grad.cpu().numpy()
    rng = np.random.default_rng(0)
    for _ in range(6):
        t, j = rng.integers(0, 20), rng.integers(0, 6)
        eps = 2e-3
        op, om_ = opt.copy(), opt.copy()
        op[t, j] += eps
        om_[t, j] -= eps
        cp_, _, _ = eng.cost_grad(nn, y0[None], op, ts, xref, stage.desc, key=key, N=32, want_grad=False)
        cm_, _, _ = eng.cost_grad(nn, y0[None], om_, ts, xref, stage.desc, key=key, N=32, want_grad=False)
        fd = (float(cp_[0]) - float(cm_[0])) / (2 * eps)
        assert abs(fd - grad[t, j]) < 2e-2 * np.max(np.abs(grad)) + 1e-5


def test_batched_problems_and_strides(cuda):
    """P problems, problem-minor opt layout (as the batched device APG uses), one key per problem."""
    from sde4mbrl_b200.engine import Engine
    from oracle import threefry as tf
    P, H, N = 5, 20, 2
    pm, model, nn, y0, opt, dw, cp, stage, xref = _setup("hexa", H, N, seed=9)
    rng = np.random.default_rng(3)
    y0s = np.stack([hp.start_state("hexa", rng) for _ in range(P)])
    opts = np.stack([hp.controls("hexa", H, rng) for _ in range(P)])
    keys = tf.split(tf.PRNGKey(4), P)
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    opt_pm = torch.tensor(opts).cuda().permute(1, 2, 0).contiguous().permute(2, 0, 1)  # [P,H,n] view, P minor
    cost, grad, _ = eng.cost_grad(nn, y0s, opt_pm, ts, xref, stage.desc, key=keys, N=N)
    assert grad.stride() == opt_pm.stride()
    for p in range(P):
        dwp = tf.rollout_noise(keys[p], N, H, 13)
        c64, g64, _ = _oracle("hexa", pm, nn, y0s[p], opts[p], dwp, cp, xref, torch.float64, True)
        assert abs(float(cost[p]) - c64) <= 2e-5 * abs(c64)
        assert hp.rel_err(grad[p].cpu().numpy(), g64) < 1e-3


def test_state_constraints_withprox_and_noprox(cuda):
    """nsde.py:1400-1419: slack (proximal) and one-sided penalties, gradient wrt slack columns."""
    from sde4mbrl_b200.engine import Engine
    from sde4mbrl_b200 import _lib as L
    H, N = 20, 4
    pm, model, nn, y0, opt, dw, cp, stage, xref = _setup("cartpole", H, N, seed=13)
    ts = orc.compute_timesteps(pm)
    eng = Engine(model)
    for slack_prox in (True, False):
        mpc = {"state_constr": {"state_id": [0, 3], "state_bound": [[-0.05, 0.05], [-1.0, 0.5]],
                                "state_penalty": [3.0, 0.7], "slack_proximal": slack_prox, "constr_pen": 2.0,
                                "slack_scaling": [1.0, 2.0]}}
        _, sc = orc.constraint_setup(4, 1, mpc)
        ns = 2 if slack_prox else 0
        constr = dict(mode=L.CONSTR_WITHPROX if slack_prox else L.CONSTR_NOPROX, n_slack=ns, weight=sc["weight"],
                      state_idx=sc["state_idx"], penalty=sc["penalty"], state_lb=sc["state_lb"], state_ub=sc["state_ub"],
                      slack_scaling=np.broadcast_to(sc["slack_scaling"], (2,)))
        st = costs.with_constraints(stage, constr, 1)
        rng = np.random.default_rng(1)
        o = np.concatenate([opt, rng.normal(0, 0.3, (H, ns)).astype(np.float32)], axis=1)
        om = hp.oracle_model("cartpole", pm, nn, torch.float64)
        ot = torch.tensor(o, dtype=torch.float64, requires_grad=True)
        c, _ = orc.compute_cost(om, torch.tensor(y0, dtype=torch.float64), ot, torch.tensor(dw, dtype=torch.float64),
                                hp.stage_cost_fn("cartpole", cp), torch.tensor(xref, dtype=torch.float64), sc=sc)
        (g64,) = torch.autograd.grad(c, ot)
        cost, grad, _ = eng.cost_grad(nn, y0[None], o, ts, xref, st.desc, N=N, noise=hp.to_native_noise(dw))
        assert abs(float(cost[0]) - float(c)) <= 3e-5 * abs(float(c))
        assert hp.rel_err(grad.cpu().numpy(), g64.numpy()) < 1e-3


@pytest.mark.parametrize("lanes", [1, 8])
def test_particle_sharding(cuda, lanes):
    """One problem's particles split over two 'ranks' (particle_offset / particle_total): the SUM of
    the shard results equals the unsharded mean, with in-kernel threefry keys split(rng, N_total)."""
    from sde4mbrl_b200.engine import Engine
    from oracle import threefry as tf
    N, H = 64, 20
    pm, model, nn, y0, opt, dw, cp, stage, xref = _setup("hexa", H, N, seed=21)
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    key = tf.PRNGKey(6)
    c, g, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, key=key, N=N, lanes=lanes)
    c, g = c.clone(), g.clone()
    ca, ga, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, key=key, N=40, particle_offset=0,
                              particle_total=N, lanes=lanes)
    ca, ga = ca.clone(), ga.clone()
    cb, gb, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, key=key, N=24, particle_offset=40,
                              particle_total=N, lanes=lanes)
    assert abs(float(ca[0] + cb[0]) - float(c[0])) <= 1e-6 * abs(float(c[0]))  # reduction order only
    assert hp.rel_err((ga + gb).cpu().numpy(), g.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("variant", ["group_sigmoids", "res_sig", "both"])
@pytest.mark.parametrize("lanes,use_key", [(1, False), (8, False), (8, True), (4, True)])
def test_shaped_rotor_cost(cuda, variant, lanes, use_key):
    """The optional shaping of one_step_cost_f (sde_mpc_design.py:84-130): per-group sigmoid(e*w - 7)*w_sig and the
    outer tanh(tot - lo)*mult, through the generic kernels (given noise), and the lane-split MODE-1 kernels
    (in-kernel threefry + shaped cost, which cannot use the branch-free plain-cost instantiation)."""
    from sde4mbrl_b200.engine import Engine
    from oracle import threefry as tf
    N, H = 24, 20
    pm, model, nn, y0, opt, dw, _, _, xref = _setup("hexa", H, N, seed=5)
    cp = {"uref": [0.33] * 6, "uerr": 300.0, "perr": [30.0, 30.0, 60.0], "verr": [40.0, 40.0, 40.0], "qerr": [50.0, 50.0, 500.0],
          "werr": [30.0, 30.0, 30.0], "res_mult": 0.01, "u_slew_coeff": 0.1}
    if variant in ("group_sigmoids", "both"):
        cp.update({"perr_sig": [2.0, 2.0, 4.0], "verr_sig": 1.5, "qerr_sig": [1.0, 1.0, 3.0], "werr_sig": 0.5, "uerr_sig": 0.7})
    if variant in ("res_sig", "both"):
        cp["res_sig"] = [3.0, 2.5] if variant == "res_sig" else [1.0, 2.5]
        if variant == "res_sig":
            for k in ("perr", "verr", "qerr", "werr"):
                cp[k] = [0.1 * v for v in cp[k]]
            cp["uerr"] = 30.0
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    assert (stage.desc.sig_mask != 0) == (variant != "res_sig") and bool(stage.desc.use_res_sig) == (variant != "group_sigmoids")
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    if use_key:
        key = tf.PRNGKey(17)
        dw = tf.rollout_noise(key, N, H, 13)
        cost, grad, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, key=key, N=N, lanes=lanes)
    else:
        cost, grad, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, N=N, noise=hp.to_native_noise(dw), lanes=lanes)
    c64, g64, _ = _oracle("hexa", pm, nn, y0, opt, dw, cp, xref, torch.float64, True)
    assert abs(float(cost[0]) - c64) <= 3e-5 * abs(c64) + 1e-7
    assert hp.rel_err(grad.cpu().numpy(), g64) < 2e-3
    assert np.abs(g64).max() > 1e-4  # the shaped cost is not saturated: the gradient test means something


@pytest.mark.parametrize("lanes,use_key", [(1, False), (8, False), (8, True)])
def test_terminal_cost(cuda, lanes, use_key):
    """Terminal cost of compute_cost_function (nsde.py:1517-1521) in the form load_mpc_from_cfgfile builds from
    `cost_params_end` (sde_mpc_design.py:353-361): the stage-cost expression without control terms, at x_H."""
    from sde4mbrl_b200.engine import Engine
    from oracle import threefry as tf
    N, H = 20, 20
    pm, model, nn, y0, opt, dw, cp, stage, xref = _setup("hexa", H, N, seed=9)
    cpe = {"perr": [80.0, 80.0, 120.0], "verr": [10.0, 10.0, 10.0], "qerr": [5.0, 5.0, 50.0], "werr": 2.0, "res_mult": 0.05}
    term = costs.one_step_cost_f(cpe, 6)
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    if use_key:
        key = tf.PRNGKey(23)
        dw = tf.rollout_noise(key, N, H, 13)
        kw = dict(key=key)
    else:
        kw = dict(noise=hp.to_native_noise(dw))
    cost, grad, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, terminal=term.desc, N=N, lanes=lanes, **kw)
    cost0, grad0, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, N=N, lanes=lanes, **kw)
    om = hp.oracle_model("hexa", pm, nn, torch.float64)
    o = torch.tensor(opt, dtype=torch.float64, requires_grad=True)
    xr64 = torch.tensor(xref, dtype=torch.float64)
    c64, _ = orc.compute_cost(om, torch.tensor(y0, dtype=torch.float64), o, torch.tensor(dw, dtype=torch.float64),
                              hp.stage_cost_fn("hexa", cp), xr64,
                              terminal_cost=lambda x, xr: orc.rotor_one_step_cost(x, None, xr, cpe))
    c64 = c64 + orc.u_slew_cost(o, ts, cp, 6)
    (g64,) = torch.autograd.grad(c64, o)
    assert abs(float(cost[0]) - float(c64)) <= 2e-5 * abs(float(c64))
    assert hp.rel_err(grad.cpu().numpy(), g64.numpy()) < 1e-3
    assert float(cost[0]) > float(cost0[0]) * 1.05  # the terminal term is a visible part of the total
    cv, _, _ = eng.cost_grad(nn, y0[None], opt, ts, xref, stage.desc, terminal=term.desc, N=N, lanes=lanes, want_grad=False, **kw)
    assert abs(float(cv[0]) - float(cost[0])) <= 2e-6 * abs(float(cost[0]))


@pytest.mark.parametrize("N,lanes", [(32, 1), (5, 0), (3, 8)])
def test_problem_mask_skips_and_leaves_outputs_untouched(cuda, N, lanes):
    """sde4_cost_args.problem_mask: masked-out problems keep their previous cost / grad entries, the others are
    evaluated as without a mask (whole warps of masked problems exit early; partly wanted warps compute everything)."""
    from sde4mbrl_b200.engine import Engine
    from oracle import threefry as tf
    P, H = 9, 20
    pm, model, nn, _, _, _, cp, stage, xref = _setup("hexa", H, N, seed=3)
    rng = np.random.default_rng(8)
    y0 = np.stack([hp.start_state("hexa", rng) for _ in range(P)])
    opt = np.stack([hp.controls("hexa", H, rng) for _ in range(P)])
    keys = tf.split(tf.PRNGKey(2), P)
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    c_all, g_all, _ = eng.cost_grad(nn, y0, opt, ts, xref, stage.desc, key=keys, N=N, lanes=lanes)
    c_all, g_all = c_all.clone(), g_all.clone()
    mask = torch.tensor([1, 0, 0, 1, 1, 0, 1, 0, 0], dtype=torch.int32, device="cuda")
    c_out = torch.full((P,), -7.0, device="cuda")
    g_out = torch.full((P, H, 6), -7.0, device="cuda")
    eng.cost_grad(nn, y0, opt, ts, xref, stage.desc, key=keys, N=N, lanes=lanes, cost_out=c_out, grad_out=g_out,
                  problem_mask=mask)
    on = mask.bool()
    assert torch.equal(c_out[on], c_all[on]) and torch.equal(g_out[on], g_all[on])
    assert bool((c_out[~on] == -7.0).all()) and bool((g_out[~on] == -7.0).all())
    cv = torch.full((P,), -7.0, device="cuda")
    eng.cost_grad(nn, y0, opt, ts, xref, stage.desc, key=keys, N=N, lanes=lanes, want_grad=False, cost_out=cv, problem_mask=mask)
    assert bool((cv[~on] == -7.0).all()) and hp.rel_err(cv[on].cpu().numpy(), c_all[on].cpu().numpy()) < 2e-6

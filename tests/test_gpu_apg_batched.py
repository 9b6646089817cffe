"""Batched device-side APG (P problems in lock step, masks instead of lax.cond) vs the host APG
(sde4mbrl_b200/apg.py, itself checked against the oracle APG) run problem by problem."""
import contextlib
import io as _io

import numpy as np
import pytest
import torch

import helpers as hp
from oracle import threefry as tf
from sde4mbrl_b200 import configs, costs, sde_models

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("merge", [True, False, "speculative"])
@pytest.mark.parametrize("N,iters,rtol,box", [(1, 8, 0.0, None), (4, 6, 0.0, None), (1, 12, 0.05, None),
                                              (4, 8, 0.0, (0.325, 0.338))])
def test_batched_apg_matches_per_problem_host_apg(cuda, N, iters, rtol, merge, box):
    """`merge`: x and v evaluated by one value+gradient launch (latency-bound sizes) or, as in the reference, value
    only followed by the gradient at the chosen point (throughput-bound sizes; there the cost of x is taken from the
    accepted line-search candidate unless the box prox moved it -- `box`: tight control bounds, the prox clips);
    "speculative": x_k / v_k of every step size with their gradients in the one launch that also gives the line
    search its costs (smallest sizes)."""
    from sde4mbrl_b200 import apg, nsde, random as R
    from sde4mbrl_b200.apg_batched import BatchedAPG
    P, H = 5, 20
    pm = configs.hexa_model(horizon=H, num_particles=N, stepsize=0.05)
    mpc = configs.hexa_mpc(horizon=H, dt=0.05, num_particles=N, max_iter=iters)
    mpc["apg_mpc"]["rtol"] = rtol  # rtol > 0: some problems stop early -> exercises the not_done masks
    lo, hi = (1e-4, 1.0) if box is None else box
    mpc["input_constr"]["input_bound"] = [[lo, hi]] * 6
    if box is not None:
        mpc["apg_mpc"]["linesearch"]["init_stepsize"] = 0.3
    with contextlib.redirect_stdout(_io.StringIO()):
        _, multi_cost, vprox, _, construct_opt = nsde.create_online_cost_sampling_fn(pm, mpc, sde_constr=sde_models.SDERotorModel)
    nn = configs.random_params(sde_models.SDERotorModel(pm), seed=3)
    cp, op = mpc["cost_params"], mpc["apg_mpc"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    rng = np.random.default_rng(7)
    y0 = np.stack([hp.start_state("hexa", rng) for _ in range(P)])
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    xref[:, :3] = [0.0, 0.5, 1.4]
    keys = tf.split(tf.PRNGKey(31), P)
    x0 = construct_opt(u=np.full(6, 0.33, np.float32))
    eng = multi_cost.engine
    if merge == "speculative":
        solver = BatchedAPG(eng, nn, stage, op, multi_cost.time_steps, N, lb=lo, ub=hi, speculative=True)
    else:
        solver = BatchedAPG(eng, nn, stage, op, multi_cost.time_steps, N, lb=lo, ub=hi, merge_xv_grad=merge)
    solver.init(x0, y0, xref, keys).run(use_graph=False)
    assert solver.speculative == (merge == "speculative")
    x_eager = solver.x_opt().clone()
    solver.init(x0, y0, xref, keys).run(use_graph=True)  # CUDA-graph replay must give the same iterates
    assert torch.equal(solver.x_opt(), x_eager)
    solver.init(x0, y0 * 1.0, xref, keys).run(use_graph=True)  # re-init with new host buffers reuses the graph
    assert torch.equal(solver.x_opt(), x_eager)
    xb = solver.x_opt().cpu().numpy()
    n_done = 0
    for p in range(P):
        cost_p = lambda o, p=p: multi_cost(nn, y0[p], o, keys[p], stage, extra_cost_args=xref)[0]
        prox = lambda x, stepsize=None: vprox(x)
        s = apg.apg(apg.init_apg(x0, cost_p, op), cost_p, op, proximal_fn=prox)
        sb = solver.state_of(p)
        assert sb.num_steps == s.num_steps and sb.not_done == s.not_done
        assert sb.num_iter_no_improv == s.num_iter_no_improv
        # the kernels form y - s*g with one FMA, torch with two roundings: 1-ulp differences in the
        # iterates grow over the iterations (most visibly in |grad|^2), hence the tolerances
        for f in ("opt_cost", "curr_cost", "init_cost", "stepsize", "avg_stepsize", "momentum", "avg_linesearch"):
            np.testing.assert_allclose(float(getattr(sb, f)), float(getattr(s, f)), rtol=1e-4, atol=1e-7, err_msg=f)
        np.testing.assert_allclose(float(sb.grad_sqr), float(s.grad_sqr), rtol=5e-3)
        assert hp.rel_err(xb[p], s.x_opt.cpu().numpy()) < 2e-4
        assert hp.rel_err(sb.yk.cpu().numpy(), s.yk.cpu().numpy()) < 2e-4
        n_done += int(not s.not_done)
        assert float(s.opt_cost) < float(s.init_cost)
    if rtol > 0:
        assert n_done > 0  # the masked path was exercised
    if box is not None:
        assert (xb <= lo + 1e-7).any() or (xb >= hi - 1e-7).any()  # the prox really clipped


@pytest.mark.parametrize("N", [4, 32])
def test_progressive_line_search_gives_the_same_solution(cuda, N):
    """One launch per step size with the already-accepted problems masked out (problem_mask) vs all step sizes in
    one stacked launch: same accepted candidates, same iterates."""
    from sde4mbrl_b200 import nsde
    from sde4mbrl_b200.apg_batched import BatchedAPG
    P, H, iters = 37, 20, 10
    pm = configs.hexa_model(horizon=H, num_particles=N, stepsize=0.05)
    mpc = configs.hexa_mpc(horizon=H, dt=0.05, num_particles=N, max_iter=iters)
    mpc["apg_mpc"]["rtol"] = 0.02  # some problems finish early: finished problems are masked out as well
    mpc["apg_mpc"]["linesearch"]["init_stepsize"] = 0.5  # large first step: the line search has to back-track
    with contextlib.redirect_stdout(_io.StringIO()):
        _, multi_cost, _, _, construct_opt = nsde.create_online_cost_sampling_fn(pm, mpc, sde_constr=sde_models.SDERotorModel)
    nn = configs.random_params(sde_models.SDERotorModel(pm), seed=3)
    cp, op = mpc["cost_params"], mpc["apg_mpc"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    rng = np.random.default_rng(11)
    y0 = np.stack([hp.start_state("hexa", rng) for _ in range(P)])
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    xref[:, :3] = [0.0, 0.5, 1.4]
    keys = tf.split(tf.PRNGKey(5), P)
    x0 = construct_opt(u=np.full(6, 0.33, np.float32))
    res = []
    for prog in (False, True):
        s = BatchedAPG(multi_cost.engine, nn, stage, op, multi_cost.time_steps, N, lb=1e-4, ub=1.0, progressive_ls=prog,
                       lanes=1)
        s.init(x0, y0, xref, keys).run()
        res.append((s.x_opt().cpu().numpy(), s.t["opt_cost"].cpu().numpy(), s.t["num_steps"].cpu().numpy(),
                    s.t["avg_linesearch"].cpu().numpy(), s.t["not_done"].cpu().numpy()))
    (xa, ca, na, la, da), (xb, cb, nb, lb_, db) = res
    np.testing.assert_array_equal(na, nb)
    np.testing.assert_array_equal(da, db)
    np.testing.assert_allclose(la, lb_, rtol=1e-6)
    np.testing.assert_allclose(ca, cb, rtol=1e-6)
    assert hp.rel_err(xa, xb) < 1e-6
    assert (la > 1.0).any() and (da == 0).any()  # back-tracking and early stopping both happened


@pytest.mark.parametrize("mode", ["merge", "two_launch", "speculative"])
@pytest.mark.parametrize("rtol", [0.0, 0.05])
def test_batched_apg_against_oracle_apg_and_oracle_cost(cuda, mode, rtol):
    """The device-side solver against the ORACLE end to end -- oracle/apg_oracle.py (apg.py restated with Python control
    flow) driving the torch oracle's cost and autograd gradient, on the noise the kernels draw (regenerated by the
    oracle's jax.random restatement) -- field by field, including the early-stop masks (rtol > 0).  No product code on
    the reference side of this comparison."""
    from oracle import apg_oracle, nsde_oracle as orc
    from sde4mbrl_b200 import nsde
    from sde4mbrl_b200.apg_batched import BatchedAPG
    P, H, N, iters = 4, 12, 4, 7
    pm = configs.hexa_model(horizon=H, num_particles=N, stepsize=0.05)
    mpc = configs.hexa_mpc(horizon=H, dt=0.05, num_particles=N, max_iter=iters)
    mpc["apg_mpc"]["rtol"] = rtol
    with contextlib.redirect_stdout(_io.StringIO()):
        _, multi_cost, _, _, construct_opt = nsde.create_online_cost_sampling_fn(pm, mpc, sde_constr=sde_models.SDERotorModel)
    nn = configs.random_params(sde_models.SDERotorModel(pm), seed=5)
    cp, op = mpc["cost_params"], mpc["apg_mpc"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    rng = np.random.default_rng(17)
    y0 = np.stack([hp.start_state("hexa", rng) for _ in range(P)])
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    xref[:, :3] = [0.0, 0.5, 1.4]
    keys = tf.split(tf.PRNGKey(77), P)
    x0 = construct_opt(u=np.full(6, 0.33, np.float32))
    x0 = np.asarray(x0.cpu() if isinstance(x0, torch.Tensor) else x0, np.float32)
    kw = {"merge": dict(merge_xv_grad=True), "two_launch": dict(merge_xv_grad=False), "speculative": dict(speculative=True)}[mode]
    solver = BatchedAPG(multi_cost.engine, nn, stage, op, multi_cost.time_steps, N, lb=1e-4, ub=1.0, **kw)
    solver.init(x0, y0, xref, keys).run()
    xb = solver.x_opt().cpu().numpy()
    om = hp.oracle_model("hexa", pm, nn)
    ts = orc.compute_timesteps(pm)
    lb, ub = torch.full((6,), 1e-4), torch.ones(6)
    prox = lambda x, stepsize=None: torch.minimum(torch.maximum(x, lb), ub)
    n_done = 0
    for p in range(P):
        _, dw = orc.sample_rollouts(om, y0[p], x0, keys[p], N)

        def cost_fn(o, p=p, dw=dw):
            c, _ = orc.compute_cost(om, torch.tensor(y0[p]), o, dw, hp.stage_cost_fn("hexa", cp), torch.tensor(xref))
            return c + orc.u_slew_cost(o, ts, cp, 6)
        s = apg_oracle.apg(apg_oracle.init_apg(torch.tensor(x0), cost_fn, op), cost_fn, op, proximal_fn=prox)
        sb = solver.state_of(p)
        assert sb.num_steps == int(s.num_steps) and bool(sb.not_done) == bool(s.not_done)
        assert sb.num_iter_no_improv == int(s.num_iter_no_improv)
        for f in ("opt_cost", "curr_cost", "init_cost", "stepsize", "avg_stepsize", "momentum", "avg_linesearch"):
            np.testing.assert_allclose(float(getattr(sb, f)), float(getattr(s, f)), rtol=2e-4, atol=1e-7, err_msg=f)
        assert hp.rel_err(xb[p], s.x_opt.detach().numpy()) < 2e-3
        n_done += int(not bool(s.not_done))
    if rtol > 0:
        assert n_done > 0

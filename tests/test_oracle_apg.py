"""APG restatement on the least-squares problem of the reference's (unrunnable) tests/opt_test.py:17-35:
0.5*||y - Q x||^2 with Q in R^{30x20}; the oracle APG and the product's host APG must produce the
same iterate sequence, and both must beat plain gradient descent."""
import numpy as np
import torch

from oracle import apg_oracle
from sde4mbrl_b200 import apg as apg_host

OPT = {"max_iter": 60, "max_no_improvement_iter": 60, "moment_scale": None, "beta_init": 0.25, "atol": 1e-8,
       "rtol": 1e-6, "linesearch": {"init_stepsize": 0.01, "max_stepsize": 1.0, "coef": 0.01, "decrease_factor": 0.7,
                                    "increase_factor": 1.3, "reset_option": "increase", "maxls": 4}}


def _problem():
    rng = np.random.default_rng(0)
    Q = torch.tensor(rng.uniform(-4, 4, (30, 20)), dtype=torch.float32)
    y = torch.tensor(rng.uniform(-2, 2, 30), dtype=torch.float32)
    x0 = torch.tensor(rng.uniform(-2, 2, 20), dtype=torch.float32)
    return (lambda x: 0.5 * torch.sum((y - Q @ x) ** 2)), x0, Q, y


def test_apg_converges_and_host_matches_oracle():
    cost, x0, Q, y = _problem()
    prox = lambda x, stepsize=None: torch.clamp(x, -1.5, 1.5)
    tr = []
    so = apg_oracle.apg(apg_oracle.init_apg(x0, cost, OPT, proximal_fn=prox), cost, OPT, proximal_fn=prox, trace=tr)
    sh = apg_host.apg(apg_host.init_apg(x0, cost, OPT, proximal_fn=prox), cost, OPT, proximal_fn=prox)
    assert float(so.opt_cost) < 0.05 * float(so.init_cost)
    assert so.num_steps == sh.num_steps
    np.testing.assert_allclose(float(sh.opt_cost), float(so.opt_cost), rtol=1e-5)
    np.testing.assert_allclose(sh.x_opt.numpy(), so.x_opt.numpy(), rtol=1e-4, atol=1e-5)
    assert float(torch.max(torch.abs(sh.x_opt))) <= 1.5 + 1e-6
    # box-constrained least squares optimum (projected gradient reference)
    x = x0.clone()
    L = float(torch.linalg.matrix_norm(Q, 2) ** 2)
    for _ in range(5000):
        x = torch.clamp(x - (Q.T @ (Q @ x - y)) / L, -1.5, 1.5)
    assert float(so.opt_cost) <= float(cost(x)) * 1.02 + 1e-3


def test_apg_state_fields_and_bookkeeping():
    cost, x0, _, _ = _problem()
    s0 = apg_host.init_opt_state(x0, OPT)
    assert s0.num_steps == 1 and s0.not_done and float(s0.stepsize) == np.float32(0.01)
    s1 = apg_host.init_apg(x0, cost, OPT)
    assert float(s1.init_cost) == float(s1.curr_cost) == float(s1.opt_cost)
    s2 = apg_host.one_step_apg(s1, cost, OPT, None)
    assert s2.num_steps == 2 and float(s2.opt_cost) <= float(s1.opt_cost)
    assert apg_host.APGState._fields == apg_oracle.APGState._fields

"""Cross-check of the oracle against the REFERENCE ITSELF (wuwushrek/sde4mbrl run through jax on the CPU).

Skipped unless jax + dm-haiku import and a copy of the reference is on disk (baseline/ref_jax_cpu.available());
the build image has neither, which is why DESIGN.md calls parity "unpinned".  Wherever they exist these tests pin
oracle/ -- and through it every GPU parity test -- to the reference: rollouts (nsde.py:987-1030), the MPC cost and
its gradient (nsde.py:1341-1597, apg.py:82), and an APG solve (apg.py:53-161)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_jax_cpu as ref  # noqa: E402

import helpers as hp  # noqa: E402
from oracle import apg_oracle, nsde_oracle as orc, threefry as tf  # noqa: E402
from sde4mbrl_b200 import configs  # noqa: E402

_ok, _why = ref.available()
pytestmark = pytest.mark.skipif(not _ok, reason="reference cross-check skipped: " + _why)


def _hexa(H, N, seed):
    rng = np.random.default_rng(seed)
    pm, model, nn = hp.build("hexa", H, N, seed=seed + 1)
    y0, u = hp.start_state("hexa", rng), hp.controls("hexa", H, rng)
    return pm, nn, y0, u


def test_rollout_matches_reference():
    H, N = 20, 8
    pm, nn, y0, u = _hexa(H, N, 3)
    key = tf.PRNGKey(11)
    want = ref.rollout("hexa", pm, nn, y0, u, key, N)
    om = hp.oracle_model("hexa", pm, nn)
    got, _ = orc.sample_rollouts(om, y0, u, key, N, mode=ref.rng_mode())
    assert hp.rel_err(got.numpy(), want) < 2e-5


def test_cost_and_gradient_match_reference():
    H, N = 20, 8
    pm, nn, y0, opt = _hexa(H, N, 5)
    mpc = configs.hexa_mpc(horizon=H, dt=float(pm["stepsize"]), num_particles=N)
    cp = mpc["cost_params"]
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    xref[:, :3] = [0.0, 0.5, 1.4]
    key = tf.PRNGKey(7)
    c_ref, g_ref = ref.rotor_cost_value_and_grad(pm, mpc, cp, nn, y0, opt, key, xref)
    om32 = hp.oracle_model("hexa", pm, nn)
    _, dw = orc.sample_rollouts(om32, y0, opt, key, N, mode=ref.rng_mode())
    om = hp.oracle_model("hexa", pm, nn, torch.float64)
    o = torch.tensor(opt, dtype=torch.float64, requires_grad=True)
    ts = orc.compute_timesteps(pm)
    c, _ = orc.compute_cost(om, torch.tensor(y0, dtype=torch.float64), o, dw.double(), hp.stage_cost_fn("hexa", cp),
                            torch.tensor(xref, dtype=torch.float64))
    c = c + orc.u_slew_cost(o, ts, cp, 6)
    (g,) = torch.autograd.grad(c, o)
    assert abs(c.item() - c_ref) <= 2e-5 * abs(c_ref)
    assert hp.rel_err(g.numpy(), g_ref) < 1e-3


def test_apg_solve_matches_reference():
    H, N = 10, 4
    pm, nn, y0, _ = _hexa(H, N, 9)
    mpc = configs.hexa_mpc(horizon=H, dt=float(pm["stepsize"]), num_particles=N, max_iter=6)
    cp = mpc["cost_params"]
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    key = tf.PRNGKey(2)
    x0 = np.full((H, 6), 0.33, np.float32)
    st_ref = ref.rotor_apg_solve(pm, mpc, cp, mpc["apg_mpc"], nn, y0, x0, key, xref)
    om32 = hp.oracle_model("hexa", pm, nn)
    _, dw = orc.sample_rollouts(om32, y0, x0, key, N, mode=ref.rng_mode())
    ts = orc.compute_timesteps(pm)

    def cost_fn(o):
        c, _ = orc.compute_cost(om32, torch.tensor(y0), o, dw, hp.stage_cost_fn("hexa", cp), torch.tensor(xref))
        return c + orc.u_slew_cost(o, ts, cp, 6)
    lb, ub = torch.full((6,), 1e-4), torch.ones(6)
    prox = lambda x, stepsize=None: torch.minimum(torch.maximum(x, lb), ub)
    st = apg_oracle.apg(apg_oracle.init_apg(torch.tensor(x0), cost_fn, mpc["apg_mpc"]), cost_fn, mpc["apg_mpc"],
                        proximal_fn=prox)
    assert abs(float(st.opt_cost) - float(st_ref["opt_cost"])) <= 2e-4 * abs(float(st_ref["opt_cost"]))
    assert int(st.num_steps) == int(st_ref["num_steps"])
    assert hp.rel_err(st.x_opt.numpy(), st_ref["x_opt"]) < 2e-3

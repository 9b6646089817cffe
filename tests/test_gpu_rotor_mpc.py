"""sde4mbrl_b200/rotor_mpc.py on the GPU: the reference's loaders (pickle + yaml) and receding-horizon MPC
step driven by the reference's SHIPPED hexacopter model (tests/golden/ref_weights_hexa_ahg_v4.npz, re-packed
here into the reference's pickle format)."""
import contextlib
import io as _io
import os
import pickle

import numpy as np
import pytest
import torch
import yaml

import helpers as hp
from oracle import nsde_oracle as orc
from oracle import threefry as tf
from sde4mbrl_b200 import configs, io

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _write_cfg(tmp_path, traj=False):
    nn, nominal = io.load_npz_fixture(os.path.join(GOLD, "ref_weights_hexa_ahg_v4.npz"))
    pkl = tmp_path / "hexa_sde.pkl"
    with open(pkl, "wb") as f:
        pickle.dump({"sde": nn, "nominal": nominal}, f)
    cfg = configs.hexa_mpc(horizon=20, dt=0.05, num_particles=4, max_iter=8)
    cfg["learned_model_params"] = str(pkl)
    cfg["one_step_params"] = {"stepsize": 0.01}
    if traj:
        import pandas as pd
        t = np.arange(0, 6.0, 0.1)
        cols = {"t": t, "x": 0.2 * t, "y": 0.5 * np.sin(t), "z": 1.0 + 0.05 * t, "vx": np.full_like(t, 0.2),
                "vy": 0.5 * np.cos(t), "vz": np.full_like(t, 0.05), "qw": np.ones_like(t), "qx": 0 * t, "qy": 0 * t,
                "qz": 0 * t, "wx": 0 * t, "wy": 0 * t, "wz": 0 * t}
        pd.DataFrame(cols).to_csv(tmp_path / "traj.csv", index=False)
        cfg["trajectory_path"] = str(tmp_path / "traj.csv")
    path = tmp_path / "mpc.yaml"
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    return str(path), nn, nominal


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(_io.StringIO()):
        return fn(*a, **k)


def test_load_predictor_function_matches_oracle(cuda, tmp_path):
    from sde4mbrl_b200 import random as R, rotor_mpc as rm
    path, nn, nominal = _write_cfg(tmp_path)
    pkl = yaml.safe_load(open(path))["learned_model_params"]
    fn, tgrid = rm.load_predictor_function(pkl, modified_params={"num_particles": 6, "horizon": 12}, return_control=True,
                                           return_time_steps=True)
    rng = np.random.default_rng(0)
    y0 = hp.start_state("hexa", rng)
    u = rng.uniform(0.25, 0.45, (12, 6)).astype(np.float32)
    yev, uev = fn(y0, u, R.PRNGKey(3))
    assert tuple(yev.shape) == (6, 13, 13) and tuple(uev.shape) == (6, 12, 6) and tgrid.shape == (13,)
    pm = io.update_params(nominal, {"num_particles": 6, "horizon": 12})
    want, _ = orc.sample_rollouts(orc.make_model("sde_rotor_model", pm, nn), y0, u, tf.PRNGKey(3), 6)
    assert hp.rel_err(yev.cpu().numpy(), want.numpy()) < 2e-5
    # nominal (system-identification) model: no noise -> all particles identical
    fn0 = rm.load_predictor_function(pkl, prior_dist=True, modified_params={"num_particles": 3, "horizon": 5})
    y = fn0(y0, u[:5], R.PRNGKey(1))
    assert torch.equal(y[0], y[1]) and torch.equal(y[0], y[2])


@pytest.mark.parametrize("host_args", [False, True])
def test_setpoint_mpc_closed_loop(cuda, tmp_path, host_args):
    """m_reset / m_mpc / one_step_prediction from load_mpc_from_cfgfile: the MPC drives the learned hexacopter
    toward the set point; key chaining follows sde_mpc_design.py:227."""
    from sde4mbrl_b200 import random as R, rotor_mpc as rm
    path, _, _ = _write_cfg(tmp_path)
    cfg, (m_reset, m_mpc), get_state_at_t, one_step = _quiet(rm.load_mpc_from_cfgfile, path, convert_to_enu=False)
    assert get_state_at_t is None and cfg["_time_steps"].shape == (20,)
    x = configs.hover_state()
    xdes = configs.hover_state()
    xdes[:3] = [0.0, 0.5, 1.4]
    key = tf.PRNGKey(5) if host_args else R.PRNGKey(5)
    sol = m_reset(x, uref=np.full(6, 0.33, np.float32))
    if host_args:
        sol = sol._replace(x_opt=sol.x_opt.cpu(), xk=sol.xk.cpu(), yk=sol.yk.cpu(), grad_yk=sol.grad_yk.cpu())
    d0 = np.linalg.norm(x[:3] - xdes[:3])
    k_expected = tf.split(tf.PRNGKey(5))[0]
    for it in range(15):
        uopt, sol, key, states = m_mpc(x, key, sol, xdes=xdes)
        if it == 0:
            got = key.cpu().numpy().view(np.uint32) if isinstance(key, torch.Tensor) else key
            np.testing.assert_array_equal(got, k_expected)
            assert tuple(uopt.shape) == (20, 6) and tuple(states.shape) == (4, 21, 14)
            assert float(sol.opt_cost) <= float(sol.init_cost)
        u0 = uopt[0].detach().cpu().numpy()
        assert u0.min() >= 1e-4 - 1e-9 and u0.max() <= 1.0 + 1e-9
        for _ in range(5):  # 5 x 0.01 s of the one-step model per 0.05 s MPC step
            x = one_step(x, u0, key).cpu().numpy()
    assert np.isfinite(x).all() and abs(np.linalg.norm(x[6:10]) - 1) < 1e-5
    assert np.linalg.norm(x[:3] - xdes[:3]) < 0.8 * d0  # moved toward the set point


def test_trajectory_mode_targets(cuda, tmp_path):
    from sde4mbrl_b200 import random as R, rotor_mpc as rm
    path, _, _ = _write_cfg(tmp_path, traj=True)
    cfg, (m_reset, m_mpc), get_state_at_t, _ = _quiet(rm.load_mpc_from_cfgfile, path, convert_to_enu=False)
    xc, xevol = get_state_at_t(1.23, _return_evol=True)
    assert xevol.shape == (21, 13) and abs(xc[0] - 0.2 * 1.23) < 1e-5 and abs(xevol[-1, 0] - 0.2 * (1.23 + 1.0)) < 1e-5
    x = get_state_at_t(0.0)
    sol = m_reset(x)
    (uopt, sol, key, states), xref = m_mpc(x, R.PRNGKey(0), sol, curr_t=0.0, return_ref=True)
    np.testing.assert_allclose(xref, get_state_at_t(0.0, _return_evol=True)[1])
    assert tuple(uopt.shape) == (20, 6) and float(sol.opt_cost) <= float(sol.init_cost)


def test_apg_mpc_device_solver_matches_host_loop(cuda, tmp_path):
    """apg_mpc with the device-side solver (default) and with the host apg.py loop: same optimum, same key chain,
    warm start (momentum / average step size of the previous solution) included."""
    from sde4mbrl_b200 import random as R, rotor_mpc as rm
    path, _, _ = _write_cfg(tmp_path)
    cfg, (m_reset, m_mpc), _, _ = _quiet(rm.load_mpc_from_cfgfile, path, convert_to_enu=False)
    x = configs.hover_state()
    x[3:6] = [0.2, -0.1, 0.05]
    xdes = configs.hover_state()
    xdes[:3] = [0.0, 0.5, 1.4]
    res = {}
    for dev_solver in (True, False):
        rm.USE_DEVICE_APG = dev_solver
        try:
            sol = m_reset(x, uref=np.full(6, 0.33, np.float32))
            key = R.PRNGKey(5)
            for _ in range(2):  # the second call starts from the first solution (warm start)
                uopt, sol, key, states = m_mpc(x, key, sol, xdes=xdes)
            res[dev_solver] = (uopt.cpu().numpy(), float(sol.opt_cost), int(sol.num_steps), states.cpu().numpy(),
                               key.cpu().numpy())
        finally:
            rm.USE_DEVICE_APG = True
    ud, cd_, nd, sd, kd = res[True]
    uh, ch, nh, sh, kh = res[False]
    np.testing.assert_array_equal(kd, kh)
    assert nd == nh
    assert abs(cd_ - ch) <= 2e-4 * abs(ch)
    assert hp.rel_err(ud, uh) < 5e-3
    assert hp.rel_err(sd[..., :13], sh[..., :13]) < 5e-3

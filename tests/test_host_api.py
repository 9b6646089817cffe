"""Host-side mirror of the reference interface (no GPU needed): time steps, constraint set-up,
cost descriptors, parameter packing, argument errors."""
import os

import numpy as np
import pytest

from sde4mbrl_b200 import _lib as L
from sde4mbrl_b200 import configs, costs, nsde, sde_models


def test_compute_timesteps():
    p = {"horizon": 5, "stepsize": 0.1}
    np.testing.assert_allclose(nsde.compute_timesteps(p), [0.1] * 5)
    p = {"horizon": 5, "stepsize": 0.1, "num_short_dt": 2, "short_step_dt": 0.01, "long_step_dt": 0.2}
    np.testing.assert_allclose(nsde.compute_timesteps(p), [0.01, 0.01, 0.2, 0.2, 0.2])
    with pytest.raises(AssertionError):
        nsde.compute_timesteps({"horizon": 2, "stepsize": 0.1, "num_short_dt": 3})


def test_initialize_problem_constraints(capsys):
    (hu, lb, ub), (hx, sp, idx, pen, slb, sub, w, ss) = nsde.initialize_problem_constraints(
        4, 2, {"input_constr": {"input_id": [1], "input_bound": [[-1.0, 2.0]]},
               "state_constr": {"state_id": [0, 3], "state_bound": [[-1, 1], [-4, 2]], "state_penalty": [1.0, 2.0],
                                "slack_proximal": True, "slack_scaling": [1.0, 2.0], "constr_pen": 3.0}})
    assert hu and hx and sp and w == 3.0
    assert lb[0] == -np.inf and lb[1] == -1.0 and ub[1] == 2.0
    np.testing.assert_allclose(slb, [-1.0, -2.0])
    np.testing.assert_allclose(sub, [1.0, 1.0])
    (hu, _, _), (hx, *_rest) = nsde.initialize_problem_constraints(4, 2, {})
    assert not hu and not hx


def test_rotor_cost_descriptor_follows_cost_params():
    cp = configs.hexa_mpc()["cost_params"]
    c = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6).desc
    assert c.kind == L.COST_ROTOR_TRACK and abs(c.res_mult - 0.01) < 1e-9 and c.has_slew == 1
    assert list(c.w_x)[:13] == [20, 20, 40, 3, 3, 3, 0] + [pytest.approx(0.1), pytest.approx(0.1), 20, 3, 3, 3]
    assert all(abs(c.w_u[j] - 0.01) < 1e-9 and abs(c.uref[j] - 0.33) < 1e-7 for j in range(6))
    with pytest.raises(L.Sde4Error):
        costs.one_step_cost_f(dict(cp, urate_err=1.0), 6)
    with pytest.raises(L.Sde4Error):
        costs.one_step_cost_f(cp, 6)(None, None, None)  # not a Python callable


def test_pack_layout_roundtrip():
    m = sde_models.CartPoleSDE(configs.cartpole_model(horizon=3, num_particles=2))
    nn = m.init_params(seed=1)
    d = m.describe()
    off, total = m.layout(d)
    p = m.pack(nn, d)
    sc = nn["cart_pole_sde"]["scaler"]
    np.testing.assert_array_equal(p[off[1]:off[1] + 4], sc)
    w0 = nn["cart_pole_sde/~init_residual_networks/res_forces/~/linear_0"]["w"]      # [3, 8]
    np.testing.assert_array_equal(p[off[3]:off[3] + 24].reshape(3, 8), w0)
    w2 = nn["cart_pole_sde/~init_residual_networks/res_forces/~/linear_2"]["w"]      # [24, 2], stored transposed
    base = off[3] + 3 * 8 + 8 + 8 * 24 + 24
    np.testing.assert_array_equal(p[base:base + 48].reshape(2, 24), w2.T)


def test_motor_polynomial_and_fixed_params():
    pm = configs.hexa_model(horizon=2, num_particles=1)
    m = sde_models.SDERotorModel(pm)
    nn = configs.random_params(m, seed=0)
    p = m.pack(nn)
    # order 1 with constant term: thrust = coeffMot0*u + coeffMot1 -> c3=c2=0, c1=coeffMot0, c0=coeffMot1
    assert p[7] == 0 and p[8] == 0 and p[9] == np.float32(9.169) and p[10] == np.float32(0.786)
    pm2 = dict(pm, fixed_init_params=True)
    m2 = sde_models.SDERotorModel(pm2)
    p2 = m2.pack(nn)
    assert p2[0] == np.float32(2.295)  # get_param returns the YAML constant (sde_rotor_model.py:146-158)


def test_errors_for_unsupported_requests():
    with pytest.raises(L.Sde4Error):
        nsde.create_sampling_fn(configs.hexa_model(2, 1), sde_constr=object)
    with pytest.raises(L.Sde4Error):
        sde_models.SDERotorModel(dict(configs.hexa_model(2, 1), control_augmented_state=True)).describe()
    # control_dependent_noise widens the density input by n_u (nsde.py:362); architectures without such an
    # instantiation are rejected by the library, not silently evaluated without the controls
    m = sde_models.CartPoleSDE(dict(configs.cartpole_model(2, 1), control_dependent_noise=True))
    assert m.describe().density.n_in == 3 + 1 and m.layout(m.describe())[1] > 0
    with pytest.raises(L.Sde4Error):
        mb = sde_models.CartPoleSDE(dict(configs.cartpole_model(2, 1, side_info=False), control_dependent_noise=True))
        mb.layout(mb.describe())
    with pytest.raises(L.Sde4Error):
        sde_models.SDERotorModel(dict(configs.hexa_model(2, 1), sde_solver="runge_kutta"))


def test_rotor_mpc_frames_and_targets():
    """Host-side helpers of sde4mbrl_b200/rotor_mpc.py (sde_mpc_design.py:37-70, rotor_uav/utils.py)."""
    from sde4mbrl_b200 import rotor_mpc as rm
    rng = np.random.default_rng(0)
    # NED <-> ENU: positions (x, y, z) -> (y, x, -z); body rates (x, y, z) -> (x, -y, -z); involution
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    x = np.concatenate((rng.standard_normal(6), q, rng.standard_normal(3))).astype(np.float32)
    y = rm.ned2enu(x)
    np.testing.assert_allclose(y[:3], [x[1], x[0], -x[2]], atol=1e-6)
    np.testing.assert_allclose(y[3:6], [x[4], x[3], -x[5]], atol=1e-6)
    np.testing.assert_allclose(y[10:13], [x[10], -x[11], -x[12]], atol=1e-6)
    assert abs(np.linalg.norm(y[6:10]) - 1) < 1e-6
    back = rm.ned2enu(y)
    # quaternions are equal up to sign
    np.testing.assert_allclose(back[:6], x[:6], atol=1e-6)
    assert min(np.abs(back[6:10] - x[6:10]).max(), np.abs(back[6:10] + x[6:10]).max()) < 1e-6
    # level attitude with yaw psi in NED is level with yaw pi/2 - psi in ENU
    psi = 0.3
    lvl = np.concatenate((np.zeros(6), rm.quat_from_euler(0.0, 0.0, psi), np.zeros(3)))
    qe = rm.ned2enu(lvl)[6:10]
    want = rm.quat_from_euler(0.0, 0.0, np.pi / 2 - psi)
    assert min(np.abs(qe - want).max(), np.abs(qe + want).max()) < 1e-6
    # extract_targets: piecewise-linear look-up == np.interp inside the table, linear extrapolation outside
    t = np.cumsum(rng.uniform(0.05, 0.2, 40))
    xs = rng.standard_normal((40, 13))
    steps = np.cumsum(np.full(20, 0.05))
    cur, seq = rm.extract_targets(t[3] + 0.01, t, xs, steps)
    assert seq.shape == (21, 13)
    for i in range(13):
        np.testing.assert_allclose(seq[:, i], np.interp(np.concatenate(([t[3] + 0.01], t[3] + 0.01 + steps)), t, xs[:, i]),
                                   atol=1e-5)
    np.testing.assert_allclose(cur, seq[0])
    _, late = rm.extract_targets(t[-1] - 0.2, t, xs, steps)
    slope = (xs[-1] - xs[-2]) / (t[-1] - t[-2])
    np.testing.assert_allclose(late[-1], xs[-1] + slope * (t[-1] - 0.2 + steps[-1] - t[-1]), atol=1e-4)
    tv, sv = rm.parse_trajectory({"t": t, **{n: xs[:, i] for i, n in enumerate(rm.state_names)}})
    np.testing.assert_array_equal(sv, xs)


def test_weight_penalty_matching_follows_get_penalty_parameters():
    """Penalty coefficients and nominal values of the weight-decay term (nsde.py:1155-1168, utils.py:119-146): a key
    matching a module name covers the whole module and wins over leaf-name matches; nominal_parameters_val and
    special_parameters_val move the point the parameters are pulled towards."""
    import contextlib
    import io
    from sde4mbrl_b200 import configs, nsde, sde_models
    pm = configs.cartpole_model(horizon=1, num_particles=1)
    lp = {"seed": 0, "stepsize": 0.02, "data_stepsize": 0.02, "num_particles": 1, "num_particles_test": 1, "horizon": 4,
          "obs_weights": [1, 1, 1, 1, 1], "pen_data": 0.0, "pen_weights": 1.0, "default_weights": 2.0,
          "special_parameters_pen": {"density": 0.0, "linear_1": 5.0, "b": 3.0},
          "default_parameters_val": 0.25, "nominal_parameters_val": {"cart_pole_sde": {"scaler": -1.0}},
          "special_parameters_val": {"linear_2": 0.5}}
    with contextlib.redirect_stdout(io.StringIO()):
        nn, loss_fn, _, _ = nsde.create_model_loss_fn(pm, lp, sde_constr=sde_models.CartPoleSDE, verbose=False)
    pen = loss_fn.weight_penalty if hasattr(loss_fn, "weight_penalty") else None
    assert pen is not None
    tot, grads = pen(nn, True)
    want = 0.0
    for mod, leaves in nn.items():
        for name, arr in leaves.items():
            if "linear_1" in mod:
                coeff = 5.0            # module-level match; the LATER key of the dict wins over "density"
            elif "density" in mod:
                coeff = 0.0            # the whole module, leaf names not looked at
            elif "b" in name:
                coeff = 3.0            # leaf-level match in the remaining modules (no module name holds a "b")
            else:
                coeff = 2.0
            nom = 0.25
            if mod == "cart_pole_sde" and name == "scaler":
                nom = -1.0
            if "linear_2" in mod:
                nom = 0.5
            a = np.asarray(arr, np.float64)
            want += coeff * float(np.sum((a - nom) ** 2))
            np.testing.assert_allclose(grads[mod][name], 2.0 * coeff * (a - nom), rtol=1e-6, atol=1e-7)
    assert abs(tot - want) <= 1e-9 * max(1.0, abs(want))


def test_gen_inst_emits_a_registration_for_an_edited_yaml(tmp_path):
    """tools/gen_inst.py: a configuration with edited hidden_layers becomes an Arch<...> alias + registration line of
    csrc/inst_user.cu (compiled by the Makefile when present); widths above the compiled range are refused."""
    import subprocess
    import sys
    import yaml
    from sde4mbrl_b200 import configs
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pm = configs.cartpole_model(horizon=10, num_particles=4)
    pm["diffusion_density_nn"]["density_nn"]["hidden_layers"] = [16, 24]
    cfg = tmp_path / "cart.yaml"
    cfg.write_text(yaml.safe_dump({"model": pm}))
    out = tmp_path / "inst_user.cu"
    tool = os.path.join(root, "tools", "gen_inst.py")
    subprocess.run([sys.executable, tool, "cartpole", str(cfg), "--name", "cart_a", "--train", "--out", str(out)], check=True, capture_output=True)
    subprocess.run([sys.executable, tool, "cartpole", str(cfg), "--name", "cart_b", "--lanes", "8", "--out", str(out)], check=True, capture_output=True)
    subprocess.run([sys.executable, tool, "cartpole", str(cfg), "--name", "cart_a", "--out", str(out)], check=True, capture_output=True)  # replaces
    text = out.read_text()
    assert text.count("// >>> cart_a") == 1 and text.count("// >>> cart_b") == 1
    assert "Net<3, 16, 24, 1, SWISH>" in text and "SDE4_REGISTER(cart_a, CartpoleModel<User_cart_a>)" in text
    assert "SDE4_REGISTER_LS(cart_b, CartpoleModel, User_cart_b, 8)" in text
    pm["diffusion_density_nn"]["density_nn"]["hidden_layers"] = [64, 64]
    cfg.write_text(yaml.safe_dump({"model": pm}))
    r = subprocess.run([sys.executable, tool, "cartpole", str(cfg), "--out", str(out)], capture_output=True, text=True)
    assert r.returncode != 0 and "outside the compiled range" in (r.stderr + r.stdout)


def test_bound_cost_declines_what_the_device_solver_does_not_cover():
    """``multi_sampling.bind(...)`` (the closure apg.py is handed, as an object): ``solve_on_device`` must return None --
    i.e. leave the solve to the host loop of apg.py -- for a configuration without line search, a proximal operator that
    is not the box, and a state that is not a fresh ``init_apg`` state.  None of these reaches the GPU."""
    import contextlib, io
    import torch
    from sde4mbrl_b200 import apg
    H = 5
    pm = configs.hexa_model(horizon=H, num_particles=2, stepsize=0.05)
    mpc = configs.hexa_mpc(horizon=H, dt=0.05, num_particles=2, max_iter=3)
    with contextlib.redirect_stdout(io.StringIO()):
        _, multi_cost, vprox, _, _ = nsde.create_online_cost_sampling_fn(pm, mpc, sde_constr=sde_models.SDERotorModel)
    cp = mpc["cost_params"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    xref = np.zeros((H + 1, 13), np.float32)
    cost_b = multi_cost.bind({}, np.zeros(13, np.float32), np.array([0, 1], np.uint32), stage, extra_cost_args=xref)
    assert callable(cost_b) and hasattr(cost_b, "value_and_grad")
    op = mpc["apg_mpc"]
    x0 = torch.full((H, 6), 0.3)
    st = apg.init_opt_state(x0, op)  # a fresh state (no cost evaluation)
    no_ls = {k: v for k, v in op.items() if k != "linesearch"}
    no_ls["stepsize"] = 0.01
    assert cost_b.solve_on_device(st, no_ls, None) is None                       # no line search
    assert cost_b.solve_on_device(st, op, lambda x, stepsize=None: x) is None    # a proximal operator that is not the box
    assert cost_b.solve_on_device(st._replace(num_steps=4), op, vprox) is None   # in the middle of a solve
    assert cost_b.solve_on_device(st._replace(xk=x0 + 1.0), op, vprox) is None   # xk != yk

"""Training-loss path (BASELINE configs[4] second half): data term of
create_model_loss_fn(...).loss_fn and its gradient w.r.t. the network parameters (K2 training
variant + K5) vs the oracle's float64 autograd."""
import numpy as np
import pytest
import torch

import helpers as hp
from oracle import nsde_oracle as orc
from oracle import threefry as tf
from sde4mbrl_b200 import configs, sde_models

pytestmark = pytest.mark.gpu


def _tree64(nn):
    return {m: {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=True) for k, v in d.items()}
            for m, d in nn.items()}


# scale = "yaml": the reference's initialiser ranges (weights ~1e-3 .. 1e-1), where an ABSOLUTE activation error is
# largest relative to the hidden activations -- and the parameter gradients are what it would show up in
@pytest.mark.parametrize("scale", ["fan_in", "yaml"])
@pytest.mark.parametrize("kind,cls,oname,B,N,H", [("cartpole", sde_models.CartPoleSDE, "cart_pole_sde", 12, 2, 10),
                                                    ("cartpole_bb", sde_models.CartPoleSDE, "cart_pole_sde", 5, 1, 30),
                                                    ("msd", sde_models.MassSpringDamper, "mass_spring_damper", 7, 3, 20)])
def test_loss_and_param_grad(cuda, kind, cls, oname, B, N, H, scale):
    from sde4mbrl_b200 import nsde, random as R
    mk = hp.KINDS[kind][0]
    pm = mk(horizon=1, num_particles=1)
    ow = [9, 5, 1, 1.0, 10] if oname == "cart_pole_sde" else [0.5, 2.0]
    lp = {"seed": 1, "stepsize": pm["stepsize"], "data_stepsize": pm["stepsize"], "num_particles": N, "num_particles_test": 4,
          "horizon": H, "obs_weights": ow, "pen_data": 1.0, "pen_weights": 1e-3, "default_weights": 1.0,
          "special_parameters_pen": {"scaler": 0, "density": 0}}
    _, loss_fn, nonneg, test_fn = nsde.create_model_loss_fn(pm, lp, sde_constr=cls, verbose=False)
    nn = configs.random_params(cls(pm), seed=4, scale=scale)
    rng = np.random.default_rng(0)
    nx, nu = pm["n_x"], pm["n_u"]
    y = (rng.normal(0, 0.5, (B, H + 1, nx)) + (np.array([0, 0, 3.0, 0]) if nx == 4 else 0)).astype(np.float32)
    u = rng.uniform(-1, 1, (B, H, nu)).astype(np.float32)
    total, info, grads = loss_fn.value_and_grad(nn, y, u, R.PRNGKey(3))
    # oracle: same noise chain, float64 autograd w.r.t. the parameter tree
    tree = _tree64(nn)
    om = orc.make_model(oname, pm, tree, torch.float64)
    dw = torch.tensor(orc.loss_noise(tf.PRNGKey(3), B, N, H, nx), dtype=torch.float64)
    Ld, _ = orc.data_loss(om, oname, torch.tensor(y, dtype=torch.float64), torch.tensor(u, dtype=torch.float64), dw, ow)
    coeffs = {m: {k: (0.0 if ("scaler" in m or "scaler" in k or "density" in m) else 1.0) for k in d} for m, d in tree.items()}
    nominal = {m: {k: torch.zeros_like(v) for k, v in d.items()} for m, d in tree.items()}
    Lw = orc.weight_penalty(tree, nominal, coeffs)
    Lt = Ld + 1e-3 * Lw
    Lt.backward()
    assert abs(info["dataLoss"] - float(Ld)) < 3e-5 * abs(float(Ld))
    assert abs(total - float(Lt)) < 3e-5 * abs(float(Lt))
    gmax = max(float(v.grad.abs().max()) for d in tree.values() for v in d.values())
    for m, d in tree.items():
        for k, v in d.items():
            got, want = grads[m][k], v.grad.numpy()
            assert got.shape == want.shape
            # per-leaf: 2e-3 of the leaf's own max, plus a floor at 1e-5 of the global max (the scaler's
            # gradient is ~1e-9 of the others: exp(scaler)*1e-7, nsde.py:413)
            assert np.max(np.abs(got - want)) <= 2e-3 * np.max(np.abs(want)) + 1e-5 * gmax, (m, k)
    # loss_fn without gradient gives the same value; test_fn runs and returns the reference's keys
    t2, _ = loss_fn(nn, y, u, R.PRNGKey(3))
    assert abs(t2 - total) < 1e-6 * abs(total)
    err, td = test_fn(nn, y, u, R.PRNGKey(5))
    assert set(td) == {"TestErrorData", "TestStdData"} and np.isfinite(err)


@pytest.mark.parametrize("mu_type,N,partial,cdn", [("dnn", 2, 0, False), ("global", 1, 0, False), ("constant", 1, 0, False),
                                                   ("dnn", 2, 5, False), ("global", 3, 2, False),
                                                   ("dnn", 2, 0, True), ("global", 2, 3, True)])
def test_density_regulariser_terms_and_grads(cuda, mu_type, N, partial, cdn):
    """Full training loss of cartpole_sde.yaml:95-137 (data term + distance-aware regulariser + weight penalty):
    gradient-norm / strong-convexity / mu terms with the reference's key chain for the ball samples, and the
    gradient w.r.t. every parameter, vs the oracle's float64 autograd."""
    from sde4mbrl_b200 import nsde, random as R
    B, H = 4, 8
    pm = configs.cartpole_model(horizon=1, num_particles=1)
    if cdn:  # control_dependent_noise (nsde.py:362): density / mu nets and the ball live in [reduced state, u]
        pm["control_dependent_noise"] = True
    ow = [9, 5, 1, 1.0, 10]
    lm = {"dnn": {"type": "dnn", "init_value": 0.3, "activation_fn": "tanh", "hidden_layers": [8, 8]},
          "global": True, "constant": False}[mu_type]
    lp = {"seed": 1, "stepsize": pm["stepsize"], "data_stepsize": pm["stepsize"], "num_particles": N, "num_particles_test": 4,
          "horizon": H, "obs_weights": ow, "pen_data": 1.0, "pen_weights": 1e-3, "default_weights": 1.0,
          "special_parameters_pen": {"scaler": 0, "density": 0},
          "density_loss": {"learn_mucoeff": lm, "mu_coeff": 10.0, "ball_radius": [0.1, 0.1, 0.4] + ([0.3] if cdn else []), "ball_nsamples": 6},
          "pen_grad_density": 1.0, "pen_density_scvex": 1.0, "pen_mu_type": "lin_inv", "pen_mu_coeff": 1000.0,
          "pen_scvex_mult": 0.01}
    if partial:  # num_sample2consider + density_on_partial (nsde.py:753-760, 794-797)
        lp.update(num_sample2consider=partial, density_on_partial=True)
    nn0, loss_fn, _, _ = nsde.create_model_loss_fn(pm, lp, sde_constr=sde_models.CartPoleSDE, verbose=False)
    dl = lp["density_loss"]
    assert dl["learn_mucoeff"]["type"] == mu_type and (lp["pen_mu_coeff"] == 0.0) == (mu_type == "constant")
    nn = configs.random_params(sde_models.CartPoleSDE(pm), seed=4)
    # spread the density output so that sigmoid(d - 7) and the hinge are active
    nn["cart_pole_sde/~construct_diffusion_density_nn/density/~/linear_2"]["w"] *= np.float32(3.0)
    mu_names = [m for m in nn0 if "mu_coeff" in m]
    for m in mu_names:
        nn[m] = {k: np.array(v) for k, v in nn0[m].items()}
    if mu_type == "global":
        nn["cart_pole_sde"]["mu_coeff"] = np.float32(0.2)
    assert (len(mu_names) == 3) == (mu_type == "dnn")
    rng = np.random.default_rng(0)
    y = (rng.normal(0, 0.5, (B, H + 1, 4)) + np.array([0, 0, 3.0, 0])).astype(np.float32)
    u = rng.uniform(-1, 1, (B, H, 1)).astype(np.float32)
    assert loss_fn.regulariser == "native"  # (density_on_partial: the call below falls to the torch path by itself)
    total, info, grads = loss_fn.value_and_grad(nn, y, u, R.PRNGKey(3))
    tree = _tree64(nn)
    om = orc.make_model("cart_pole_sde", pm, tree, torch.float64)
    y64, u64 = torch.tensor(y, dtype=torch.float64), torch.tensor(u, dtype=torch.float64)
    dw = torch.tensor(orc.loss_noise(tf.PRNGKey(3), B, N, H, 4), dtype=torch.float64)
    idx = orc.loss_step_indices(tf.PRNGKey(3), B, N, H, partial) if partial else None
    Ld, _ = orc.data_loss(om, "cart_pole_sde", y64, u64, dw, ow, step_idx=idx)
    gn, sc, mu, _ = orc.density_loss_terms(om, y64, u64, tf.PRNGKey(3), N, dl, step_idx=idx)
    coeffs = {m: {k: (0.0 if ("scaler" in m or "scaler" in k or "density" in m) else 1.0) for k in d} for m, d in tree.items()}
    nominal = {m: {k: torch.zeros_like(v) for k, v in d.items()} for m, d in tree.items()}
    Lw = orc.weight_penalty(tree, nominal, coeffs)
    Lt, oi = orc.training_total(lp, dl, Ld, gn, sc, mu, Lw)
    Lt.backward()
    assert float(oi["densitySCVEX"]) > 0 and float(oi["gradDensity"]) > 0  # the terms are exercised
    for k in ("gradDensity", "densitySCVEX", "muCoeff"):
        assert abs(info[k] - float(oi[k])) <= 2e-4 * abs(float(oi[k])) + 1e-12, k
    assert abs(total - float(Lt)) < 5e-5 * abs(float(Lt))
    gmax = max(float(v.grad.abs().max()) for d in tree.values() for v in d.values() if v.grad is not None)
    for m, d in tree.items():
        for k, v in d.items():
            want = v.grad.numpy() if v.grad is not None else np.zeros(v.shape)
            assert np.max(np.abs(grads[m][k] - want)) <= 3e-3 * np.max(np.abs(want)) + 1e-5 * gmax, (m, k)


@pytest.mark.parametrize("kind,cls,B,N,H", [("cartpole", sde_models.CartPoleSDE, 9, 2, 12), ("cartpole_bb", sde_models.CartPoleSDE, 33, 1, 30)])
def test_lane_split_training_kernel_matches_one_thread_per_sample(cuda, kind, cls, B, N, H):
    """The lane-split training kernel (EvLT<L, true>: every lane streams the rows of its own hidden units) and the
    one-thread-per-sample one give the same loss and parameter gradients."""
    from sde4mbrl_b200 import costs
    from sde4mbrl_b200.engine import Engine
    pm = hp.KINDS[kind][0](horizon=H, num_particles=N)
    model = cls(pm)
    nn = configs.random_params(model, seed=6)
    eng = Engine(model)
    rng = np.random.default_rng(2)
    y = (rng.normal(0, 0.5, (B, H + 1, 4)) + np.array([0, 0, 3.0, 0])).astype(np.float32)
    u = rng.uniform(-1, 1, (B, H, 1)).astype(np.float32)
    ow = np.array([9, 5, 1, 1.0, 10])
    dc = costs.QuadraticCost(list(1.0 / ow ** 2), [0.0], [0.0], feat="cartpole")
    keys = tf.split(tf.PRNGKey(8), B)
    ts = orc.compute_timesteps(pm)
    l1, g1 = eng.loss_grad(nn, y, u, ts, dc.desc, keys, N=N, lanes=1)
    l1, g1 = l1.clone(), g1.clone()
    l0, g0 = eng.loss_grad(nn, y, u, ts, dc.desc, keys, N=N, lanes=0)
    assert hp.rel_err(l0.cpu().numpy(), l1.cpu().numpy()) < 2e-5
    g0, g1 = g0.cpu().numpy(), g1.cpu().numpy()
    assert np.abs(g1).max() > 0
    assert np.max(np.abs(g0 - g1)) <= 2e-4 * np.abs(g1).max()


@pytest.mark.parametrize("lanes,dens", [(0, (32, 32, "swish")), (1, (32, 32, "swish")), (0, (16, 16, "tanh"))])
def test_rotor_training_loss_and_physical_scalar_gradients(cuda, lanes, dens):
    """Training loss of the multirotor nSDE: gradients w.r.t. the three networks, the scaler AND the learnable physical
    scalars (mass, inertia, arm lengths, drag / motor coefficients, aerodynamic drag terms) vs float64 autograd of
    the oracle; lane-split (8 lanes) and one-thread-per-sample training kernels."""
    from sde4mbrl_b200 import nsde, random as R
    B, N, H = 6, 2, 8
    pm = configs.hexa_model(horizon=1, num_particles=1, stepsize=0.01, density=dens[:2], density_act=dens[2])
    lp = {"seed": 1, "stepsize": pm["stepsize"], "data_stepsize": pm["stepsize"], "num_particles": N, "num_particles_test": 2,
          "horizon": H, "obs_weights": [2.6, 3.1, 2.7, 3.2, 3.5, 1.5, 1, 0.3, 0.3, 0.4, 4, 3.5, 1.4], "pen_data": 1.0,
          "pen_weights": 0.0}
    _, loss_fn, _, _ = nsde.create_model_loss_fn(pm, lp, sde_constr=sde_models.SDERotorModel, verbose=False)
    model = sde_models.SDERotorModel(pm)
    nn = configs.random_params(model, seed=4)
    rng = np.random.default_rng(0)
    y = np.stack([np.stack([hp.start_state("hexa", rng) for _ in range(H + 1)]) for _ in range(B)]).astype(np.float32)
    u = rng.uniform(0.25, 0.45, (B, H, 6)).astype(np.float32)
    eng = loss_fn.engine
    keys = tf.split(tf.PRNGKey(3), B)
    from sde4mbrl_b200 import costs
    ow = np.asarray(lp["obs_weights"], np.float64)
    dc = costs.QuadraticCost(list(1.0 / ow ** 2), [0.0] * 6, [0.0] * 6)
    loss_b, gpacked = eng.loss_grad(nn, y, u, orc.compute_timesteps(dict(pm, horizon=H)), dc.desc, keys, N=N, lanes=lanes)
    grads = model.unpack_grads(gpacked.cpu().numpy(), eng.desc)
    tree = _tree64(nn)
    om = orc.make_model("sde_rotor_model", dict(pm, horizon=H), tree, torch.float64)
    dw = torch.tensor(orc.loss_noise(tf.PRNGKey(3), B, N, H, 13), dtype=torch.float64)
    Ld, _ = orc.data_loss(om, "sde_rotor_model", torch.tensor(y, dtype=torch.float64), torch.tensor(u, dtype=torch.float64), dw,
                          lp["obs_weights"])
    Ld.backward()
    assert abs(float(loss_b.mean()) - float(Ld)) < 3e-5 * abs(float(Ld))
    top = tree["sde_rotor_model"]
    phys = [k for k in top if k != "scaler"]
    assert {"mass", "Ixx", "Iyy", "Izz", "Lmx", "CoeffDrag", "coeffMot0", "coeffMot1"} <= set(phys)
    for k in phys:
        want = float(top[k].grad)
        got = float(grads["sde_rotor_model"][k])
        assert abs(got - want) <= 2e-3 * abs(want) + 1e-6 * max(abs(float(t.grad)) for t in (top[n] for n in phys)), (k, got, want)
    gmax = max(float(v.grad.abs().max()) for m, d in tree.items() if m != "sde_rotor_model" for v in d.values())
    for m, d in tree.items():
        if m == "sde_rotor_model":
            continue
        for k, v in d.items():
            assert np.max(np.abs(grads[m][k] - v.grad.numpy())) <= 2e-3 * float(v.grad.abs().max()) + 1e-5 * gmax, (m, k)


@pytest.mark.parametrize("fixture,n_u", [("ref_weights_hexa_ahg_v4.npz", 6), ("ref_weights_iris_sitl.npz", 4)])
def test_training_gradients_on_shipped_rotor_models(cuda, fixture, n_u):
    """The reference's own learned multirotor models (hexacopter 16x16 tanh density, quadrotor 32x32 swish): training
    loss value and every parameter gradient -- networks, scaler, physical scalars -- vs the oracle's autograd."""
    import os
    from sde4mbrl_b200 import io, nsde
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    nn, nominal = io.load_npz_fixture(os.path.join(gold, fixture))
    B, N, H = 5, 2, 6
    pm = dict(nominal)
    pm.pop("density_loss", None)
    pm.pop("num_sample2consider", None)
    lp = {"seed": 0, "stepsize": pm["stepsize"], "data_stepsize": pm["stepsize"], "num_particles": N, "num_particles_test": 2,
          "horizon": H, "obs_weights": list(pm.get("state_scaling", [1.0] * 13)), "pen_data": 1.0, "pen_weights": 0.0,
          "num_sample2consider": 4}  # the reference's training configs subsample the time indexes
    _, loss_fn, _, _ = nsde.create_model_loss_fn(pm, lp, sde_constr=sde_models.SDERotorModel, verbose=False)
    rng = np.random.default_rng(1)
    y = np.stack([np.stack([hp.start_state("hexa", rng) for _ in range(H + 1)]) for _ in range(B)]).astype(np.float32)
    u = rng.uniform(0.25, 0.45, (B, H, n_u)).astype(np.float32)
    from sde4mbrl_b200 import random as R
    total, info, grads = loss_fn.value_and_grad(nn, y, u, R.PRNGKey(9))
    tree = _tree64(nn)
    om = orc.make_model("sde_rotor_model", pm, tree, torch.float64)
    dw = torch.tensor(orc.loss_noise(tf.PRNGKey(9), B, N, H, 13), dtype=torch.float64)
    Ld, _ = orc.data_loss(om, "sde_rotor_model", torch.tensor(y, dtype=torch.float64), torch.tensor(u, dtype=torch.float64), dw,
                          lp["obs_weights"], step_idx=orc.loss_step_indices(tf.PRNGKey(9), B, N, H, 4))
    Ld.backward()
    assert abs(info["dataLoss"] - float(Ld)) < 5e-5 * abs(float(Ld))
    checked = 0
    gm = max(float(v.grad.abs().max()) for d in tree.values() for v in d.values() if v.grad is not None)
    for m, d in tree.items():
        for k, v in d.items():
            if v.grad is None:
                continue
            want = v.grad.numpy()
            assert np.max(np.abs(np.asarray(grads[m][k]) - want)) <= 3e-3 * np.max(np.abs(want)) + 1e-6 * gm, (m, k)
            checked += 1
    assert checked >= 20


@pytest.mark.parametrize("kind,cls,oname,B,N,H,m,lanes", [("cartpole", sde_models.CartPoleSDE, "cart_pole_sde", 6, 3, 12, 5, 0),
                                                          ("cartpole", sde_models.CartPoleSDE, "cart_pole_sde", 6, 3, 12, 5, 1),
                                                          ("msd", sde_models.MassSpringDamper, "mass_spring_damper", 4, 2, 9, 1, 0)])
def test_num_sample2consider_subsampled_loss(cuda, kind, cls, oname, B, N, H, m, lanes):
    """``num_sample2consider`` < horizon (nsde.py:753-760): every (minibatch element, particle) keeps its own random
    subset of time indexes -- drawn on the device with the reference's key chain -- and only those enter the data
    term and its gradient."""
    from sde4mbrl_b200 import nsde, random as R
    pm = hp.KINDS[kind][0](horizon=1, num_particles=1)
    ow = [9, 5, 1, 1.0, 10] if oname == "cart_pole_sde" else [0.5, 2.0]
    lp = {"seed": 1, "stepsize": pm["stepsize"], "data_stepsize": pm["stepsize"], "num_particles": N, "num_particles_test": 4,
          "horizon": H, "obs_weights": ow, "pen_data": 1.0, "pen_weights": 0.0, "num_sample2consider": m}
    _, loss_fn, _, _ = nsde.create_model_loss_fn(pm, lp, sde_constr=cls, verbose=False)
    nn = configs.random_params(cls(pm), seed=4)
    rng = np.random.default_rng(0)
    nx, nu = pm["n_x"], pm["n_u"]
    y = (rng.normal(0, 0.5, (B, H + 1, nx)) + (np.array([0, 0, 3.0, 0]) if nx == 4 else 0)).astype(np.float32)
    u = rng.uniform(-1, 1, (B, H, nu)).astype(np.float32)
    if lanes:
        eng = loss_fn.engine
        orig = eng.loss_grad
        eng.loss_grad = lambda *a, **k: orig(*a, **dict(k, lanes=lanes))
    total, info, grads = loss_fn.value_and_grad(nn, y, u, R.PRNGKey(11))
    idx = orc.loss_step_indices(tf.PRNGKey(11), B, N, H, m)
    assert all(len(set(idx[b, n])) == m for b in range(B) for n in range(N))
    # the device indicator itself, bit for bit
    kbn = R.split_many(R.split(R.PRNGKey(11), B), N)
    mask = R.choice_mask_many(R.split_many(kbn, 3)[:, :, 1].reshape(-1, 2), H, m).cpu().numpy()  # [H, B*N]
    want_mask = np.zeros((H, B * N), np.float32)
    for b in range(B):
        for n in range(N):
            want_mask[idx[b, n], b * N + n] = 1.0
    np.testing.assert_array_equal(mask, want_mask)
    tree = _tree64(nn)
    om = orc.make_model(oname, pm, tree, torch.float64)
    dw = torch.tensor(orc.loss_noise(tf.PRNGKey(11), B, N, H, nx), dtype=torch.float64)
    Ld, _ = orc.data_loss(om, oname, torch.tensor(y, dtype=torch.float64), torch.tensor(u, dtype=torch.float64), dw, ow,
                          step_idx=idx)
    Ld.backward()
    assert abs(info["dataLoss"] - float(Ld)) < 3e-5 * abs(float(Ld))
    gmax = max(float(v.grad.abs().max()) for d in tree.values() for v in d.values())
    for mod, d in tree.items():
        for k, v in d.items():
            want = v.grad.numpy()
            assert np.max(np.abs(grads[mod][k] - want)) <= 2e-3 * np.max(np.abs(want)) + 1e-5 * gmax, (mod, k)


@pytest.mark.parametrize("kind,B,H", [("cartpole", 6, 12), ("cartpole_bb", 5, 20), ("hexa", 4, 8)])
def test_training_gradients_of_the_neural_ode_variants(cuda, kind, B, H):
    """Models without a diffusion term (the reference's `*_ode` configurations: no `diffusion_density_nn`, one
    deterministic particle): data term and parameter gradients vs the oracle's autograd."""
    from sde4mbrl_b200 import nsde, random as R
    mk, cls, oname = hp.KINDS[kind]
    pm = mk(horizon=1, num_particles=1)
    pm.pop("diffusion_density_nn", None)
    nx, nu = pm["n_x"], pm["n_u"]
    pm["noise_prior_params"] = [0.0] * nx
    ow = [9, 5, 1, 1.0, 10] if oname == "cart_pole_sde" else [1.0] * nx
    lp = {"seed": 1, "stepsize": pm["stepsize"], "data_stepsize": pm["stepsize"], "num_particles": 1, "num_particles_test": 1,
          "horizon": H, "obs_weights": ow, "pen_data": 1.0, "pen_weights": 0.0}
    _, loss_fn, _, _ = nsde.create_model_loss_fn(pm, lp, sde_constr=cls, verbose=False)
    model = cls(pm)
    assert not loss_fn.engine.noisy
    nn = configs.random_params(model, seed=4)
    rng = np.random.default_rng(0)
    if kind == "hexa":
        y = np.stack([np.stack([hp.start_state("hexa", rng) for _ in range(H + 1)]) for _ in range(B)]).astype(np.float32)
        u = rng.uniform(0.25, 0.45, (B, H, nu)).astype(np.float32)
    else:
        y = (rng.normal(0, 0.5, (B, H + 1, nx)) + np.array([0, 0, 3.0, 0])).astype(np.float32)
        u = rng.uniform(-1, 1, (B, H, nu)).astype(np.float32)
    total, info, grads = loss_fn.value_and_grad(nn, y, u, R.PRNGKey(3))
    tree = _tree64(nn)
    om = orc.make_model(oname, pm, tree, torch.float64)
    dw = torch.zeros((B, 1, H, nx), dtype=torch.float64)
    Ld, _ = orc.data_loss(om, oname, torch.tensor(y, dtype=torch.float64), torch.tensor(u, dtype=torch.float64), dw, ow)
    Ld.backward()
    assert abs(info["dataLoss"] - float(Ld)) < 3e-5 * abs(float(Ld))
    gmax = max(float(v.grad.abs().max()) for d in tree.values() for v in d.values() if v.grad is not None)
    checked = 0
    for m, d in tree.items():
        for k, v in d.items():
            if v.grad is None:
                continue
            want = v.grad.numpy()
            assert np.max(np.abs(np.asarray(grads[m][k]) - want)) <= 2e-3 * np.max(np.abs(want)) + 1e-5 * gmax, (m, k)
            checked += 1
    assert checked >= 6


@pytest.mark.parametrize("kind,cls,oname,B,N,H,nsd,m", [("cartpole", sde_models.CartPoleSDE, "cart_pole_sde", 6, 1, 10, 3, None),
                                                        ("cartpole", sde_models.CartPoleSDE, "cart_pole_sde", 5, 3, 8, 4, 5),
                                                        ("msd", sde_models.MassSpringDamper, "mass_spring_damper", 4, 2, 9, 2, None)])
def test_u_sampling_strategy_random(cuda, kind, cls, oname, B, N, H, nsd, m):
    """``u_sampling_strategy: random`` with a dataset finer than the solver step (nsde.py:43-70, :735): every
    (minibatch element, particle) draws, per solver step, which of the ``num_steps2data`` recorded controls it applies
    (``randint`` on ``rng_sample2consider``, the device draw checked bit for bit); with ``num_sample2consider`` the same
    key also subsamples the time indexes."""
    from sde4mbrl_b200 import nsde, random as R
    pm = hp.KINDS[kind][0](horizon=1, num_particles=1)
    ow = [9, 5, 1, 1.0, 10] if oname == "cart_pole_sde" else [0.5, 2.0]
    lp = {"seed": 1, "stepsize": pm["stepsize"], "data_stepsize": pm["stepsize"] / nsd, "num_particles": N, "num_particles_test": 4,
          "horizon": H, "obs_weights": ow, "pen_data": 1.0, "pen_weights": 0.0, "u_sampling_strategy": "random"}
    if m is not None:
        lp["num_sample2consider"] = m
    _, loss_fn, _, test_fn = nsde.create_model_loss_fn(pm, lp, sde_constr=cls, verbose=False)
    assert pm["num_steps2data"] == nsd and pm["u_sampling_strategy"] == "random"
    nn = configs.random_params(cls(pm), seed=4)
    rng = np.random.default_rng(0)
    nx, nu = pm["n_x"], pm["n_u"]
    Hd = H * nsd
    y = (rng.normal(0, 0.5, (B, Hd + 1, nx)) + (np.array([0, 0, 3.0, 0]) if nx == 4 else 0)).astype(np.float32)
    u = rng.uniform(-1, 1, (B, Hd, nu)).astype(np.float32)
    total, info, grads = loss_fn.value_and_grad(nn, y, u, R.PRNGKey(5))
    uidx = orc.loss_u_indices(tf.PRNGKey(5), B, N, H, nsd)
    assert uidx.min() >= 0 and uidx.max() == nsd - 1
    kbn = R.split_many(R.split(R.PRNGKey(5), B), N)
    got_idx = R.randint_many(R.split_many(kbn, 3)[:, :, 1], H, 0, nsd).cpu().numpy()
    np.testing.assert_array_equal(got_idx, uidx)
    sidx = orc.loss_step_indices(tf.PRNGKey(5), B, N, H, m) if m is not None else None
    tree = _tree64(nn)
    om = orc.make_model(oname, pm, tree, torch.float64)
    dw = torch.tensor(orc.loss_noise(tf.PRNGKey(5), B, N, H, nx), dtype=torch.float64)
    Ld, _ = orc.data_loss(om, oname, torch.tensor(y, dtype=torch.float64), torch.tensor(u, dtype=torch.float64), dw, ow,
                          num_steps2data=nsd, step_idx=sidx, u_idx=uidx)
    Ld.backward()
    assert abs(info["dataLoss"] - float(Ld)) < 3e-5 * abs(float(Ld))
    gmax = max(float(v.grad.abs().max()) for d in tree.values() for v in d.values())
    for mod, d in tree.items():
        for k, v in d.items():
            want = v.grad.numpy()
            assert np.max(np.abs(grads[mod][k] - want)) <= 2e-3 * np.max(np.abs(want)) + 1e-5 * gmax, (mod, k)
    # test_fn (nsde.py:1306-1310): one index per solver step shared by the minibatch, drawn from split(rng)[1]
    err, tinfo = test_fn(nn, y, u, R.PRNGKey(6))
    ks = tf.split(tf.PRNGKey(6), 2)
    tix = tf.randint(ks[1], (H,), 0, nsd)
    u_t = u.reshape(B, H, nsd, nu)[:, np.arange(H), tix]
    kb = tf.split(ks[0], B)
    om32 = orc.make_model(oname, pm, {mm: {k: torch.tensor(np.asarray(v)) for k, v in d.items()} for mm, d in nn.items()})
    yv = torch.tensor(y[:, ::nsd])
    ys = torch.stack([orc.sample_rollouts(om32, yv[b, 0], torch.tensor(u_t[b]), kb[b], 4)[0] for b in range(B)])  # [B, 4, H+1, nx]
    tfm = lambda z: orc.state_transform_for_loss(oname, z)
    w = torch.tensor(np.asarray(ow, np.float32))
    want_err = float((((tfm(yv) - tfm(ys.mean(dim=1))) / w) ** 2).sum(dim=(-1, -2)).mean())
    assert abs(err - want_err) < 1e-4 * abs(want_err) + 1e-6
    assert set(tinfo) == {"TestErrorData", "TestStdData"}


def test_graphed_regulariser_replays_match_the_eager_path(cuda):
    """The CUDA-graph replay of the regulariser (density_loss.GraphedRegulariser) tracks new parameters, data and keys:
    three consecutive calls with different inputs give the eager path's totals, logged scalars and gradients."""
    from sde4mbrl_b200 import nsde, random as R
    B, H, N = 6, 9, 2

    def make(graph):
        pm = configs.cartpole_model(horizon=1, num_particles=1)
        lp = {"seed": 1, "stepsize": pm["stepsize"], "data_stepsize": pm["stepsize"], "num_particles": N, "num_particles_test": 4,
              "horizon": H, "obs_weights": [9, 5, 1, 1.0, 10], "pen_data": 1.0, "pen_weights": 1e-3, "default_weights": 1.0,
              "density_loss": {"learn_mucoeff": {"type": "dnn", "init_value": 0.3, "activation_fn": "tanh", "hidden_layers": [8, 8]},
                               "mu_coeff": 10.0, "ball_radius": [0.1, 0.1, 0.4], "ball_nsamples": 6},
              "pen_grad_density": 1.0, "pen_density_scvex": 1.0, "pen_mu_type": "lin_inv", "pen_mu_coeff": 1000.0,
              "pen_scvex_mult": 0.01, "regulariser": graph}
        nn0, loss_fn, _, _ = nsde.create_model_loss_fn(pm, lp, sde_constr=sde_models.CartPoleSDE, verbose=False)
        return pm, nn0, loss_fn

    pm, nn0, f_graph = make("graph")
    _, _, f_eager = make("eager")
    _, _, f_native = make("native")
    assert (f_graph.regulariser, f_eager.regulariser, f_native.regulariser) == ("graph", "eager", "native")
    rng = np.random.default_rng(0)
    for rep in range(3):
        nn = configs.random_params(sde_models.CartPoleSDE(pm), seed=4 + rep)
        nn["cart_pole_sde/~construct_diffusion_density_nn/density/~/linear_2"]["w"] *= np.float32(3.0)
        for m in nn0:
            if "mu_coeff" in m:
                nn[m] = {k: np.array(v) * np.float32(1.0 + rep) for k, v in nn0[m].items()}
        y = (rng.normal(0, 0.5, (B, H + 1, 4)) + np.array([0, 0, 3.0, 0])).astype(np.float32)
        u = rng.uniform(-1, 1, (B, H, 1)).astype(np.float32)
        tg, ig, gg = f_graph.value_and_grad(nn, y, u, R.PRNGKey(20 + rep))
        te, ie, ge = f_eager.value_and_grad(nn, y, u, R.PRNGKey(20 + rep))
        assert abs(tg - te) <= 1e-5 * abs(te)
        for k in ie:
            assert abs(ig[k] - ie[k]) <= 1e-5 * abs(ie[k]) + 1e-12, k
        for m in ge:
            for k in ge[m]:
                assert np.max(np.abs(gg[m][k] - ge[m][k])) <= 1e-4 * np.max(np.abs(ge[m][k])) + 1e-9, (m, k)
        assert ig["densitySCVEX"] > 0
        # the library's own kernels (sde4_density_reg) on the same inputs; float32 with MUFU activations against
        # torch's float32 library kernels
        tn, inn, gn_ = f_native.value_and_grad(nn, y, u, R.PRNGKey(20 + rep))
        assert abs(tn - te) <= 5e-5 * abs(te)
        for k in ie:
            assert abs(inn[k] - ie[k]) <= 3e-4 * abs(ie[k]) + 1e-12, k
        gmax = max(float(np.max(np.abs(v))) for d in ge.values() for v in d.values())
        for m in ge:
            for k in ge[m]:
                assert np.max(np.abs(gn_[m][k] - ge[m][k])) <= 3e-3 * np.max(np.abs(ge[m][k])) + 1e-5 * gmax, (m, k)

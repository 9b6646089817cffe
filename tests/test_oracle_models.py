"""CPU checks of the oracle itself: golden regression, physical sanity on the reference's shipped
weights, float64-autograd vs finite differences, and the C restatement vs the torch oracle."""
import glob
import os

import numpy as np
import pytest
import torch

import helpers as hp
from oracle import nsde_oracle as orc
from oracle import threefry as tf
from sde4mbrl_b200 import configs, costs, io, sde_models

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "golden_*.npz"))))
def test_oracle_reproduces_golden(name):
    g = np.load(os.path.join(GOLD, name))
    kind = name.split("_")[1]
    H, N, seed, mode = int(g["H"]), int(g["N"]), int(g["seed"]), int(g["rng_mode"])
    pm, model, nn = hp.build(kind, H, N, seed=seed)
    om = hp.oracle_model(kind, pm, nn)
    x, dw = orc.sample_rollouts(om, g["y0"], g["u"], g["key"], N, mode=mode)
    np.testing.assert_array_equal(dw.numpy(), g["dw"])  # threefry noise is bit-reproducible
    assert hp.rel_err(x.numpy(), g["xevol"]) < 1e-6


def test_hover_equilibrium_with_shipped_hexa_weights():
    """Reference fixture sanity: the learned hexacopter hovers at u = 0.33."""
    nn, nominal = io.load_npz_fixture(os.path.join(GOLD, "ref_weights_hexa_ahg_v4.npz"))
    om = orc.make_model("sde_rotor_model", nominal, nn, torch.float64)
    x = torch.tensor(configs.hover_state(), dtype=torch.float64)[None]
    f = om.compositional_drift(x, torch.full((1, 6), 0.33, dtype=torch.float64))[0].numpy()
    assert np.all(np.abs(f[:3]) < 1e-9)
    assert np.max(np.abs(f[3:6])) < 0.5      # |thrust/m - g| small: 6*(9.17*0.33+0.786)/2.33 ~ 9.8
    assert np.all(f[6:10] == 0) and np.max(np.abs(f[10:])) < 1.0  # learned trim moments are small
    sig = om.diffusion(x, torch.zeros(1, 6, dtype=torch.float64))[0].numpy()
    prior = np.array(nominal["noise_prior_params"])
    assert np.all(sig <= prior + 1e-12) and np.all(sig[6:10] == 0)


def test_autograd_gradient_matches_finite_differences():
    kind, H, N = "cartpole", 12, 3
    rng = np.random.default_rng(0)
    pm, model, nn = hp.build(kind, H, N, seed=5)
    om = hp.oracle_model(kind, pm, nn, torch.float64)
    y0 = torch.tensor(hp.start_state(kind, rng), dtype=torch.float64)
    dw = torch.tensor(rng.standard_normal((N, H, 4)))
    cp = configs.cartpole_mpc()["cost_params"]
    xref = torch.zeros(H + 1, 4, dtype=torch.float64)
    f = lambda o: orc.compute_cost(om, y0, o, dw, hp.stage_cost_fn(kind, cp), xref)[0]
    o = torch.tensor(hp.controls(kind, H, rng), dtype=torch.float64, requires_grad=True)
    (g,) = torch.autograd.grad(f(o), o)
    for t in (0, 5, 11):
        e = torch.zeros_like(o)
        e[t, 0] = 1e-6
        fd = (f(o.detach() + e) - f(o.detach() - e)) / 2e-6
        assert abs(float(fd) - float(g[t, 0])) < 1e-5 * max(1.0, abs(float(g[t, 0])))


@pytest.mark.parametrize("kind,mode", [("hexa", 0), ("hexa16", 1)])
def test_c_oracle_matches_torch_oracle(kind, mode):
    from oracle import c_oracle
    H, N = 30, 21
    rng = np.random.default_rng(2)
    pm, model, nn = hp.build(kind, H, N, seed=8)
    desc = model.describe()
    packed = model.pack(nn, desc)
    cp = configs.hexa_mpc()["cost_params"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    y0, opt = hp.start_state(kind, rng), hp.controls(kind, H, rng)
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    ts = orc.compute_timesteps(pm)
    key = tf.PRNGKey(77)
    c, g, xt = c_oracle.rotor_cost_grad(desc, stage.desc, packed, y0, opt, ts, xref, N, key=key, rng_mode=mode,
                                        want_xtraj=True)
    dw = tf.rollout_noise(key, N, H, 13, mode)
    om = hp.oracle_model(kind, pm, nn, torch.float64)
    o = torch.tensor(opt, dtype=torch.float64, requires_grad=True)
    c64, xtr = orc.compute_cost(om, torch.tensor(y0, dtype=torch.float64), o, torch.tensor(dw, dtype=torch.float64),
                                hp.stage_cost_fn(kind, cp), torch.tensor(xref, dtype=torch.float64))
    c64 = c64 + orc.u_slew_cost(o, ts, cp, 6)
    (g64,) = torch.autograd.grad(c64, o)
    assert abs(c - c64.item()) < 1e-5 * abs(c64.item())
    assert hp.rel_err(g, g64.numpy()) < 1e-4
    assert hp.rel_err(xt, xtr[..., :13].detach().numpy()) < 2e-5


def test_shipped_weights_pack_for_every_fixture():
    for f, cls in (("ref_weights_hexa_ahg_v4.npz", sde_models.SDERotorModel),
                   ("ref_weights_iris_sitl.npz", sde_models.SDERotorModel),
                   ("ref_weights_cartpole_bb_learned_si.npz", sde_models.CartPoleSDE)):
        nn, nominal = io.load_npz_fixture(os.path.join(GOLD, f))
        m = cls(nominal)
        packed = m.pack(nn)
        assert packed.dtype == np.float32 and np.isfinite(packed).all() and np.abs(packed).sum() > 0


def test_density_regulariser_oracle_against_definitions():
    """oracle.density_loss_terms vs the definitions (nsde.py:632-652) evaluated by hand for one (b, n, t):
    finite-difference gradient of sigmoid(density_net - 7), the ball drawn from the documented key chain, the hinge."""
    pm = configs.cartpole_model(horizon=1, num_particles=1)
    nn = configs.random_params(sde_models.CartPoleSDE(pm), seed=2)
    nn["cart_pole_sde/~construct_diffusion_density_nn/density/~/linear_2"]["w"] *= np.float32(3.0)
    tree = {m: {k: torch.tensor(np.asarray(v), dtype=torch.float64) for k, v in d.items()} for m, d in nn.items()}
    om = orc.make_model("cart_pole_sde", pm, tree, torch.float64)
    dl = {"learn_mucoeff": {"type": "constant", "quad_approx": False}, "mu_coeff": 10.0, "ball_radius": [0.1, 0.1, 0.4],
          "ball_nsamples": 5}
    rng = np.random.default_rng(0)
    B, N, Hd = 2, 2, 3
    y = torch.tensor(rng.normal(0, 0.5, (B, Hd + 1, 4)) + np.array([0, 0, 3.0, 0]))
    u = torch.tensor(rng.uniform(-1, 1, (B, Hd, 1)))
    key = tf.PRNGKey(9)
    gn, sc, mu, den = orc.density_loss_terms(om, y, u, key, N, dl)
    assert tuple(gn.shape) == (B, N) and float(mu[0, 0]) == 10.0 and bool((sc >= 0).all())
    # by hand for b = 1, n = 1
    b, n = 1, 1
    f = lambda z: float(orc.density_function(om, torch.tensor(z)))
    gsum, ssum = 0.0, 0.0
    kt = tf.split(tf.split(tf.split(tf.split(key, B)[b], N)[n], 3)[2], Hd)
    for t in range(Hd):
        z = om.noise_aug_state_ctrl(y[b, t], u[b, t]).numpy()
        g = np.array([(f(z + 1e-6 * e) - f(z - 1e-6 * e)) / 2e-6 for e in np.eye(3)])
        gsum += (g ** 2).sum()
        ball = tf.normal(tf.split(kt[t])[1], (5, 3)).astype(np.float64) * np.array([0.1, 0.1, 0.4])
        for j in range(5):
            cond = f(z + ball[j]) - f(z) - g @ ball[j] - 0.5 * 10.0 * (ball[j] ** 2).sum()
            ssum += min(cond, 0.0) ** 2
    assert abs(float(gn[b, n]) - gsum / Hd) <= 1e-6 * gsum / Hd
    assert abs(float(sc[b, n]) - ssum) <= 1e-6 * ssum + 1e-15
    # both particles see the same data points, different balls
    assert float(gn[b, 0]) == float(gn[b, 1]) and float(sc[b, 0]) != float(sc[b, 1])

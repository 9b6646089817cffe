"""K1 parity: CUDA rollouts vs the CPU oracle on identical inputs.

Tolerances (float32, stated per the north star): with IDENTICAL noise the trajectories must agree
to 2e-5 relative to the trajectory magnitude over <= 50 steps and 1e-4 over 100 steps (chaotic
amplification of 1-ulp differences in tanh / exp / rsqrt); with in-kernel threefry noise the same
bounds hold because the uint32 stream is bit exact and the normals agree to a few ulp."""
import numpy as np
import pytest
import torch

import helpers as hp
from oracle import nsde_oracle as orc
from oracle import threefry as tf

pytestmark = pytest.mark.gpu

CASES = [("msd", 100, 200), ("cartpole", 50, 96), ("cartpole_bb", 30, 33), ("hexa", 100, 64), ("hexa16", 20, 40)]


@pytest.mark.parametrize("kind,H,N", CASES)
def test_rollout_given_noise(cuda, kind, H, N):
    from sde4mbrl_b200.engine import Engine
    rng = np.random.default_rng(1)
    pm, model, nn = hp.build(kind, H, N, seed=3)
    om = hp.oracle_model(kind, pm, nn)
    y0 = hp.start_state(kind, rng)
    u = hp.controls(kind, H, rng)
    dw = rng.standard_normal((N, H, model.n_x)).astype(np.float32)
    want, _ = orc.sample_rollouts(om, y0, u, None, N, noise=dw)
    eng = Engine(model)
    xbuf = eng.rollout(nn, y0[None], u, orc.compute_timesteps(pm), N=N, noise=hp.to_native_noise(dw))
    got = hp.from_native_x(xbuf, model.n_x)
    tol = 1e-4 if H > 50 else 2e-5
    assert hp.rel_err(got, want.numpy()) < tol
    # rotor: quaternion stays on the unit sphere (projection_fn)
    if kind.startswith("hexa"):
        assert np.max(np.abs(np.linalg.norm(got[:, 1:, 6:10], axis=-1) - 1.0)) < 5e-7


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("kind,H,N", [("msd", 100, 200), ("cartpole", 50, 64), ("hexa", 100, 37)])
def test_rollout_threefry_noise(cuda, kind, H, N, mode):
    from sde4mbrl_b200.engine import Engine
    rng = np.random.default_rng(2)
    pm, model, nn = hp.build(kind, H, N, seed=4)
    om = hp.oracle_model(kind, pm, nn)
    y0 = hp.start_state(kind, rng)
    u = hp.controls(kind, H, rng)
    key = tf.PRNGKey(2050)
    want, _ = orc.sample_rollouts(om, y0, u, key, N, mode=mode)
    eng = Engine(model, rng_mode=mode)
    xbuf = eng.rollout(nn, y0[None], u, orc.compute_timesteps(pm), key=key, N=N)
    got = hp.from_native_x(xbuf, model.n_x)
    assert hp.rel_err(got, want.numpy()) < (1e-4 if H > 50 else 2e-5)


def test_rollout_many_problems(cuda):
    """P start states x N particles, one key per problem (the batched generalisation used by the
    simulator / training paths, nsde.py:1322-1323)."""
    from sde4mbrl_b200.engine import Engine
    rng = np.random.default_rng(5)
    kind, H, N, P = "cartpole", 30, 3, 7
    pm, model, nn = hp.build(kind, H, N, seed=6)
    om = hp.oracle_model(kind, pm, nn)
    y0 = np.stack([hp.start_state(kind, rng) for _ in range(P)])
    u = np.stack([hp.controls(kind, H, rng) for _ in range(P)])
    keys = tf.split(tf.PRNGKey(1), P)
    eng = Engine(model)
    xbuf = eng.rollout(nn, y0, u, orc.compute_timesteps(pm), key=keys, N=N)
    got = hp.from_native_x(xbuf, 4).reshape(P, N, H + 1, 4)
    for p in range(P):
        want, _ = orc.sample_rollouts(om, y0[p], u[p], keys[p], N)
        assert hp.rel_err(got[p], want.numpy()) < 2e-5


def test_ode_mode_is_deterministic(cuda):
    from sde4mbrl_b200.engine import Engine
    from sde4mbrl_b200 import _lib as L
    rng = np.random.default_rng(7)
    pm, model, nn = hp.build("hexa", 20, 8, seed=1)
    eng = Engine(model)
    y0 = hp.start_state("hexa", rng)
    u = hp.controls("hexa", 20, rng)
    xbuf = eng.rollout(nn, y0[None], u, orc.compute_timesteps(pm), N=8, noise_mode=L.NOISE_NONE)
    got = hp.from_native_x(xbuf, 13)
    assert np.all(got == got[:1])
    pm0 = dict(pm, noise_prior_params=[0.0] * 13)
    om = hp.oracle_model("hexa", pm0, nn)
    want, _ = orc.sample_rollouts(om, y0, u, None, 1, noise=np.zeros((1, 20, 13), np.float32))
    assert hp.rel_err(got[:1], want.numpy()) < 2e-5

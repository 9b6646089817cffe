"""Tensor-core forward rollout (csrc/kernels_tc.cuh: the [32, 32] hidden layers as tcgen05 3xTF32 products, accumulator
in tensor memory) against the oracle and against the FP32 one-thread-per-sample kernel, incl. a ragged last tile, given
and in-kernel noise, both initialisation scales, and the simulator-summary instantiation."""
import numpy as np
import pytest
import torch

import helpers as hp
from oracle import nsde_oracle as orc
from oracle import threefry as tf
from sde4mbrl_b200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scale", ["fan_in", "yaml"])
@pytest.mark.parametrize("kind,H,N", [("cartpole", 30, 200), ("hexa", 25, 150)])
def test_tc_rollout_given_noise(cuda, kind, H, N, scale):
    from sde4mbrl_b200.engine import Engine
    rng = np.random.default_rng(1)
    pm, model, nn = hp.build(kind, H, N, seed=3, scale=scale)
    om = hp.oracle_model(kind, pm, nn)
    y0 = hp.start_state(kind, rng)
    u = hp.controls(kind, H, rng)
    dw = rng.standard_normal((N, H, model.n_x)).astype(np.float32)
    want, _ = orc.sample_rollouts(om, y0, u, None, N, noise=dw)
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    got = hp.from_native_x(eng.rollout(nn, y0[None], u, ts, N=N, noise=hp.to_native_noise(dw), lanes=L.LANES_TC), model.n_x)
    assert hp.rel_err(got, want.numpy()) < 2e-5
    fp32 = hp.from_native_x(eng.rollout(nn, y0[None], u, ts, N=N, noise=hp.to_native_noise(dw), lanes=1), model.n_x)
    assert hp.rel_err(got, fp32) < 1e-5


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("kind,H,N", [("cartpole", 50, 300), ("hexa", 40, 129)])
def test_tc_rollout_threefry_noise(cuda, kind, H, N, mode):
    from sde4mbrl_b200.engine import Engine
    rng = np.random.default_rng(2)
    pm, model, nn = hp.build(kind, H, N, seed=4)
    om = hp.oracle_model(kind, pm, nn)
    y0 = hp.start_state(kind, rng)
    u = hp.controls(kind, H, rng)
    key = tf.PRNGKey(2050)
    want, _ = orc.sample_rollouts(om, y0, u, key, N, mode=mode)
    eng = Engine(model, rng_mode=mode)
    got = hp.from_native_x(eng.rollout(nn, y0[None], u, orc.compute_timesteps(pm), key=key, N=N, lanes=L.LANES_TC), model.n_x)
    assert hp.rel_err(got, want.numpy()) < 2e-5


def test_tc_rollout_many_start_states_and_auto_dispatch(cuda):
    """Many start states with their own controls and keys (the simulator configuration): the automatic choice (>= two
    tiles per SM -> tensor-core kernel) and the explicit FP32 kernel agree; architectures without [32, 32] networks and
    the Stratonovich solvers refuse the explicit request."""
    from sde4mbrl_b200 import random as R
    from sde4mbrl_b200._lib import Sde4Error
    from sde4mbrl_b200.engine import Engine
    H, N = 12, 1
    P = 2 * 128 * torch.cuda.get_device_properties(0).multi_processor_count + 77
    rng = np.random.default_rng(5)
    pm, model, nn = hp.build("cartpole", H, N, seed=2)
    y0 = (rng.uniform(-1, 1, (P, 4)) * np.array([2, 3, np.pi, 6])).astype(np.float32)
    u = rng.uniform(-1, 1, (P, H, 1)).astype(np.float32)
    keys = R.split(R.PRNGKey(7), P)
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    auto = eng.rollout(nn, y0, u, ts, key=keys, N=N).clone()
    tcr = eng.rollout(nn, y0, u, ts, key=keys, N=N, lanes=L.LANES_TC).clone()
    fp32 = eng.rollout(nn, y0, u, ts, key=keys, N=N, lanes=1).clone()
    assert torch.equal(auto, tcr)  # same kernel
    assert float((tcr - fp32).abs().max()) < 2e-5 * float(fp32.abs().max())
    with pytest.raises(Sde4Error):
        eng.rollout(nn, y0[:4], u[:4], ts, key=keys[:4], N=N, lanes=L.LANES_TC, solver="stratonovich_heun")
    pm2, model2, nn2 = hp.build("msd", H, N, seed=2)
    with pytest.raises(Sde4Error):
        Engine(model2).rollout(nn2, np.zeros((1, 2), np.float32), np.zeros((H, 1), np.float32), orc.compute_timesteps(pm2),
                               key=R.PRNGKey(1), N=4, lanes=L.LANES_TC)


def test_tc_simulator_summary(cuda):
    """sde4_sim_rollout through the tensor-core instantiation (automatic for large batches) against the FP32 one."""
    from sde4mbrl_b200 import random as R
    from sde4mbrl_b200.engine import Engine
    H, N = 8, 4
    P = 128 * torch.cuda.get_device_properties(0).multi_processor_count
    rng = np.random.default_rng(6)
    pm, model, nn = hp.build("cartpole", H, N, seed=2)
    y0 = (rng.uniform(-1, 1, (P, 4)) * np.array([2, 3, np.pi, 6])).astype(np.float32)
    u = rng.uniform(-1, 1, (P, H, 1)).astype(np.float32)
    keys = R.split(R.PRNGKey(7), P)
    eng = Engine(model)
    ts = orc.compute_timesteps(pm)
    a = eng.sim_rollout(nn, y0, u, ts, keys, N=N, reward="cartpole_swingup", want_std=True, want_step=True)
    b = eng.sim_rollout(nn, y0, u, ts, keys, N=N, reward="cartpole_swingup", want_std=True, want_step=True, lanes=1)
    for k in ("mean", "std", "reward_mean", "reward_of_mean"):
        np.testing.assert_allclose(a[k].cpu().numpy(), b[k].cpu().numpy(), rtol=3e-5, atol=3e-5)
    assert torch.equal(a["done_of_mean"], b["done_of_mean"])

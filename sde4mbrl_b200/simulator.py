"""The nSDE as an MBRL simulator (SURVEY note N1, 8(f) rank 3).

Reference: ``cartpole_sde_gym`` (``sde4mbrlExamples/cartpole/cartpole_sde.py:197-267``) wraps the
one-step sampler as a gym env step: ``next_state = mean over particles``, ``std over particles``
(``_my_pred_fn``, :219-223), with ``modified_params = {'horizon': 1, 'num_particles': n, 'stepsize': tau}``.
``predict`` is the batched generalisation over many start states (the vmap the reference applies in
``create_model_loss_fn.test_fn``, ``nsde.py:1322-1323``) in the shape of mbrl-lib's ``ModelEnv`` step
(``mbrlLibUtils/modified_model_env.py:16-18``): one launch advances P states by H steps.
"""
import numpy as np
import torch

from . import random as R
from .engine import Engine, as_f32, as_key
from .io import update_params
from .nsde import compute_timesteps


class SDESimulator:
    def __init__(self, sde_constr, nominal, nn_params, num_particles=1, tau=None, horizon=1):
        mod = {"horizon": horizon, "num_particles": num_particles}
        if tau is not None:
            mod["stepsize"] = tau
        self.params_model = update_params(nominal, mod)
        for k in ("num_short_dt", "short_step_dt", "long_step_dt"):
            self.params_model.pop(k, None)
        self.model = sde_constr(self.params_model)
        self.engine = Engine(self.model, R.rng_mode())
        self.pk = self.engine.packed(nn_params)
        self.N, self.H = num_particles, horizon
        self.ts = compute_timesteps(self.params_model)

    def pred_fn(self, x, u, rng):
        """``_my_pred_fn``: one start state -> (mean, std) over particles of the last state
        (population std, as ``jnp.std``)."""
        u = as_f32(u).reshape(self.H, -1)
        self.engine.rng_mode = R.rng_mode()
        xbuf = self.engine.rollout(self.pk, as_f32(x).reshape(1, -1), u, self.ts, key=as_key(rng), N=self.N)
        last = xbuf[self.H]  # [n_x, N]
        return last.mean(dim=1), last.std(dim=1, unbiased=False)

    def predict(self, obs, act, rng, return_particles=False):
        """P start states ``obs[P, n_x]``, controls ``act[P, H, n_u]`` (or ``[P, n_u]`` for H = 1),
        ``rng`` one key per state ``[P, 2]`` or a single key (then split on the device).
        Returns next states ``[P, n_x]`` (mean over particles) -- or all particles ``[P, N, H+1, n_x]``."""
        obs = as_f32(obs)
        P = obs.shape[0]
        act = as_f32(act)
        if act.dim() == 2:
            act = act.reshape(P, 1, -1)
        keys = as_key(rng)
        if keys.dim() == 1:
            keys = R.split(keys, P)
        self.engine.rng_mode = R.rng_mode()
        xbuf = self.engine.rollout(self.pk, obs, act, self.ts, key=keys, N=self.N)
        n_x = self.engine.n_x
        if return_particles:
            return xbuf.view(self.H + 1, n_x, P, self.N).permute(2, 3, 0, 1)
        return xbuf[self.H].view(n_x, P, self.N).mean(dim=2).t()

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
        r = self.engine.sim_rollout(self.pk, as_f32(x).reshape(1, -1), u, self.ts, as_key(rng), N=self.N, want_std=True)
        return r["mean"][0], r["std"][0]  # reduced on the device by k_sim_reduce; no trajectory is written

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
        n_x = self.engine.n_x
        if return_particles:
            xbuf = self.engine.rollout(self.pk, obs, act, self.ts, key=keys, N=self.N)
            return xbuf.view(self.H + 1, n_x, P, self.N).permute(2, 3, 0, 1)
        # fused path: only the last states leave the kernel ([n_x, P*N] instead of [H+1, n_x, P*N]), the particle mean
        # is taken by the reduction kernel
        return self.engine.sim_rollout(self.pk, obs, act, self.ts, keys, N=self.N)["mean"]

    def step_summary(self, obs, act, rng, reward=None):
        """One fused call for mbrl's ``ModelEnv.step``: next states (particle mean) and, for a reward the kernels know
        (``'cartpole_swingup'``), reward and termination of the mean next state."""
        obs = as_f32(obs)
        P = obs.shape[0]
        act = as_f32(act)
        if act.dim() == 2:
            act = act.reshape(P, 1, -1)
        keys = as_key(rng)
        if keys.dim() == 1:
            keys = R.split(keys, P)
        self.engine.rng_mode = R.rng_mode()
        return self.engine.sim_rollout(self.pk, obs, act, self.ts, keys, N=self.N, reward=reward, want_step=reward is not None)

    def sequence_rewards(self, obs, act, rng, reward):
        """Per start state / control sequence: the per-particle reward summed over the steps up to termination, averaged
        over the particles, accumulated INSIDE the rollout kernel (mbrl ``evaluate_action_sequences``)."""
        return self.step_summary(obs, act, rng, reward=reward)["reward_mean"]


# ---------------------------------------------------------------------------------------------
# mbrl-lib ``ModelEnv`` contract (mbrlLibUtils/modified_model_env.py:12-57) around ``SDESimulator``
# ---------------------------------------------------------------------------------------------
def cartpole_swingup_reward(act, next_obs):
    """``mbrlLibUtils/reward_functions.py:3-9`` on observations [x, xdot, sin th, cos th, thdot].  Kept verbatim:
    the reference takes ``cos`` of observation 3, which already is ``cos th`` (sic)."""
    assert next_obs.dim() == act.dim() == 2
    return (torch.cos(next_obs[:, 3]) - 0.1 * torch.abs(next_obs[:, 0])).float().view(-1, 1)


def cartpole_swingup_termination(act, next_obs):
    """``mbrlLibUtils/termination_functions.py:3-14``: done when |x| >= 120."""
    assert next_obs.dim() == 2
    x = next_obs[:, 0]
    return (~((x > -120.0) & (x < 120.0)))[:, None]


def cartpole_obs(state):
    """``CartPoleSDEEnv.get_obs`` (cartpole_sde.py:255-257): [x, xdot, sin th, cos th, thdot]."""
    return torch.stack((state[..., 0], state[..., 1], torch.sin(state[..., 2]), torch.cos(state[..., 2]), state[..., 3]), dim=-1)


class SDEModelEnv:
    """``reset(initial_states) -> model_state`` / ``step(actions, model_state) -> (next_obs, rewards, dones,
    model_state)`` / ``evaluate_action_sequences`` with the nSDE as the dynamics model.  States are the SDE
    states; ``obs_fn`` maps them to what the reward / termination functions expect."""

    def __init__(self, simulator, termination_fn, reward_fn, obs_fn=None, seed=0):
        self.sim = simulator
        self.termination_fn, self.reward_fn = termination_fn, reward_fn
        self.obs_fn = obs_fn if obs_fn is not None else (lambda s: s)
        self._rng = R.PRNGKey(seed)
        # the reward / termination pair the rollout kernel evaluates itself (cartpole swing-up on cartpole_obs)
        self._fused_kind = "cartpole_swingup" if (reward_fn is cartpole_swingup_reward and termination_fn is cartpole_swingup_termination
                                                  and obs_fn is cartpole_obs and simulator.engine.n_x == 4) else None

    def _next_key(self):
        ks = R.split(self._rng)
        self._rng = ks[0]
        return ks[1]

    def reset(self, initial_states):
        return {"state": as_f32(initial_states).reshape(-1, self.sim.engine.n_x).clone()}

    def step(self, actions, model_state):
        """One simulator step for every state of the batch (mean over the particles, as ``_my_pred_fn``)."""
        actions = as_f32(actions).reshape(model_state["state"].shape[0], -1)
        if self._fused_kind is not None:  # reward / termination of the mean next state come out of the same launches
            r = self.sim.step_summary(model_state["state"], actions, self._next_key(), reward=self._fused_kind)
            nxt = r["mean"]
            return self.obs_fn(nxt), r["reward_of_mean"].view(-1, 1), r["done_of_mean"].bool().view(-1, 1), {"state": nxt}
        nxt = self.sim.predict(model_state["state"], actions, self._next_key())
        obs = self.obs_fn(nxt)
        rewards = self.reward_fn(actions, obs)
        dones = self.termination_fn(actions, obs)
        return obs, rewards, dones, {"state": nxt}

    def evaluate_action_sequences(self, action_sequences, initial_state, num_particles):
        """Total reward of every candidate sequence ``[B, H, n_u]`` from one start state, averaged over
        ``num_particles`` particles -- ONE rollout launch (B problems x particles x H steps) instead of H model
        steps (mbrl ``ModelEnv.evaluate_action_sequences``); rewards stop accumulating after termination."""
        seqs = as_f32(action_sequences)
        B, Hs, _ = seqs.shape
        sim = self.sim
        if (sim.H, sim.N) != (Hs, num_particles):
            sim = SDESimulator(type(self.sim.model), self.sim.params_model, self.sim.pk, num_particles=num_particles,
                               horizon=Hs)
            self._seq_sim = sim
        x0 = as_f32(initial_state).reshape(1, -1).expand(B, -1)
        if self._fused_kind is not None:  # rewards accumulated in the kernel: nothing of the trajectories is written
            return sim.sequence_rewards(x0, seqs, self._next_key(), self._fused_kind)
        traj = sim.predict(x0, seqs, self._next_key(), return_particles=True)  # [B, N, H+1, n_x]
        obs = self.obs_fn(traj[:, :, 1:, :]).reshape(B * num_particles * Hs, -1)
        act = seqs[:, None, :, :].expand(B, num_particles, Hs, seqs.shape[-1]).reshape(B * num_particles * Hs, -1)
        rew = self.reward_fn(act, obs).view(B, num_particles, Hs)
        done = self.termination_fn(act, obs).view(B, num_particles, Hs)
        alive = (torch.cumsum(done.to(torch.int32), dim=2) - done.to(torch.int32)) == 0  # the terminal step still counts
        return (rew * alive).sum(dim=2).mean(dim=1)

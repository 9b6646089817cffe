"""Drop-in mirror of the sampling / cost factories of ``sde4mbrl/nsde.py``.

Same function names, argument meaning and error behaviour as the reference; the closures they
return launch the B200 kernels through the C-ABI (``libsde4b200.so``) instead of tracing JAX.
Arrays in are numpy / torch; arrays out are torch CUDA tensors whose *logical* shapes equal the
reference's (``[N, H+1, n_x]`` ...) and which are strided views of the kernels' sample-minor
buffers.

Differences that cannot be hidden (all raise instead of silently changing behaviour):
  * ``sde_constr`` must be one of the descriptor classes in ``sde4mbrl_b200.sde_models``;
  * a callable ``us(y)`` (``sde_solver.py:292``) is accepted by the sampling functions only, receives the batch of
    particle observations as a CUDA tensor, and is stepped from the host (one launch per step); the cost / gradient
    factories need control arrays;
  * ``extra_args`` / ``extra_dyn_args`` must be ``None`` (no shipped model uses them);
  * ``_cost_fn`` must be a ``costs.StageCost`` (the shipped cost families), not a Python callable.
"""
import copy

import numpy as np
import torch

from . import _lib as L
from . import costs as _costs
from . import density_loss as _density
from . import random as _random
from ._lib import Sde4Error
from .engine import Engine, as_f32, as_key, is_host, host_f32, host_key
from .sde_models import ControlledSDE


def compute_timesteps(params):
    """``sde4mbrl/nsde.py:15-40``."""
    horizon = params["horizon"]
    stepsize = params["stepsize"]
    num_short_dt = params.get("num_short_dt", horizon)
    assert num_short_dt <= horizon, "The number of short dt is greater than horizon"
    num_long_dt = horizon - num_short_dt
    short_step_dt = params.get("short_step_dt", stepsize)
    long_step_dt = params.get("long_step_dt", stepsize)
    return np.array([short_step_dt] * num_short_dt + [long_step_dt] * num_long_dt, dtype=np.float32)


def initialize_problem_constraints(n_x, n_u, params_model):
    """``sde4mbrl/nsde.py:75-130`` (numpy instead of jnp)."""
    has_ubound = "input_constr" in params_model
    input_lb, input_ub = None, None
    if has_ubound:
        print("Found input bound constraints...\n")
        input_lb = [-np.inf for _ in range(n_u)]
        input_ub = [np.inf for _ in range(n_u)]
        input_dict = params_model["input_constr"]
        assert len(input_dict["input_id"]) <= n_u and \
            len(input_dict["input_id"]) == len(input_dict["input_bound"]), \
            "The number of constrained inputs identifier does not match the number of bounds"
        for idx, (u_lb, u_ub) in zip(input_dict["input_id"], input_dict["input_bound"]):
            input_lb[idx], input_ub[idx] = u_lb, u_ub
        input_lb = np.array(input_lb, np.float32)
        input_ub = np.array(input_ub, np.float32)
    has_xbound = "state_constr" in params_model
    weight_constr = 1
    slack_scaling = 1
    slack_proximal, state_idx, penalty_coeff, state_lb, state_ub = False, None, None, None, None
    if has_xbound:
        weight_constr = params_model["state_constr"].get("constr_pen", 1)
        print("Found states bound constraints...\n")
        slack_dict = params_model["state_constr"]
        assert len(slack_dict["state_id"]) <= n_x and \
            len(slack_dict["state_id"]) == len(slack_dict["state_bound"]), \
            "The number of the constrained states identifier does not match the number of bounds"
        for _idx in slack_dict["state_id"]:
            assert _idx < n_x, "The state identifier is out of bound {}/{}".format(_idx, n_x)
        state_idx = np.array(slack_dict["state_id"])
        slack_proximal = slack_dict["slack_proximal"]
        penalty_coeff = slack_dict["state_penalty"]
        state_lb = np.array([x[0] for x in slack_dict["state_bound"]], np.float32)
        state_ub = np.array([x[1] for x in slack_dict["state_bound"]], np.float32)
        slack_scaling = np.array(params_model["state_constr"].get("slack_scaling", 1), np.float32) if slack_proximal else 1
        state_lb = state_lb / slack_scaling
        state_ub = state_ub / slack_scaling
    return (has_ubound, input_lb, input_ub), \
        (has_xbound, slack_proximal, state_idx, penalty_coeff, state_lb, state_ub, weight_constr, slack_scaling)


def _check_constr(sde_constr):
    if not (isinstance(sde_constr, type) and issubclass(sde_constr, ControlledSDE)) or sde_constr is ControlledSDE:
        raise Sde4Error("sde_constr must be MassSpringDamper, CartPoleSDE or SDERotorModel from "
                        "sde4mbrl_b200.sde_models (arbitrary Python drift/diffusion callables cannot run in a CUDA kernel)")


def _controls(u, H, n_u):
    if callable(u):
        raise Sde4Error("controls given as a function of the observation are not supported by the CUDA path")
    u = as_f32(u)
    if u.dim() == 1:
        u = u[None]  # sde_solver.py:264-265
    assert u.shape[-2] == H, "Dimension mismatch on the input of the sde solver"
    return u


def _make_multi_sampling(engine, params_model, num_samples, squeeze_single):
    n_x, n_u = engine.n_x, engine.n_u

    def multi_sampling(_nn_params, y, u, rng, extra_args=None):
        if extra_args is not None:
            raise Sde4Error("extra_args must be None (no shipped model uses extra drift/diffusion arguments)")
        time_steps = compute_timesteps(params_model)
        time_steps.setflags(write=False)
        H = time_steps.shape[0]
        y = as_f32(y)
        batched = y.dim() == 2  # extension: a leading batch of start states (vmap in nsde.py:1322-1323)
        assert y.dim() in (1, 2), "Not the right dimension for the initial observation and the random key"
        if callable(u):
            return _closed_loop(_nn_params, y, u, as_key(rng), time_steps)
        u = _controls(u, H, n_u)
        key = as_key(rng)
        assert key.dim() == (2 if (batched and key.dim() == 2) else 1), "RNG must be a single key for vmapping"
        engine.rng_mode = _random.rng_mode()
        xbuf = engine.rollout(_nn_params, y, u, time_steps, key=key, N=num_samples)
        P = y.shape[0] if batched else 1
        N = num_samples
        x = xbuf.view(H + 1, n_x, P, N).permute(2, 3, 0, 1)  # [P, N, H+1, n_x] strided view
        if u.dim() == 2:
            uev = u.view(1, 1, H, n_u).expand(P, N, H, n_u)
        else:
            uev = u.view(u.shape[0], 1, H, n_u).expand(P, N, H, n_u)
        if not batched:
            x, uev = x[0], uev[0]
        if squeeze_single and num_samples == 1:
            x, uev = x.select(-3, 0), uev.select(-3, 0)
        return x, x, uev  # identity observation map: yevol is xevol

    def _closed_loop(_nn_params, y, u_fn, key, time_steps):
        """Controls given as a function of the observation, ``u_t = us(y_t)`` (sde_solver.py:292).  A Python callable
        cannot run inside the rollout kernel, so the horizon is stepped from the host: one K1 launch per step over all
        particles (each particle = one problem with its own control), with the Brownian increments drawn up front on the
        device by the reference's key chain (``split(rng, N)[p]`` nsde.py:1026 -> ``split(.)[0]``, ``normal(., (H, n_x))``
        sde_solver.py:276,283).  The reference vmaps ``us`` over particles; here it receives the batch ``y[N, n_y]`` (a
        CUDA tensor) and returns ``[N, n_u]`` (or ``[n_u]``, shared)."""
        assert y.dim() == 1 and key.dim() == 1, "closed-loop controls: one start observation and one key"
        engine.rng_mode = _random.rng_mode()
        N, H = num_samples, time_steps.shape[0]
        noise = None
        if engine.noisy:
            kb = _random.split_many(_random.split(key, N), 2)[:, 0].contiguous()
            noise = _random.normal_many(kb, (H, n_x)).permute(1, 2, 0).contiguous()  # [H, n_x, N]
        x = y[None].expand(N, n_x).contiguous()
        xs, us = [x], []
        for t in range(H):
            ut = as_f32(u_fn(x))
            if ut.dim() == 1:
                ut = ut[None].expand(N, n_u)
            assert tuple(ut.shape) == (N, n_u), "us(y) must return [N, n_u] or [n_u]"
            xb = engine.rollout(_nn_params, x, ut.reshape(N, 1, n_u).contiguous(), time_steps[t:t + 1], N=1,
                                noise=None if noise is None else noise[t:t + 1])
            x = xb[1].t().contiguous()
            xs.append(x)
            us.append(ut)
        xev, uev = torch.stack(xs, dim=1), torch.stack(us, dim=1)
        if squeeze_single and num_samples == 1:
            xev, uev = xev[0], uev[0]
        return xev, xev, uev

    multi_sampling.engine = engine
    return multi_sampling


def create_sampling_fn(params_model, sde_constr=ControlledSDE, seed=0, num_samples=None, **extra_args_sde_constr):
    """``sde4mbrl/nsde.py:987-1030``: returns ``(nn_params, multi_sampling)`` with
    ``multi_sampling(nn_params, y, u, rng, extra_args=None) -> (xevol, yevol, uevol)``,
    shapes ``[N, H+1, n_x]``, ``[N, H+1, n_y]``, ``[N, H, n_u]``."""
    _check_constr(sde_constr)
    num_samples = params_model["num_particles"] if num_samples is None else num_samples
    model = sde_constr(params_model, **extra_args_sde_constr)
    nn_params = model.init_params(seed)
    engine = Engine(model, _random.rng_mode())
    return nn_params, _make_multi_sampling(engine, params_model, num_samples, squeeze_single=False)


def create_one_step_sampling(params_model, sde_constr=ControlledSDE, seed=0, num_samples=None,
                             **extra_args_sde_constr):
    """``sde4mbrl/nsde.py:933-985``: horizon forced to 1, outputs squeezed when ``num_samples == 1``."""
    _check_constr(sde_constr)
    params_model = copy.deepcopy(params_model)
    params_model["horizon"] = 1
    params_model.pop("num_short_dt", None)
    params_model.pop("short_step_dt", None)
    params_model.pop("long_step_dt", None)
    num_samples = params_model["num_particles"] if num_samples is None else num_samples
    model = sde_constr(params_model, **extra_args_sde_constr)
    engine = Engine(model, _random.rng_mode())
    return _make_multi_sampling(engine, params_model, num_samples, squeeze_single=True)


class _CostSampling(torch.autograd.Function):
    """mean-over-particles MPC cost, differentiable in ``opt_params`` (the reference relies on
    ``jax.value_and_grad`` of this function, ``sde4mbrl/apg.py:82,224``).  The forward launch
    computes value AND gradient in one fused kernel when a gradient is required."""

    @staticmethod
    def forward(ctx, opt, fn):
        need = bool(ctx.needs_input_grad[0])
        cost, grad, xtraj = fn(opt.detach(), need)
        ctx.grad = grad
        ctx.mark_non_differentiable(xtraj)
        return cost, xtraj

    @staticmethod
    def backward(ctx, gcost, _gx):
        if ctx.grad is None:
            return None, None
        g = ctx.grad
        if g.dim() == 3:
            return g * gcost.view(-1, 1, 1), None
        return g * gcost, None


def create_online_cost_sampling_fn(params_model, params_mpc, sde_constr=ControlledSDE, seed=0,
                                   **extra_args_sde_constr):
    """``sde4mbrl/nsde.py:1341-1597``.  Returns
    ``(nn_params, multi_sampling, vmapped_prox, constr_cost, construct_opt_params)``.

    ``multi_sampling(nn_params, y, opt_params, rng, _cost_fn, _terminal_cost=None,
    extra_dyn_args=None, extra_cost_args=None, extra_cost_terminal_args=None) -> (mean_cost, xtraj)``
    where ``extra_cost_args`` is the reference trajectory ``xref[H+1, n_x]`` and ``xtraj`` is
    ``[N, H+1, n_x+1]`` (last column = stage cost, nsde.py:1524-1525).  Like the reference this
    MUTATES ``params_model`` (nsde.py:1372-1385)."""
    _check_constr(sde_constr)
    n_u = params_model["n_u"]
    n_x = params_model["n_x"]
    num_sample = params_mpc.get("num_particles", params_model.get("num_particles", 1))
    params_model["num_particles"] = num_sample
    print("Using [ N = {} ] particles for the loss".format(num_sample))
    params_model["horizon"] = params_mpc.get("horizon", params_model.get("horizon", 1))
    print("Using [ T = {} ] horizon for the loss".format(params_model["horizon"]))
    params_model["num_short_dt"] = params_mpc["num_short_dt"]
    params_model["short_step_dt"] = params_mpc["short_step_dt"]
    params_model["long_step_dt"] = params_mpc["long_step_dt"]
    time_steps = compute_timesteps(params_model)
    time_steps.setflags(write=False)  # lets the engine cache its device copy by identity
    H = time_steps.shape[0]

    (has_ubound, input_lb, input_ub), \
        (has_xbound, slack_proximal, state_idx, penalty_coeff, state_lb, state_ub, weight_constr, slack_scaling) = \
        initialize_problem_constraints(n_x, n_u, params_mpc)
    penalty_coeff = np.asarray(penalty_coeff, np.float32) if penalty_coeff is not None else None
    has_slack = bool(has_xbound and slack_proximal)
    n_slack = len(state_idx) if has_slack else 0
    opt_params_size = n_u + n_slack

    constr = None
    if has_xbound:
        ns = len(state_idx)
        constr = dict(mode=L.CONSTR_WITHPROX if slack_proximal else L.CONSTR_NOPROX, n_slack=n_slack,
                      weight=weight_constr, state_idx=[int(i) for i in state_idx], penalty=penalty_coeff,
                      state_lb=state_lb, state_ub=state_ub,
                      slack_scaling=np.broadcast_to(np.asarray(slack_scaling, np.float32), (ns,)))

    class _Box:
        """``constr_vect`` (nsde.py:1422): clip to [lb, ub] on the device the argument lives on (numpy
        arguments are moved to the GPU, as every other entry point does)."""

        def __init__(self, lb, ub, first=0):
            self.lb_np, self.ub_np = np.asarray(lb, np.float32), np.asarray(ub, np.float32)
            self.first = first  # columns before `first` pass through unclipped
            self._dev = {}

        def bounds(self, device):
            b = self._dev.get(device)
            if b is None:
                b = (torch.as_tensor(self.lb_np, device=device), torch.as_tensor(self.ub_np, device=device))
                self._dev[device] = b
            return b

        def __call__(self, u_slack):
            a = u_slack.to(torch.float32) if isinstance(u_slack, torch.Tensor) else as_f32(u_slack)
            lb, ub = self.bounds(a.device)
            if self.first:
                return torch.cat((a[..., :self.first], torch.minimum(torch.maximum(a[..., self.first:], lb), ub)), dim=-1)
            return torch.minimum(torch.maximum(a, lb), ub)

    # proximal operators (nsde.py:1435-1460); they act on [H, n_opt] (or [..., H, n_opt]) tensors
    proximal_fn = None
    constr_cost = None
    if has_ubound and (not has_xbound):
        proximal_fn = _Box(input_lb, input_ub)
    if (not has_ubound) and has_xbound and slack_proximal:
        proximal_fn = _Box(state_lb, state_ub, first=n_u)
    if has_ubound and has_xbound and (not slack_proximal):
        proximal_fn = _Box(input_lb, input_ub)
    if has_ubound and has_xbound and slack_proximal:
        proximal_fn = _Box(np.concatenate((np.asarray(input_lb, np.float32), np.asarray(state_lb, np.float32))),
                           np.concatenate((np.asarray(input_ub, np.float32), np.asarray(state_ub, np.float32))))
    if has_xbound:
        constr_cost = "constr_cost_withprox" if slack_proximal else "constr_cost_noprox"  # evaluated in-kernel

    model = sde_constr(params_model, **extra_args_sde_constr)
    nn_params = model.init_params(seed)
    engine = Engine(model, _random.rng_mode())

    def _prepare(_nn_params, y, opt_params, rng, _cost_fn, _terminal_cost, extra_dyn_args, extra_cost_args):
        """Argument checks shared by ``multi_sampling`` and ``multi_sampling.value_and_grad``.  Returns
        ``(opt, run)`` with ``run(opt, need_grad) -> (cost, grad or None, xbuf)``."""
        if extra_dyn_args is not None:
            raise Sde4Error("extra_dyn_args must be None (no shipped model uses them)")
        if not isinstance(_cost_fn, _costs.StageCost):
            raise Sde4Error("_cost_fn must be a sde4mbrl_b200.costs.StageCost (e.g. costs.one_step_cost_f(cost_params, n_u)); "
                            "arbitrary Python cost callables cannot run inside the CUDA kernel")
        if _terminal_cost is not None and not isinstance(_terminal_cost, _costs.StageCost):
            raise Sde4Error("_terminal_cost must be a costs.StageCost or None")
        if extra_cost_args is None:
            raise Sde4Error("extra_cost_args (the reference trajectory xref[H+1, n_x]) is required")
        stage = _cost_fn if constr is None else _costs.with_constraints(_cost_fn, constr, n_u)
        term = None if _terminal_cost is None else _terminal_cost.desc
        engine.rng_mode = _random.rng_mode()
        host = is_host(y) and is_host(opt_params) and is_host(rng) and is_host(extra_cost_args)
        if host:
            # every argument lives in host memory (numpy arrays / CPU tensors, as when the reference's jitted
            # function is fed numpy arrays): one packed copy each way through sde4_cost_grad_host; the cost
            # (and its gradient) come back as CPU tensors, the trajectories stay on the device
            opt = opt_params if isinstance(opt_params, torch.Tensor) else torch.from_numpy(host_f32(opt_params))
        else:
            opt = opt_params if isinstance(opt_params, torch.Tensor) else as_f32(opt_params)
        opt = opt.to(dtype=torch.float32)
        assert opt.dim() in (2, 3), "The parameters must be a two dimension array"
        assert opt.shape[-1] == opt_params_size, "Shape of the opt params do not match"
        if host:
            y_h, xref_h = host_f32(y), host_f32(extra_cost_args)

            def run(opt_h, need_grad):
                cost, grad, xtraj = engine.cost_grad_host(_nn_params, y_h, opt_h, time_steps, xref_h, stage.desc,
                                                          terminal=term, key=rng, N=num_sample, want_grad=need_grad,
                                                          want_xtraj=True)
                cost_t = torch.from_numpy(cost)
                return (cost_t if opt_h.dim() == 3 else cost_t[0]), (None if grad is None else torch.from_numpy(grad)), xtraj
        else:
            key = as_key(rng)
            assert key.dim() == 1 or opt.dim() == 3, "RNG must be a single key for vmapping"
            y_t = as_f32(y)
            xref = as_f32(extra_cost_args)
            if opt.device.type != "cuda":
                opt = opt.to(y_t.device)

            def run(opt_d, need_grad):
                cost, grad, xtraj = engine.cost_grad(_nn_params, y_t, opt_d, time_steps, xref, stage.desc, terminal=term,
                                                     key=key, N=num_sample, want_grad=need_grad, want_xtraj=True)
                return (cost if opt_d.dim() == 3 else cost[0]), grad, xtraj
        return opt, run

    def _xtraj_view(xbuf, opt):
        P = opt.shape[0] if opt.dim() == 3 else 1
        xtraj = xbuf.view(H + 1, n_x + 1, P, num_sample).permute(2, 3, 0, 1)
        return xtraj[0] if opt.dim() == 2 else xtraj

    def multi_sampling(_nn_params, y, opt_params, rng, _cost_fn, _terminal_cost=None, extra_dyn_args=None,
                       extra_cost_args=None, extra_cost_terminal_args=None):
        opt, run = _prepare(_nn_params, y, opt_params, rng, _cost_fn, _terminal_cost, extra_dyn_args, extra_cost_args)
        cost, xbuf = _CostSampling.apply(opt, run)
        return cost, _xtraj_view(xbuf, opt)

    def value_and_grad(_nn_params, y, opt_params, rng, _cost_fn, _terminal_cost=None, extra_dyn_args=None,
                       extra_cost_args=None, extra_cost_terminal_args=None):
        """``jax.value_and_grad(multi_sampling, argnums=2, has_aux=True)`` (how ``apg.py:82,224`` obtains the
        gradient): ``((mean_cost, xtraj), grad)`` from ONE fused launch, without building an autograd graph."""
        if multi_sampling.particle_reducer is not None:
            return _value_and_grad_sharded(multi_sampling.particle_reducer, _nn_params, y, opt_params, rng, _cost_fn,
                                           _terminal_cost, extra_dyn_args, extra_cost_args)
        if is_host(y) and is_host(opt_params) and is_host(rng) and is_host(extra_cost_args):
            return _value_and_grad_host(_nn_params, y, opt_params, rng, _cost_fn, _terminal_cost, extra_dyn_args, extra_cost_args)
        opt, run = _prepare(_nn_params, y, opt_params, rng, _cost_fn, _terminal_cost, extra_dyn_args, extra_cost_args)
        cost, grad, xbuf = run(opt.detach(), True)
        return (cost, _xtraj_view(xbuf, opt)), grad

    def _value_and_grad_host(_nn_params, y, opt_params, rng, _cost_fn, _terminal_cost, extra_dyn_args, extra_cost_args):
        """``value_and_grad`` for arguments in host memory, with as little Python around ``sde4_cost_grad_host`` as the
        checks allow: on a 0.3 ms evaluation the generic route (``_prepare`` / ``run`` / ``_xtraj_view``: a dozen
        tensor operations of 1-3 us each) cost 31 us per call."""
        if extra_dyn_args is not None:
            raise Sde4Error("extra_dyn_args must be None (no shipped model uses them)")
        if not isinstance(_cost_fn, _costs.StageCost):
            raise Sde4Error("_cost_fn must be a sde4mbrl_b200.costs.StageCost (e.g. costs.one_step_cost_f(cost_params, n_u)); "
                            "arbitrary Python cost callables cannot run inside the CUDA kernel")
        if _terminal_cost is not None and not isinstance(_terminal_cost, _costs.StageCost):
            raise Sde4Error("_terminal_cost must be a costs.StageCost or None")
        if extra_cost_args is None:
            raise Sde4Error("extra_cost_args (the reference trajectory xref[H+1, n_x]) is required")
        stage = _cost_fn if constr is None else _costs.with_constraints(_cost_fn, constr, n_u)
        engine.rng_mode = _random.rng_mode()
        o_h = host_f32(opt_params)
        assert o_h.ndim in (2, 3), "The parameters must be a two dimension array"
        assert o_h.shape[-1] == opt_params_size, "Shape of the opt params do not match"
        cost, grad, xbuf = engine.cost_grad_host(_nn_params, y, o_h, time_steps, extra_cost_args, stage.desc,
                                                 terminal=None if _terminal_cost is None else _terminal_cost.desc,
                                                 key=rng, N=num_sample, want_grad=True, want_xtraj=True)
        G = xbuf.shape[2]
        if o_h.ndim == 2:  # xbuf[H+1, n_x+1, N] seen as [N, H+1, n_x+1]
            xtraj = torch.as_strided(xbuf, (num_sample, H + 1, n_x + 1), (1, (n_x + 1) * G, G))
            return (torch.from_numpy(cost.reshape(())), xtraj), torch.from_numpy(grad)
        xtraj = torch.as_strided(xbuf, (G // num_sample, num_sample, H + 1, n_x + 1), (num_sample, 1, (n_x + 1) * G, G))
        return (torch.from_numpy(cost), xtraj), torch.from_numpy(grad)

    _stg = {}

    def _value_and_grad_sharded(red, _nn_params, y, opt_params, rng, _cost_fn, _terminal_cost, extra_dyn_args, extra_cost_args):
        """One problem whose particles span the ranks of ``red`` (a ``dist.CostGradReducer`` over H*n_opt + 1 floats):
        every rank evaluates its ``num_particles`` particles of ``world * num_particles`` (keys
        ``split(rng, N_total)[rank*N + n]``, both means scaled by 1/N_total), K3 writes [cost | grad] straight into the
        reducer's buffer, ONE all-reduce (``jnp.mean`` over all particles, nsde.py:1586).  Host arguments travel as one
        packed pinned copy each way; the returned trajectories are this rank's particles."""
        from . import dist as _dist
        if extra_dyn_args is not None:
            raise Sde4Error("extra_dyn_args must be None (no shipped model uses them)")
        if not isinstance(_cost_fn, _costs.StageCost) or (_terminal_cost is not None and not isinstance(_terminal_cost, _costs.StageCost)):
            raise Sde4Error("_cost_fn / _terminal_cost must be costs.StageCost objects")
        if extra_cost_args is None:
            raise Sde4Error("extra_cost_args (the reference trajectory xref[H+1, n_x]) is required")
        stage = _cost_fn if constr is None else _costs.with_constraints(_cost_fn, constr, n_u)
        term = None if _terminal_cost is None else _terminal_cost.desc
        engine.rng_mode = _random.rng_mode()
        rank, world = _dist.world()
        n_o, n_ref = H * opt_params_size, (H + 1) * n_x
        if red.n != n_o + 1:
            raise Sde4Error("particle_reducer must hold H*n_opt + 1 = %d floats" % (n_o + 1))
        host = is_host(y) and is_host(opt_params) and is_host(rng) and is_host(extra_cost_args)
        dev = red.device
        if host and getattr(red, "_xg", None) is not None:
            # fused transport: the host-buffer entry point itself (sde4_cost_grad_host with sde4_xgpu_reduce) -- one packed
            # pinned H2D copy, K2, K3 with the cross-GPU sum in its last CTA writing the TOTAL [cost | grad] straight into
            # the mapped staging buffer, one stream synchronise
            o_h = opt_params.detach().numpy() if isinstance(opt_params, torch.Tensor) else opt_params
            o_h = host_f32(o_h)
            assert o_h.ndim == 2 and o_h.shape[-1] == opt_params_size, "one problem: opt_params must be [H, n_opt]"
            kw = red.launch_kwargs(1)
            cost, grad, xbuf = engine.cost_grad_host(_nn_params, host_f32(y).reshape(1, n_x), o_h, time_steps,
                                                     host_f32(extra_cost_args), stage.desc, terminal=term, key=rng,
                                                     N=num_sample, want_grad=True, want_xtraj=True,
                                                     particle_offset=rank * num_sample, particle_total=world * num_sample,
                                                     xgpu=kw["xgpu"])
            xtraj = xbuf.view(H + 1, n_x + 1, 1, num_sample).permute(2, 3, 0, 1)[0]
            return (torch.from_numpy(cost)[0], xtraj), torch.from_numpy(grad)
        if host:
            st = _stg.get(dev)
            if st is None:
                nin = n_x + n_o + n_ref + 2
                st = _stg[dev] = (torch.empty(nin, dtype=torch.float32).pin_memory(), torch.empty(nin, dtype=torch.float32, device=dev),
                                  torch.empty(n_o + 1, dtype=torch.float32).pin_memory())
            pin, dbuf, out = st
            hv = pin.numpy()
            hv[:n_x] = host_f32(y).reshape(-1)
            o_h = opt_params.detach().numpy() if isinstance(opt_params, torch.Tensor) else opt_params
            hv[n_x:n_x + n_o] = np.asarray(o_h, np.float32).reshape(-1)
            hv[n_x + n_o:n_x + n_o + n_ref] = host_f32(extra_cost_args).reshape(-1)
            hv[n_x + n_o + n_ref:].view(np.uint32)[:] = host_key(rng).reshape(2)
            dbuf.copy_(pin, non_blocking=True)
            y_t, opt_d = dbuf[:n_x], dbuf[n_x:n_x + n_o].view(H, opt_params_size)
            xref, key = dbuf[n_x + n_o:n_x + n_o + n_ref].view(H + 1, n_x), dbuf[n_x + n_o + n_ref:].view(torch.int32)
        else:
            y_t, opt_d, xref, key = as_f32(y), as_f32(opt_params), as_f32(extra_cost_args), as_key(rng)
        assert opt_d.dim() == 2 and opt_d.shape[-1] == opt_params_size, "one problem: opt_params must be [H, n_opt]"
        kw = red.launch_kwargs(1)
        kw["grad_out"] = kw["grad_out"].view(1, H, opt_params_size)
        _, _, xbuf = engine.cost_grad(_nn_params, y_t.reshape(1, n_x), opt_d.detach().contiguous(), time_steps, xref, stage.desc,
                                      terminal=term, key=key, N=num_sample, want_grad=True, want_xtraj=True,
                                      particle_offset=rank * num_sample, particle_total=world * num_sample, **kw)
        tot = red.reduce()
        xtraj = xbuf.view(H + 1, n_x + 1, 1, num_sample).permute(2, 3, 0, 1)[0]
        if host:
            out.copy_(tot, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            res = out.clone()
            return (res[0], xtraj), res[1:].view(H, opt_params_size)
        return (tot[0], xtraj), tot[1:].view(H, opt_params_size)

    multi_sampling.value_and_grad = value_and_grad
    # set to a dist.CostGradReducer(H*n_opt + 1, device) to shard the particles of ONE problem over the process group
    # (value_and_grad only); None: single-GPU evaluation
    multi_sampling.particle_reducer = None
    # what a device-side solver (apg_batched.BatchedAPG) needs to evaluate the same cost without this closure
    multi_sampling.final_stage = lambda _cost_fn: _cost_fn if constr is None else _costs.with_constraints(_cost_fn, constr, n_u)
    multi_sampling.num_particles = num_sample
    multi_sampling.n_opt = opt_params_size
    multi_sampling.engine = engine
    multi_sampling.time_steps = time_steps

    def device_solver(_nn_params, stage, term, optimizer_params, box=None):
        """The device-side APG solver (``apg_batched.BatchedAPG``) of this cost for a final stage cost / terminal cost /
        ``apg_mpc`` options / box, cached on everything that shapes it; the parameter VALUES are refreshed on every call
        (a tree whose leaves were updated in place, or a new tree of the same architecture, reuses the solver and its
        CUDA graph)."""
        from .apg_batched import BatchedAPG
        import json
        bkey = None if box is None else (box.first, box.lb_np.tobytes(), box.ub_np.tobytes())
        key = (bytes(stage.desc), None if term is None else bytes(term.desc),
               json.dumps(optimizer_params, sort_keys=True, default=float), bkey)
        cache = multi_sampling.__dict__.setdefault("_device_solvers", {})
        solver = cache.get(key)
        if solver is not None:
            return solver.set_params(_nn_params)
        lb = ub = None
        if box is not None:
            lb, ub = np.full(opt_params_size, -np.inf, np.float32), np.full(opt_params_size, np.inf, np.float32)
            lb[box.first:], ub[box.first:] = box.lb_np, box.ub_np
        if len(cache) > 8:
            cache.clear()
        solver = cache[key] = BatchedAPG(engine, _nn_params, stage, optimizer_params, time_steps, num_sample, lb=lb, ub=ub,
                                         terminal=term)
        return solver

    multi_sampling.device_solver = device_solver

    class _BoundCost:
        """``lambda opt: multi_sampling(nn_params, y, opt, rng, cost_fn, ...)[0]`` -- the closure ``apg.py`` is handed as
        ``cost_fn`` (sde_mpc_design.py:207-219) -- as an object that still knows its arguments.  Called, it is that
        closure; ``apg.init_apg`` / ``apg.one_step_apg`` take value and gradient from its ONE fused launch
        (``value_and_grad``) instead of an autograd graph; and ``apg.apg`` hands a fresh line-search solve (with no or a
        box ``proximal_fn``) to the device-side solver, one CUDA-graph replay per iteration instead of 5-8 host-driven
        cost calls -- same iterates, same ``APGState``."""

        def __init__(self, _nn_params, y, rng, _cost_fn, _terminal_cost=None, extra_cost_args=None):
            self.nn, self.y, self.rng, self.cost, self.term, self.xref = _nn_params, y, rng, _cost_fn, _terminal_cost, extra_cost_args

        def __call__(self, opt):
            return multi_sampling(self.nn, self.y, opt, self.rng, self.cost, self.term, extra_cost_args=self.xref)[0]

        def value_and_grad(self, opt):
            (c, _), g = value_and_grad(self.nn, self.y, opt, self.rng, self.cost, self.term, extra_cost_args=self.xref)
            return c, g

        def solve_on_device(self, st, optimizer_params, proximal_fn, check_every=4):
            """``apg.apg`` for a state as ``init_apg`` returns it; None when the device solver does not cover the call
            (no line search, a proximal operator that is not a box, a state in the middle of a solve, several problems)."""
            box = proximal_fn if isinstance(proximal_fn, _Box) else getattr(proximal_fn, "box", None)
            if "linesearch" not in optimizer_params or (proximal_fn is not None and not isinstance(box, _Box)):
                return None
            if multi_sampling.particle_reducer is not None or not isinstance(st.xk, torch.Tensor) or st.xk.dim() != 2:
                return None
            fresh = (st.num_steps == 1 and st.num_iter_no_improv == 0 and st.not_done and st.opt_cost == st.curr_cost
                     and st.curr_cost == st.init_cost and st.avg_momentum == st.momentum and st.avg_stepsize == st.stepsize
                     and (st.xk is st.yk or torch.equal(st.xk, st.yk)) and (st.xk is st.x_opt or torch.equal(st.xk, st.x_opt)))
            if not fresh:
                return None
            solver = device_solver(self.nn, multi_sampling.final_stage(self.cost), self.term, optimizer_params, box)
            solver.init(st.xk, self.y, self.xref, self.rng, momentum=np.float32(st.momentum), stepsize=np.float32(st.stepsize))
            solver.run(check_every=check_every)
            out = solver.state_of(0)
            vec = {k: getattr(out, k).clone() for k in ("x_opt", "xk", "yk", "grad_yk")}
            if not st.xk.is_cuda:
                vec = {k: v.cpu() for k, v in vec.items()}
            return out._replace(**vec)

    multi_sampling.bind = _BoundCost

    def _construct_opt_params(u=None, x=None):
        """``nsde.py:1532-1573`` (including the doubled ``+1e-4`` of the 1-D ``u`` path)."""
        num_var = H
        dev = engine_device()
        zero_u = torch.ones((num_var, n_u), device=dev) * 1e-4
        zero_x = torch.ones((num_var, n_slack), device=dev) * 1e-4 if has_slack else None
        ssv = None
        if has_slack:
            ss = np.asarray(slack_scaling, np.float32)
            ssv = torch.ones((num_var, n_slack), device=dev) * (as_f32(ss) if ss.ndim >= 1 else float(ss))
        if u is not None:
            u = as_f32(u)
            if u.dim() == 1:
                u = u[None].repeat(num_var, 1) + 1e-4
        if x is not None:
            x = as_f32(x)
            if x.dim() == 1:
                x = x[None].repeat(num_var + 1, 1)
            x = x.clone()
            x[0, :] = x[-1, :]
        sidx = [int(i) for i in state_idx] if has_slack else None
        if u is None and not has_slack:
            return zero_u
        if u is None and x is None:
            return torch.cat((zero_u, zero_x), dim=1)
        if u is None and x is not None:
            assert x.dim() == 2 and x.shape[0] == num_var + 1
            return torch.cat((zero_u, x[:-1, sidx] / ssv), dim=1) + 1e-4
        if u is not None and not has_slack:
            assert u.dim() == 2 and u.shape[0] == num_var
            return u + 1e-4
        if u is not None and x is None:
            assert u.dim() == 2 and u.shape[0] == num_var
            return torch.cat((u, zero_x), dim=1) + 1e-4
        assert u.dim() == 2 and x.dim() == 2 and u.shape[0] + 1 == x.shape[0]
        return torch.cat((u, x[:-1, sidx] / ssv), dim=1) + 1e-4

    vmapped_prox = proximal_fn  # the box clip broadcasts over the time axis
    if vmapped_prox is not None:
        construct_opt_params = lambda u=None, x=None: vmapped_prox(_construct_opt_params(u, x))
    else:
        construct_opt_params = _construct_opt_params
    return nn_params, multi_sampling, vmapped_prox, constr_cost, construct_opt_params


def engine_device():
    from .engine import _dev
    return _dev()


def create_model_loss_fn(model_params, loss_params, sde_constr=ControlledSDE, verbose=True, **extra_args_sde_constr):
    """``sde4mbrl/nsde.py:1032-1338``: returns ``(nn_params, loss_fn, nonneg_projection, test_fn)``.

    ``loss_fn(nn_params, y, u, rng) -> (total_loss, dict)`` evaluates the data term on the GPU and the
    weight penalty on the host.  JAX users differentiate ``loss_fn`` with ``jax.grad``; here the
    parameter gradient comes out of the same fused launch: ``loss_fn.value_and_grad(nn_params, y, u, rng)
    -> (total_loss, dict, grads)`` with ``grads`` a tree shaped like ``nn_params`` (model scalars and
    un-differentiated leaves get zeros).  Like the reference this MUTATES ``model_params`` /
    ``loss_params`` (nsde.py:1064-1134).

    The distance-aware regulariser (``density_loss`` gradient-norm / strong-convexity / mu terms) is evaluated by
    ``density_loss.py`` at the data points with the reference's key chain; ``num_sample2consider`` below the horizon
    becomes a per-sample 0/1 step weight drawn on the device (``sde4_rng_choice_mask``);
    ``u_sampling_strategy='random'`` draws each sample's per-step data index on the device (``sde4_rng_randint_many``,
    nsde.py:65) and, with more than one particle, evaluates the particles of a minibatch element in one launch per
    particle index (each particle has its own controls).  Not covered (raises): ``learn_mucoeff.quad_approx``
    (undefined in the reference itself)."""
    _check_constr(sde_constr)
    vprint = print if verbose else (lambda *a, **k: None)
    params_model = model_params
    num_sample = loss_params.get("num_particles", params_model.get("num_particles", 1))
    params_model["num_particles"] = num_sample
    step_size = loss_params.get("stepsize", params_model["stepsize"])
    params_model["stepsize"] = step_size
    data_stepsize = loss_params.get("data_stepsize", step_size)
    if abs(step_size - data_stepsize) <= 1e-6:
        num_steps2data = 1
    else:
        num_steps2data = int((step_size / data_stepsize) + 0.5)  # nsde.py:1078
    horizon = loss_params.get("horizon", params_model.get("horizon", 1))
    data_horizon = horizon * num_steps2data
    params_model["horizon"] = horizon
    params_model["num_steps2data"] = num_steps2data
    loss_params["data_horizon"] = data_horizon
    vprint("Using [ N = {} ] particles for the loss".format(num_sample))
    vprint("Using [ T = {} ] horizon for the loss".format(horizon))
    vprint("Using [ dt = {} ] stepsize for the loss".format(step_size))
    vprint("Using [ num_steps2data = {} ] steps between data points".format(num_steps2data))
    for k in ("num_sample2consider", "obs_weights", "density_on_partial", "u_sampling_strategy"):
        if k in loss_params:
            params_model[k] = loss_params[k]
    for k in ("num_short_dt", "short_step_dt", "long_step_dt"):
        params_model.pop(k, None)
    if "diffusion_density_nn" not in params_model or "density_nn" not in params_model["diffusion_density_nn"]:
        loss_params.pop("density_loss", None)
    if "density_loss" in loss_params:  # nsde.py:1118-1132
        _lm = loss_params["density_loss"].get("learn_mucoeff", False)
        if type(_lm) is bool:
            loss_params["density_loss"]["learn_mucoeff"] = {"quad_approx": False, "type": "constant" if not _lm else "global"}
        loss_params["density_loss"]["learn_mucoeff"].setdefault("quad_approx", False)
        loss_params["density_loss"]["learn_mucoeff"].setdefault("type", "constant")
        params_model["density_loss"] = loss_params["density_loss"]
        if loss_params["density_loss"]["learn_mucoeff"]["type"] == "constant":
            loss_params["pen_mu_coeff"] = 0.0
    else:
        loss_params["pen_mu_coeff"] = 0.0
    dl = loss_params.get("density_loss", None)
    n_consider = int(params_model.get("num_sample2consider", horizon))
    if dl is not None and dl["learn_mucoeff"].get("quad_approx", False):
        # the reference dispatches to self.density_loss_theory, which it never defines (nsde.py:778): AttributeError there
        raise Sde4Error("density_loss.learn_mucoeff.quad_approx: the reference has no density_loss_theory to mirror")
    strategy = params_model.get("u_sampling_strategy", "first")
    if strategy not in ("first", "mean", "median", "random"):  # nsde.py:69
        raise ValueError("Unknown u_sampling_strategy: {}. Choose from first, mean, median, random".format(strategy))

    model = sde_constr(params_model, **extra_args_sde_constr)
    nn_params = model.init_params(loss_params.get("seed", 0))
    if dl is not None:  # parameters created by mu_coeff_nn (nsde.py:657-680)
        for mod, leaves in _density.init_mu_params(model, dl, np.random.default_rng(loss_params.get("seed", 0) + 17)).items():
            nn_params.setdefault(mod, {}).update(leaves)
    engine = Engine(model, _random.rng_mode())
    time_steps = compute_timesteps(params_model)
    n_feat = 5 if model.module_name == "cart_pole_sde" else model.n_x  # state_transform_for_loss
    ow = np.broadcast_to(np.asarray(params_model.get("obs_weights", 1.0), np.float64), (n_feat,))
    data_cost = _costs.QuadraticCost(list(1.0 / ow ** 2), [0.0] * model.n_u, [0.0] * model.n_u,
                                     feat="cartpole" if model.module_name == "cart_pole_sde" else "identity")
    pen_w = loss_params.get("pen_weights", 0)
    default_w = loss_params.get("default_weights", 1.0)
    special = loss_params.get("special_parameters_pen", {})
    default_val = loss_params.get("default_parameters_val", 0.0)

    nominal_val = loss_params.get("nominal_parameters_val", {})
    special_val = loss_params.get("special_parameters_val", {})

    def _matched(tree, keyed, default, current=None):
        """``get_penalty_parameters`` (sde4mbrl/utils.py:119-146) on a two-level haiku tree, as {module: {leaf: scalar}}:
        a key that is a substring of a MODULE name sets every leaf of that module (later keys of ``keyed`` win) and the
        module is not looked into any further; in the other modules a key that is a substring of a LEAF name sets that
        leaf; everything else gets ``default`` (or keeps ``current`` when ``default`` is None)."""
        out = {}
        for mod, leaves in tree.items():
            hit = [v for k, v in keyed.items() if k in mod]
            if hit:
                out[mod] = {name: hit[-1] for name in leaves}
                continue
            out[mod] = {}
            for name in leaves:
                lh = [v for k, v in keyed.items() if k in name]
                out[mod][name] = lh[-1] if lh else (default if default is not None else current[mod][name])
        return out

    _pen_cache = {}

    def _penalty(_nn_params, want_grad):
        """pen_weights * sum_leaves coeff * ||p - nominal||^2 (nsde.py:1155-1168, 1216-1219): coefficients from
        ``special_parameters_pen`` / ``default_weights``; nominal values = ``default_parameters_val`` everywhere, then
        ``nominal_parameters_val`` ({module: {leaf: value}}, update_same_struct_dict), then ``special_parameters_val``
        matched like the coefficients."""
        sig = tuple((mod, tuple(leaves)) for mod, leaves in _nn_params.items())
        if _pen_cache.get("sig") != sig:  # the two trees depend on the NAMES only
            coeffs = _matched(_nn_params, special, default_w)
            nom = {mod: {name: default_val for name in leaves} for mod, leaves in _nn_params.items()}
            for mod, sub in nominal_val.items():
                if mod in nom and isinstance(sub, dict):
                    for name, v in sub.items():
                        nom[mod][name] = v
            nom = _matched(_nn_params, special_val, None, nom)
            _pen_cache.update(sig=sig)
            # flat float64 views in tree order: one vector operation per call instead of a handful per leaf
            index, cf, nf, o = [], [], [], 0
            for mod, leaves in _nn_params.items():
                for name, arr in leaves.items():
                    shp = np.shape(arr)
                    n = int(np.prod(shp)) if len(shp) else 1
                    index.append((mod, name, shp, o, n))
                    cf.append(np.full(n, coeffs[mod][name], np.float64))
                    nf.append(np.broadcast_to(np.asarray(nom[mod][name], np.float64), shp).reshape(-1))
                    o += n
            _pen_cache.update(index=index, cf=np.concatenate(cf) if cf else np.zeros(0), nf=np.concatenate(nf) if nf else np.zeros(0))
        index, cf, nf = _pen_cache["index"], _pen_cache["cf"], _pen_cache["nf"]
        flat = np.empty(cf.shape[0], np.float64)
        for mod, name, shp, o, n in index:
            flat[o:o + n] = np.asarray(_nn_params[mod][name]).reshape(-1)
        flat -= nf
        tot = float(np.dot(cf, flat * flat))
        grads = {}
        if want_grad:
            gflat = (2.0 * pen_w * cf * flat).astype(np.float32)
            for mod, name, shp, o, n in index:
                grads.setdefault(mod, {})[name] = gflat[o:o + n].reshape(shp)
        return tot, grads

    # The distance-aware regulariser runs on the library's own kernels (density_loss.NativeRegulariser ->
    # sde4_density_reg) whenever they cover the configuration.  loss_params["regulariser"] = "graph" / "eager" (or the
    # older "graph_regulariser": False) select the torch restatements instead: one CUDA-graph replay, or plain autograd.
    reg_mode = loss_params.get("regulariser", "native" if loss_params.get("graph_regulariser", True) else "eager")
    if reg_mode not in ("native", "graph", "eager"):
        raise Sde4Error("loss_params['regulariser'] must be 'native', 'graph' or 'eager'")
    native = None
    if dl is not None and reg_mode == "native":
        native = _density.NativeRegulariser(model, engine, loss_params, dl, num_sample)
        if not native.supported():
            native, reg_mode = None, "graph"
    graphed = _density.GraphedRegulariser(model, loss_params, dl, num_sample) if (dl is not None and reg_mode == "graph") else None

    def _prepare(y, u):
        y, u = as_f32(y), as_f32(u)
        assert y.shape[1] == u.shape[1] + 1, "Trajectory horizon must match"
        assert horizon * num_steps2data == u.shape[1], "Trajectory horizon and num_steps2data must match the number of control inputs"
        B = y.shape[0]
        y_values = y[:, ::num_steps2data]
        uu = u.reshape(B, horizon, num_steps2data, model.n_u)
        if strategy == "first":
            u_values = uu[:, :, 0]
        elif strategy == "mean":
            u_values = uu.mean(dim=2)
        elif strategy == "median":
            u_values = uu.median(dim=2).values
        else:  # 'random': the caller gathers with its own indexes (per sample in loss_fn, per call in test_fn)
            u_values = uu
        return y_values.contiguous(), u_values.contiguous()

    def _pick(uu, idx):
        """uu[B, H, num_steps2data, n_u], idx[..., H] (int32, broadcast against B) -> uu[b, t, idx[.., t]] (nsde.py:66)."""
        ii = idx.long()[..., None, None].expand(*idx.shape, 1, uu.shape[-1])
        src = uu if idx.dim() == 2 else uu[:, None].expand(uu.shape[0], idx.shape[1], *uu.shape[1:])
        return torch.gather(src, idx.dim(), ii).squeeze(idx.dim()).contiguous()

    _stage = {}  # pinned read-back buffer [loss_b | packed gradient]

    def value_and_grad(_nn_params, y, u, rng, extra_args=None, want_grad=True):
        if extra_args is not None:
            raise Sde4Error("extra_args must be None")
        key = as_key(rng)
        assert key.dim() == 1, "THe rng key is splitted inside the loss function computation"
        y_values, u_values = _prepare(y, u)
        B = y_values.shape[0]
        keys = _random.split(key, B)  # nsde.py:1196
        engine.rng_mode = _random.rng_mode()
        step_weight = None
        k_s2c = None
        if n_consider < horizon or strategy == "random":
            k_bn = _random.split_many(keys, num_sample)  # [B, N, 2]  (nsde.py:1187)
            k_s2c = _random.split_many(k_bn, 3)[:, :, 1].reshape(-1, 2)  # rng_sample2consider (nsde.py:709)
        if n_consider < horizon:  # nsde.py:753-760: every (b, n) keeps its own random subset of the time indexes
            step_weight = _random.choice_mask_many(k_s2c, horizon, n_consider)
        on_partial = step_weight is not None and params_model.get("density_on_partial", False)
        pend = None
        pk_all = engine.packed(_nn_params)  # ONE fingerprint + upload per call: the data term and the regulariser share it
        if dl is not None and native is not None and not on_partial:
            # distance-aware regulariser at the data points on the library's kernels, on a side stream: the data term's
            # launches (a few thousand threads) run beside it; both are read back behind ONE synchronisation
            if "side" not in _stage:
                _stage["side"] = torch.cuda.Stream(device=y_values.device)
            side, cur = _stage["side"], torch.cuda.current_stream(y_values.device)
            pk = pk_all  # (uploaded on the main stream, which the side stream then waits for)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                pend = native.launch(_nn_params, as_f32(y), as_f32(u), key, pk=pk)
        if strategy == "random":  # nsde.py:735: every (b, n) draws its own data sub-step per solver step
            ridx = _random.randint_many(k_s2c, horizon, 0, num_steps2data).reshape(B, num_sample, horizon)
            u_bn = _pick(u_values, ridx)  # [B, N, H, n_u]
            loss_b, gpacked = None, None
            for n in range(num_sample):  # one launch per particle index: particle n of every element, its own controls
                sw = None if step_weight is None else step_weight[:, n::num_sample].contiguous()
                lb, gp = engine.loss_grad(pk_all, y_values, u_bn[:, n].contiguous(), time_steps, data_cost.desc, keys, N=1,
                                          step_weight=sw, particle_offset=n, particle_total=num_sample)
                loss_b = lb.clone() if loss_b is None else loss_b + lb
                gpacked = gp.clone() if gpacked is None else gpacked + gp
            u_values = u_bn[:, 0]
        else:
            loss_b, gpacked = engine.loss_grad(pk_all, y_values, u_values, time_steps, data_cost.desc, keys, N=num_sample,
                                               step_weight=step_weight)
        nb, npar = loss_b.numel(), gpacked.numel()
        if _stage.get("n") != (nb, npar):
            _stage["n"] = (nb, npar)
            _stage["host"] = torch.zeros(nb + npar, dtype=torch.float32).pin_memory()
        main = torch.cuda.current_stream(loss_b.device)
        _stage["host"][:nb].copy_(loss_b, non_blocking=True)
        _stage["host"][nb:].copy_(gpacked, non_blocking=True)
        # host work that needs nothing from the device goes in front of the synchronisation
        w_loss, wgrads = _penalty(_nn_params, want_grad) if pen_w > 0 else (0.0, {})
        if pend is not None:
            main.wait_stream(_stage["side"])
        main.synchronize()
        hres = _stage["host"].numpy()
        loss_data = float(hres[:nb].mean(dtype=np.float64))
        total = 0.0
        if loss_params.get("pen_data", 0) > 0:
            total += loss_data * loss_params["pen_data"]
        dinfo, dgrads = {}, {}
        if pend is not None or (dl is not None and graphed is not None and not on_partial):
            # (graphed: one CUDA-graph replay of the torch restatement, density_loss.GraphedRegulariser)
            dval, dinfo, dgrads_np = native.finish(pend, mu_only=True) if pend is not None else graphed(_nn_params, as_f32(y), as_f32(u), key)
            total += dval
            if want_grad:
                dgrads = {mod: {k: torch.from_numpy(v) for k, v in leaves.items()} for mod, leaves in dgrads_np.items()}
        elif dl is not None:
            # distance-aware regulariser at the data points (density_loss.py): library GEMMs + autograd
            tt = _density.to_torch_tree(_nn_params, y_values.device)
            partial = None
            if on_partial:  # nsde.py:794-797
                if strategy != "first":
                    raise Sde4Error("density_on_partial needs u_sampling_strategy 'first'")
                partial = (y_values, u_values, step_weight)
            gn, sc, mu, _dv = _density.density_terms(model, tt, as_f32(y), as_f32(u), key, num_sample, dl, partial)
            dtot, dinfo = _density.weighted_terms(loss_params, dl, gn, sc, mu)
            if want_grad and dtot.requires_grad:
                dtot.backward()
                dgrads = {mod: {k: v.grad for k, v in leaves.items() if v.grad is not None} for mod, leaves in tt.items()}
            total += float(dtot.detach())
            dinfo = {k: float(v.detach()) if isinstance(v, torch.Tensor) else float(v) for k, v in dinfo.items()}
        if pen_w > 0:
            total += pen_w * w_loss
        info = {"totalLoss": total, "gradDensity": 0.0, "lossMuCoeff": 0.0, "densitySCVEX": 0.0, "weights": w_loss,
                "muCoeff": 0.0, "dataLoss": loss_data}
        info.update(dinfo)
        if not want_grad:
            return total, info, None
        pd = float(loss_params.get("pen_data", 0)) if loss_params.get("pen_data", 0) > 0 else 0.0
        gblock = pd * hres[nb:]
        if pend is not None:  # the regulariser's block has the same packed layout: add before the (single) unpacking
            gblock = gblock + native.packed_grad(pend)
        gtree = model.unpack_grads(gblock, engine.desc)
        grads = {}
        for mod, leaves in _nn_params.items():
            for name, arr in leaves.items():
                g = wgrads[mod][name] if (mod in wgrads and name in wgrads[mod]) else np.zeros(np.shape(arr), np.float32)
                if mod in gtree and name in gtree[mod]:
                    g = g + gtree[mod][name]
                if mod in dgrads and name in dgrads[mod]:
                    g = g + dgrads[mod][name].cpu().numpy()
                grads.setdefault(mod, {})[name] = g
        return total, info, grads

    def loss_fn(_nn_params, y, u, rng, extra_args=None):
        total, info, _ = value_and_grad(_nn_params, y, u, rng, extra_args, want_grad=False)
        return total, info

    loss_fn.value_and_grad = value_and_grad
    loss_fn.engine = engine
    loss_fn.weight_penalty = _penalty  # (unweighted sum, pen_weights-scaled gradients)
    loss_fn.regulariser = "native" if native is not None else (reg_mode if dl is not None else None)

    nonneg = set(params_model.get("noneg_params", []))
    nonzero = loss_params.get("nonneg_nonzero", {})

    def nonneg_projection(_params):
        """nsde.py:1181-1182: clamp the named parameters from below."""
        out = {}
        for mod, leaves in _params.items():
            out[mod] = {}
            for name, arr in leaves.items():
                lo = None
                if any(k in mod or k in name for k in nonneg):
                    lo = 0.0
                    for k, v in nonzero.items():
                        if k in mod or k in name:
                            lo = v
                out[mod][name] = np.maximum(arr, lo) if lo is not None else arr
        return out

    _, multi_sampling_sde = create_sampling_fn(params_model, sde_constr, loss_params.get("seed", 0),
                                               loss_params["num_particles_test"], **extra_args_sde_constr)

    def test_fn(_nn_params, y, u, rng, extra_args=None):
        """nsde.py:1268-1335: error of the particle-mean prediction and mean particle spread."""
        key = as_key(rng)
        y_values, u_values = _prepare(y, u)
        B = y_values.shape[0]
        if strategy == "random":  # nsde.py:1306-1310: one index per solver step, shared by the minibatch
            key, rng_u = _random.split(key, 2)
            u_values = _pick(u_values, _random.randint(rng_u, (horizon,), 0, num_steps2data).reshape(1, horizon).expand(B, horizon))
        keys = _random.split(key, B)
        _, yevol, _ = multi_sampling_sde(_nn_params, y_values[:, 0], u_values, keys)  # [B, N, H+1, n_y]

        def tf_loss(z):
            if model.module_name == "cart_pole_sde":
                return torch.stack([z[..., 0], z[..., 1], torch.sin(z[..., 2]), torch.cos(z[..., 2]), z[..., 3]], dim=-1)
            return z
        w = as_f32(np.broadcast_to(np.asarray(loss_params.get("obs_weights", 1.0), np.float32), (n_feat,)))
        err = (((tf_loss(y_values) - tf_loss(yevol.mean(dim=1))) / w) ** 2).sum(dim=(-1, -2)).mean()
        std_val = yevol.std(dim=1, unbiased=False).sum(dim=(-1, -2)).mean()
        return float(err), {"TestErrorData": float(err), "TestStdData": float(std_val)}

    return nn_params, loss_fn, nonneg_projection, test_fn

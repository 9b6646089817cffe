"""Accelerated proximal gradient (APG) solver -- mirror of ``sde4mbrl/apg.py``.

Same entry points and ``APGState`` as the reference.  The control logic (Armijo backtracking,
non-monotone momentum, stopping rules) stays on the host exactly as the north star asks; every
``cost_fn`` evaluation and its gradient hit the fused CUDA kernels through
``nsde.create_online_cost_sampling_fn(...).multi_sampling`` (differentiable via torch autograd,
which plays the role of ``jax.value_and_grad``, apg.py:82,224).  A cost function made with
``multi_sampling.bind(nn_params, y, rng, cost_fn, ...)`` instead of a lambda keeps its arguments visible: value and
gradient then come from one fused launch, and ``apg`` runs a fresh line-search solve on the device-side solver
(``apg_batched.BatchedAPG``) -- identical iterates, a third of the time (cartpole C2: 16.7 -> 5.7 ms).

The reference bounds its loops with ``lax.scan`` + ``lax.cond`` no-ops (``_while_loop``,
apg.py:451-464); running "while cond and it < max_iter" returns the identical state.
"""
from typing import NamedTuple

import numpy as np
import torch

opt_stop_criteria = lambda grad_val: torch.sum(torch.square(grad_val))  # apg.py:11


class APGState(NamedTuple):
    """apg.py:277-304 (field order kept)."""
    num_iter_no_improv: int
    num_steps: int
    momentum: float
    avg_momentum: float
    avg_linesearch: float
    avg_stepsize: float
    stepsize: float
    opt_cost: float
    x_opt: torch.Tensor
    curr_cost: float
    init_cost: float
    grad_sqr: float
    not_done: bool
    xk: torch.Tensor
    yk: torch.Tensor
    grad_yk: torch.Tensor


def _f(v):
    """host float32 scalar (the reference keeps every scalar as a float32 jnp array)."""
    if isinstance(v, torch.Tensor):
        return np.float32(v.item())
    return np.float32(v)


def _value_and_grad(cost_fn, x):
    fused = getattr(cost_fn, "value_and_grad", None)  # nsde multi_sampling.bind(...): one fused launch, no autograd graph
    if fused is not None:
        c, g = fused(x.detach())
        return _f(c), g.detach()
    xx = x.detach().clone().requires_grad_(True)
    c = cost_fn(xx)
    (g,) = torch.autograd.grad(c, xx)
    return _f(c), g.detach()


def _value(cost_fn, x):
    with torch.no_grad():
        return _f(cost_fn(x.detach()))


def init_opt_state(opt_var, optimizer_params):
    """apg.py:14-49."""
    init_step_size = _f(optimizer_params["linesearch"]["init_stepsize"] if "linesearch" in optimizer_params
                        else optimizer_params["stepsize"])
    momentum = _f(optimizer_params.get("beta_init", 0.25))
    return APGState(num_steps=1, num_iter_no_improv=0, momentum=momentum, avg_momentum=momentum,
                    avg_linesearch=_f(1.0), stepsize=init_step_size, avg_stepsize=init_step_size,
                    opt_cost=_f(0.0), x_opt=opt_var, curr_cost=_f(0.0), init_cost=_f(0.0), grad_sqr=_f(0.0),
                    not_done=True, xk=opt_var, yk=opt_var, grad_yk=torch.zeros_like(opt_var))


def init_apg(x0, cost_fn, optimizer_params, proximal_fn=None, momentum=None, stepsize=None):
    """apg.py:53-113: evaluates the cost and its gradient at ``x0``."""
    init_step_size = _f(optimizer_params["linesearch"]["init_stepsize"] if "linesearch" in optimizer_params
                        else optimizer_params["stepsize"])
    if proximal_fn is not None:
        x0 = proximal_fn(x0, stepsize=init_step_size)
    cost_x0, grad_x0 = _value_and_grad(cost_fn, x0)
    grad_x0_nsqr = _f(opt_stop_criteria(grad_x0))
    if momentum is None:
        momentum = _f(optimizer_params.get("beta_init", 0.25))
    if stepsize is None:
        stepsize = init_step_size
    return APGState(num_steps=1, num_iter_no_improv=0, momentum=_f(momentum), avg_momentum=_f(momentum),
                    avg_linesearch=_f(1.0), stepsize=_f(stepsize), avg_stepsize=_f(stepsize), opt_cost=cost_x0,
                    x_opt=x0, curr_cost=cost_x0, init_cost=cost_x0, grad_sqr=grad_x0_nsqr, not_done=True,
                    xk=x0, yk=x0, grad_yk=grad_x0)


def wolfe_cond_violated(stepsize, coef, f_cur, f_next, grad_sqnorm):
    """apg.py:324-329."""
    eps = np.finfo(np.float32).eps
    return bool(np.float32(stepsize * coef * grad_sqnorm) > np.float32(f_cur - f_next + eps))


def update_search(_stepsize, xcurr, grad, cost_fn, decrease_factor, max_stepsize):
    """apg.py:312-321."""
    _stepsize = np.float32(min(_stepsize * decrease_factor, max_stepsize))
    _xnext = xcurr - float(_stepsize) * grad
    return _stepsize, _xnext, _value(cost_fn, _xnext)


def armijo_line_search(cost_fn, xcurr, f_cur, grad, grad_sqnorm, stepsize, user_params):
    """apg.py:332-379.  Returns ``(stepsize, xnext, f_next, num_iter)``."""
    coef, decrease_factor, max_stepsize = (user_params[v] for v in ("coef", "decrease_factor", "max_stepsize"))
    xnext = xcurr - float(stepsize) * grad
    f_next = _value(cost_fn, xnext)
    num_iter, violated, it = 1, True, 0
    while violated and it < user_params["maxls"]:
        violated = wolfe_cond_violated(stepsize, coef, f_cur, f_next, grad_sqnorm)
        if violated:
            stepsize, xnext, f_next = update_search(stepsize, xcurr, grad, cost_fn, decrease_factor, max_stepsize)
        num_iter += int(violated)
        it += 1
    return stepsize, xnext, f_next, num_iter


def one_step_apg(state_ops, cost_fn, optimizer_params, proximal_fn):
    """apg.py:164-274."""
    st = state_ops
    xpast, ycurr, cost_ycurr, grad_ycurr = st.xk, st.yk, st.curr_cost, st.grad_yk
    if "linesearch" in optimizer_params:
        _stepsize = st.stepsize
        if optimizer_params["linesearch"]["reset_option"] == "increase":
            _stepsize = np.float32(min(_stepsize * optimizer_params["linesearch"]["increase_factor"],
                                       optimizer_params["linesearch"]["max_stepsize"]))
        stepsize, xcurr, cost_xcurr, linesearch_numstep = armijo_line_search(
            cost_fn, ycurr, cost_ycurr, grad_ycurr, st.grad_sqr, _stepsize, optimizer_params["linesearch"])
    else:
        stepsize = st.stepsize
        xcurr = ycurr - float(stepsize) * grad_ycurr
        if proximal_fn is None:
            cost_xcurr = _value(cost_fn, xcurr)
        linesearch_numstep = 1
    if proximal_fn is not None:
        xcurr = proximal_fn(xcurr, stepsize=stepsize)
        cost_xcurr = _value(cost_fn, xcurr)
    vcurr = xpast + float(st.momentum) * (xcurr - xpast)
    cost_vcurr = _value(cost_fn, vcurr)
    xcurr_less_vcurr = bool(cost_xcurr <= cost_vcurr)
    ynext = xcurr if xcurr_less_vcurr else vcurr
    cost_ynext = cost_xcurr if xcurr_less_vcurr else cost_vcurr
    _, grad_ynext = _value_and_grad(cost_fn, ynext)
    grad_ynext_nsqr = _f(opt_stop_criteria(grad_ynext))
    if optimizer_params.get("moment_scale", None) is None:
        moment_next = np.float32(st.num_steps / (st.num_steps + 3.0))
    else:
        ms = optimizer_params["moment_scale"]
        moment_next = np.float32(ms * st.momentum if xcurr_less_vcurr else min(st.momentum / ms, 1.0))
    stop_not_sat = bool(abs(cost_ynext - st.opt_cost) > optimizer_params["atol"] + optimizer_params["rtol"] * abs(st.opt_cost))
    cost_improv = bool(st.opt_cost > cost_ynext)
    new_opt_cost = cost_ynext if cost_improv else st.opt_cost
    new_x_opt = ynext if cost_improv else st.x_opt
    new_noimprov = 0 if cost_improv else st.num_iter_no_improv + 1
    stop_not_sat = stop_not_sat and (new_noimprov < optimizer_params["max_no_improvement_iter"])
    num_total_step = st.num_steps + 1
    return APGState(num_steps=num_total_step, num_iter_no_improv=new_noimprov, momentum=moment_next,
                    avg_momentum=np.float32((st.num_steps * st.avg_momentum + moment_next) / num_total_step),
                    avg_linesearch=np.float32((st.num_steps * st.avg_linesearch + linesearch_numstep) / num_total_step),
                    stepsize=np.float32(stepsize),
                    avg_stepsize=np.float32((st.num_steps * st.avg_stepsize + stepsize) / num_total_step),
                    opt_cost=new_opt_cost, curr_cost=cost_ynext, x_opt=new_x_opt, init_cost=st.init_cost,
                    grad_sqr=grad_ynext_nsqr, not_done=stop_not_sat, xk=xcurr, yk=ynext, grad_yk=grad_ynext)


def _while_loop(cond_fun, body_fun, init_val, max_iter):
    """apg.py:451-464 -- bounded while loop."""
    val, it = init_val, 0
    while cond_fun(val) and it < max_iter:
        val = body_fun(val)
        it += 1
    return val


def apg(optim_state, cost_fn, optimizer_params, proximal_fn=None):
    """apg.py:115-161."""
    assert isinstance(optim_state, APGState), "The input optim_state should be an instance of APG"
    on_device = getattr(cost_fn, "solve_on_device", None)  # nsde multi_sampling.bind(...): the whole solve on the GPU
    if on_device is not None:
        out = on_device(optim_state, optimizer_params, proximal_fn)
        if out is not None:
            return out
    return _while_loop(cond_fun=lambda t: t.not_done,
                       body_fun=lambda s: one_step_apg(s, cost_fn, optimizer_params, proximal_fn),
                       init_val=optim_state, max_iter=optimizer_params["max_iter"])

"""CPU restatement of ``sde4mbrl/apg.py`` -- test infrastructure only.

Accelerated proximal gradient with Armijo backtracking, restated with plain
Python control flow.  The reference expresses its loops as bounded
``lax.scan``s whose bodies become no-ops once the condition is false
(``_while_loop``, apg.py:451-464); executing "while cond and it < max_iter"
is value-identical.  PARITY UNPINNED against the reference itself (see
``oracle/__init__.py``); every step cites apg.py.

``cost_fn`` maps a torch tensor to a torch scalar and must be differentiable by
torch autograd (the reference uses ``jax.value_and_grad`` / ``jax.grad``,
apg.py:82,224).
"""
from typing import NamedTuple

import numpy as np
import torch


class APGState(NamedTuple):
    """apg.py:277-304 (same field order)."""
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


def _f32(v):
    return np.float32(v)


def value_and_grad(cost_fn, x):
    xx = x.detach().clone().requires_grad_(True)
    c = cost_fn(xx)
    (g,) = torch.autograd.grad(c, xx)
    return c.detach(), g.detach()


def value(cost_fn, x):
    with torch.no_grad():
        return cost_fn(x.detach())


def init_opt_state(opt_var, optimizer_params):
    """apg.py:14-49."""
    s0 = _f32(optimizer_params["linesearch"]["init_stepsize"] if "linesearch" in optimizer_params
              else optimizer_params["stepsize"])
    mom = _f32(optimizer_params.get("beta_init", 0.25))
    return APGState(num_steps=1, num_iter_no_improv=0, momentum=mom, avg_momentum=mom,
                    avg_linesearch=_f32(1.0), stepsize=s0, avg_stepsize=s0, opt_cost=_f32(0.0),
                    x_opt=opt_var, curr_cost=_f32(0.0), init_cost=_f32(0.0), grad_sqr=_f32(0.0),
                    not_done=True, xk=opt_var, yk=opt_var, grad_yk=torch.zeros_like(opt_var))


def init_apg(x0, cost_fn, optimizer_params, proximal_fn=None, momentum=None, stepsize=None):
    """apg.py:53-113."""
    s0 = _f32(optimizer_params["linesearch"]["init_stepsize"] if "linesearch" in optimizer_params
              else optimizer_params["stepsize"])
    if proximal_fn is not None:
        x0 = proximal_fn(x0, stepsize=s0)
    c0, g0 = value_and_grad(cost_fn, x0)
    gsq = (g0 * g0).sum()
    if momentum is None:
        momentum = _f32(optimizer_params.get("beta_init", 0.25))
    if stepsize is None:
        stepsize = s0
    return APGState(num_steps=1, num_iter_no_improv=0, momentum=momentum, avg_momentum=momentum,
                    avg_linesearch=_f32(1.0), stepsize=stepsize, avg_stepsize=stepsize, opt_cost=c0,
                    x_opt=x0, curr_cost=c0, init_cost=c0, grad_sqr=gsq, not_done=True,
                    xk=x0, yk=x0, grad_yk=g0)


def wolfe_cond_violated(stepsize, coef, f_cur, f_next, grad_sqnorm):
    """apg.py:324-329 (eps = float32 machine epsilon)."""
    eps = np.finfo(np.float32).eps
    return bool(stepsize * coef * grad_sqnorm > f_cur - f_next + eps)


def armijo_line_search(cost_fn, xcurr, f_cur, grad, grad_sqnorm, stepsize, up):
    """apg.py:332-379."""
    coef, dec, smax = up["coef"], up["decrease_factor"], up["max_stepsize"]
    xnext = xcurr - stepsize * grad
    f_next = value(cost_fn, xnext)
    num_iter, violated, it = 1, True, 0
    while violated and it < up["maxls"]:
        violated = wolfe_cond_violated(stepsize, coef, f_cur, f_next, grad_sqnorm)
        if violated:  # update_search, apg.py:312-321
            stepsize = min(stepsize * dec, smax)
            xnext = xcurr - stepsize * grad
            f_next = value(cost_fn, xnext)
        num_iter += int(violated)
        it += 1
    return stepsize, xnext, f_next, num_iter


def one_step_apg(st, cost_fn, op, proximal_fn):
    """apg.py:164-274."""
    xpast, ycurr, cost_ycurr, grad_ycurr = st.xk, st.yk, st.curr_cost, st.grad_yk
    if "linesearch" in op:
        s = st.stepsize
        if op["linesearch"]["reset_option"] == "increase":
            s = min(s * op["linesearch"]["increase_factor"], op["linesearch"]["max_stepsize"])
        stepsize, xcurr, cost_xcurr, ls_steps = armijo_line_search(
            cost_fn, ycurr, cost_ycurr, grad_ycurr, st.grad_sqr, s, op["linesearch"])
    else:
        stepsize = st.stepsize
        xcurr = ycurr - stepsize * grad_ycurr
        if proximal_fn is None:
            cost_xcurr = value(cost_fn, xcurr)
        ls_steps = 1
    if proximal_fn is not None:
        xcurr = proximal_fn(xcurr, stepsize=stepsize)
        cost_xcurr = value(cost_fn, xcurr)
    vcurr = xpast + st.momentum * (xcurr - xpast)
    cost_vcurr = value(cost_fn, vcurr)
    x_le_v = bool(cost_xcurr <= cost_vcurr)
    ynext = xcurr if x_le_v else vcurr
    cost_ynext = cost_xcurr if x_le_v else cost_vcurr
    _, grad_ynext = value_and_grad(cost_fn, ynext)
    gsq = (grad_ynext * grad_ynext).sum()
    if op.get("moment_scale", None) is None:
        moment_next = st.num_steps / (st.num_steps + 3.0)
    else:
        ms = op["moment_scale"]
        moment_next = ms * st.momentum if x_le_v else min(st.momentum / ms, 1.0)
    stop_not_sat = bool(abs(cost_ynext - st.opt_cost) > op["atol"] + op["rtol"] * abs(st.opt_cost))
    improv = bool(st.opt_cost > cost_ynext)
    new_opt_cost = cost_ynext if improv else st.opt_cost
    new_x_opt = ynext if improv else st.x_opt
    new_noimprov = 0 if improv else st.num_iter_no_improv + 1
    stop_not_sat = stop_not_sat and (new_noimprov < op["max_no_improvement_iter"])
    n1 = st.num_steps + 1
    return APGState(num_steps=n1, num_iter_no_improv=new_noimprov, momentum=moment_next,
                    avg_momentum=(st.num_steps * st.avg_momentum + moment_next) / n1,
                    avg_linesearch=(st.num_steps * st.avg_linesearch + ls_steps) / n1,
                    stepsize=stepsize, avg_stepsize=(st.num_steps * st.avg_stepsize + stepsize) / n1,
                    opt_cost=new_opt_cost, curr_cost=cost_ynext, x_opt=new_x_opt, init_cost=st.init_cost,
                    grad_sqr=gsq, not_done=stop_not_sat, xk=xcurr, yk=ynext, grad_yk=grad_ynext)


def apg(state, cost_fn, op, proximal_fn=None, trace=None):
    """apg.py:115-161: at most ``max_iter`` iterations while ``not_done``."""
    it = 0
    while state.not_done and it < op["max_iter"]:
        state = one_step_apg(state, cost_fn, op, proximal_fn)
        if trace is not None:
            trace.append((float(state.stepsize), float(state.curr_cost), bool(state.not_done)))
        it += 1
    return state

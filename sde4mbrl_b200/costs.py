"""Stage / terminal cost descriptors evaluated inside the fused cost+gradient kernel.

The reference passes arbitrary Python callables ``_cost_fn(x, u, extra_args)`` that JAX traces
(``sde4mbrl/nsde.py:1463-1470,1506-1511``).  A CUDA kernel cannot call those, so the drop-in
recognises the cost families the reference actually ships and raises for anything else:

* :func:`one_step_cost_f` -- ``sde4mbrlExamples/rotor_uav/sde_mpc_design.py:73-132`` (all keys
  except ``urate_err``, which needs the unsupported control-augmented state);
* :class:`QuadraticCost` -- weighted tracking cost for models without a reference MPC script
  (cartpole, mass-spring-damper).
"""
import numpy as np

from . import _lib as L
from ._lib import CostDesc, Sde4Error


class StageCost:
    """A cost the kernel knows.  ``desc`` is the POD handed to ``sde4_cost_grad``."""

    def __init__(self, desc, cost_params=None):
        self.desc = desc
        self.cost_params = cost_params

    def __call__(self, *a, **k):
        raise Sde4Error("StageCost objects are evaluated by the CUDA kernel; they are not Python callables")


def _bcast(v, n):
    a = np.broadcast_to(np.asarray(v, dtype=np.float64), (n,))
    return [float(x) for x in a]


def one_step_cost_f(cost_params, n_u):
    """Descriptor of ``one_step_cost_f(x, u, xref, cost_params)`` (sde_mpc_design.py:73-132)."""
    cp = cost_params
    if "urate_err" in cp:
        raise Sde4Error("'urate_err' needs control_augmented_state, which has no compiled kernel")
    d = CostDesc()
    d.kind = L.COST_ROTOR_TRACK
    d.res_mult = float(cp.get("res_mult", 1.0))
    groups = (("perr", 0, (0, 1, 2)), ("verr", 1, (3, 4, 5)), ("qerr", 2, (7, 8, 9)), ("werr", 3, (10, 11, 12)))
    for key, bit, idx in groups:
        if key in cp:
            w = _bcast(cp[key], 3)
            for k, i in enumerate(idx):
                d.w_x[i] = w[k]
            if key + "_sig" in cp:
                d.sig_mask |= 1 << bit
                ws = _bcast(cp[key + "_sig"], 3)
                for k, i in enumerate(idx):
                    d.w_x_sig[i] = ws[k]
    if "uerr" in cp:
        w = _bcast(cp["uerr"], n_u)
        ur = _bcast(cp["uref"], n_u)
        for j in range(n_u):
            d.w_u[j] = w[j]
            d.uref[j] = ur[j]
        if "uerr_sig" in cp:
            d.sig_mask |= 1 << 4
            ws = _bcast(cp["uerr_sig"], n_u)
            for j in range(n_u):
                d.w_u_sig[j] = ws[j]
    if "res_sig" in cp:
        d.use_res_sig = 1
        d.res_sig_lo, d.res_sig_mult = float(cp["res_sig"][0]), float(cp["res_sig"][1])
    return StageCost(d, cp)


class QuadraticCost(StageCost):
    """``res_mult * (sum_i w_i (phi(x) - phi(xref))_i^2 + sum_j r_j (u - uref)_j^2)``;
    ``feat='cartpole'`` uses phi = [x, xdot, sin th, cos th, thdot] (cartpole_sde.py:49-53)."""

    def __init__(self, xerr, uerr, uref=None, feat="identity", res_mult=1.0):
        d = CostDesc()
        d.kind = L.COST_QUADRATIC
        d.feat = {"identity": 0, "cartpole": 1}[feat]
        d.res_mult = float(res_mult)
        for i, w in enumerate(xerr):
            d.w_x[i] = float(w)
        for j, w in enumerate(uerr):
            d.w_u[j] = float(w)
            d.uref[j] = float(uref[j]) if uref is not None else 0.0
        super().__init__(d, dict(feat=feat, xerr=list(xerr), uerr=list(uerr),
                                 uref=list(uref) if uref is not None else [0.0] * len(uerr), res_mult=res_mult))


def with_constraints(cost, constr, n_u):
    """Copy of ``cost`` carrying the state-bound penalty set-up (nsde.py:1392-1470)."""
    import ctypes as C
    d = CostDesc()
    C.memmove(C.byref(d), C.byref(cost.desc), C.sizeof(CostDesc))
    d.constr_mode = constr["mode"]
    d.n_slack = constr["n_slack"]
    d.constr_weight = float(constr["weight"])
    for i in range(L.MAX_NX):
        d.slack_col[i] = -1
    for k, i in enumerate(constr["state_idx"]):
        d.slack_col[i] = k
        d.pen[i] = float(constr["penalty"][k])
        d.lb[i] = float(constr["state_lb"][k])
        d.ub[i] = float(constr["state_ub"][k])
        d.slack_scale[i] = float(constr["slack_scaling"][k])
    out = StageCost(d, cost.cost_params)
    return out


def with_slew(cost, cost_params, n_u):
    """Copy of ``cost`` with the particle-independent control-slew term of ``augmented_cost``
    (sde_mpc_design.py:141-148) folded into the reduction kernel."""
    import ctypes as C
    d = CostDesc()
    C.memmove(C.byref(d), C.byref(cost.desc), C.sizeof(CostDesc))
    if "u_slew_coeff" in cost_params and "urate_err" not in cost_params:
        d.has_slew = 1
        for j, w in enumerate(_bcast(cost_params["u_slew_coeff"], n_u)):
            d.u_slew_coeff[j] = w
    if "u_slew_constr" in cost_params:
        raise Sde4Error("'u_slew_constr' has no compiled kernel (unused by the shipped configs)")
    if "state_slew_active" in cost_params:
        raise Sde4Error("'state_slew_active' has no compiled kernel (unused by the shipped configs)")
    return StageCost(d, cost.cost_params)

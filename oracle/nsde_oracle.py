"""CPU restatement of the reference's neural-SDE hot path -- test infrastructure only.

Written against torch-CPU so the same code gives float32 values (parity of
trajectories / costs) and float64 autograd gradients (parity of the adjoint
kernel).  Everything is vectorised over arbitrary leading batch dims (the
reference ``jax.vmap``s over particles, ``sde4mbrl/nsde.py:1027,1582``).

PARITY UNPINNED against the reference itself (no jax in this image; the
reference ships no golden vectors) -- see ``oracle/__init__.py``.

Reference map (all paths relative to /root/reference):
  hk.nets.MLP semantics ............ call sites nsde.py:376,387,674
  ControlledSDE.diffusion .......... sde4mbrl/nsde.py:298-449
  euler_maruyama ................... sde4mbrl/sde_solver.py:243-317
  compute_timesteps ................ sde4mbrl/nsde.py:15-40
  compute_cost_function ............ sde4mbrl/nsde.py:1485-1527,1576-1586
  MassSpringDamper ................. sde4mbrlExamples/mass_spring_damper/mass_spring_model.py:21-123
  CartPoleSDE ...................... sde4mbrlExamples/cartpole/cartpole_sde.py:20-106
  SDERotorModel .................... sde4mbrlExamples/rotor_uav/sde_rotor_model.py:121-458
  motor_model ...................... sde4mbrlExamples/rotor_uav/motor_models.py:21-95
  quaternion helpers ............... sde4mbrlExamples/rotor_uav/utils.py:13-63
  one_step_cost_f / augmented_cost . sde4mbrlExamples/rotor_uav/sde_mpc_design.py:73-167
"""
import math

import numpy as np
import torch

from . import threefry as tf


# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------
def _t(v, like):
    if isinstance(v, torch.Tensor):  # parameter trees of torch tensors keep their autograd history
        return v.to(like.dtype)
    return torch.as_tensor(np.array(v, dtype=np.float64), dtype=like.dtype, device=like.device)


def activation(name):
    """Name lookup as ``getattr(jnp, n) if hasattr(jnp, n) else getattr(jax.nn, n)``
    (nsde.py:377,388).  Only the names the shipped configs use."""
    if name == "tanh":
        return torch.tanh
    if name in ("swish", "silu"):
        return lambda x: x * torch.sigmoid(x)
    if name == "sigmoid":
        return torch.sigmoid
    if name == "relu":
        return torch.relu
    raise ValueError("unsupported activation " + name)


def mlp_apply(nn_params, prefix, x, act_name):
    """``hk.nets.MLP``: h <- act(h @ w + b) for all but the last layer, the last
    is linear; ``w[in,out]``, ``b[out]``.  ``prefix`` is the
    haiku module path without ``/linear_k``."""
    act = activation(act_name)
    n = 0
    while "%s/linear_%d" % (prefix, n) in nn_params:
        n += 1
    assert n > 0, "no MLP parameters under " + prefix
    h = x
    for k in range(n):
        lp = nn_params["%s/linear_%d" % (prefix, k)]
        w = _t(lp["w"], x)
        b = _t(lp["b"], x)
        h = h @ w + b
        if k < n - 1:
            h = act(h)
    return h


def compute_timesteps(params):
    """nsde.py:15-40."""
    horizon = params["horizon"]
    stepsize = params["stepsize"]
    num_short_dt = params.get("num_short_dt", horizon)
    assert num_short_dt <= horizon
    num_long_dt = horizon - num_short_dt
    short_step_dt = params.get("short_step_dt", stepsize)
    long_step_dt = params.get("long_step_dt", stepsize)
    return np.array([short_step_dt] * num_short_dt + [long_step_dt] * num_long_dt, dtype=np.float32)


# ---------------------------------------------------------------------------
# ControlledSDE restatement
# ---------------------------------------------------------------------------
class OracleSDE:
    """Restates ``ControlledSDE`` (nsde.py:134-595) for identity encoders."""

    module_name = "controlled_sde"  # haiku snake-case module name

    def __init__(self, params_model, nn_params, dtype=torch.float32):
        self.params = params_model
        self.nn = nn_params
        self.dtype = dtype
        self.n_x = params_model["n_x"]
        self.n_u = params_model["n_u"]
        self.n_y = params_model["n_y"]
        dd = params_model.get("diffusion_density_nn", None)
        self.has_density = dd is not None and "density_nn" in dd
        if dd is not None:
            self.noise_inputs = list(dd.get("indx_noise_in", range(self.n_x)))
            self.noise_outputs = list(dd.get("indx_noise_out", range(self.n_x)))

    # -- to be provided by the model ------------------------------------
    def compositional_drift(self, x, u):
        raise NotImplementedError

    def prior_diffusion(self, x, u):
        return _t(self.params["noise_prior_params"], x).expand(x.shape)

    def reduced_state(self, x):
        return x  # nsde.py:279-296

    def projection_fn(self, x):
        return x

    def scalar(self, name):
        return self.nn[self.module_name][name]

    # -- nsde.py:298-449 -------------------------------------------------
    def noise_aug_state_ctrl(self, x, u):
        red = self.reduced_state(x)
        z = red[..., self.noise_inputs] if len(self.noise_inputs) < self.n_x else red  # :361
        if self.params.get("control_dependent_noise", False):
            z = torch.cat([z, u.expand(*z.shape[:-1], u.shape[-1])], dim=-1)  # :362
        return z

    def density_net(self, z):
        dn = self.params["diffusion_density_nn"]["density_nn"]
        pre = "%s/~construct_diffusion_density_nn/density/~" % self.module_name
        return mlp_apply(self.nn, pre, z, dn.get("activation_fn", "tanh"))[..., 0]  # :387-390,407

    def diff_density(self, x, u):
        if not self.has_density:
            return torch.ones_like(x)  # :347
        dd = self.params["diffusion_density_nn"]
        z = self.noise_aug_state_ctrl(x, u)
        density = self.density_net(z)
        n_out = len(self.noise_outputs)
        sc = dd.get("scaler_nn", None)
        if sc is not None and sc.get("type", "none") != "none":
            assert sc.get("type", "scaler") == "scaler", "oracle restates the 'scaler' type only"
            sval = torch.exp(_t(self.scalar("scaler"), x)) * 1e-7  # :413
            coeff, offset = sval[:n_out], sval[n_out:]
        else:
            coeff = torch.zeros(n_out, dtype=x.dtype)
            offset = torch.zeros(n_out, dtype=x.dtype)
        aggr = dd["density_nn"].get("aggressiveness", 1.0)  # :398
        eta = torch.sigmoid((aggr + coeff) * density[..., None] - 7.0 + offset)  # :426
        if n_out < self.n_x:  # :363 -- non-noise outputs are filled with ONE
            full = torch.ones_like(x)
            full[..., self.noise_outputs] = eta
            return full
        return eta

    def diffusion(self, x, u):
        return self.diff_density(x, u) * self.prior_diffusion(x, u)  # :449


class MassSpringDamperOracle(OracleSDE):
    """mass_spring_model.py:21-123."""
    module_name = "mass_spring_damper"

    def residual(self, z):
        if self.params.get("prior", False):
            return torch.zeros(z.shape[:-1], dtype=z.dtype)
        rf = self.params["residual_forces"]
        pre = "%s/~init_residual_networks/Fres/~" % self.module_name
        return mlp_apply(self.nn, pre, z, rf["activation_fn"])[..., 0]

    def vector_field(self, x, u, Fres=0.0):
        ip = self.params["init_params"]
        m, k, b = ip["mass"], ip["kq_coeff"], ip["bqdot_coeff"]
        u0 = u[..., 0] if self.params["control"] else torch.zeros_like(x[..., 0])  # :54-55
        if self.params.get("ground_truth", False):
            acc = (u0 - k * x[..., 0] - b * x[..., 1]) / m
        else:
            acc = (u0 - k * x[..., 0] + Fres) / m
        return torch.stack([x[..., 1] + 0 * acc, acc], dim=-1)

    def compositional_drift(self, x, u):
        if self.params.get("ground_truth", False):
            return self.vector_field(x, u)
        if self.params.get("include_side_info", False):
            return self.vector_field(x, u, self.residual(x[..., 1:]))
        Fres = self.residual(x)
        return torch.stack([x[..., 1], Fres], dim=-1)


class CartPoleOracle(OracleSDE):
    """cartpole_sde.py:20-106."""
    module_name = "cart_pole_sde"

    def reduced_state(self, x):
        if "state_scaling" not in self.params:
            return x
        s = self.params["state_scaling"]
        return torch.stack([x[..., 1], torch.sin(x[..., 2]), torch.cos(x[..., 2]), x[..., 3]], dim=-1) / s[-1]  # :41

    def compositional_drift(self, x, u):
        s = self.params.get("state_scaling", [1.0] * self.n_x)
        xdot, th, thdot = x[..., 1], x[..., 2], x[..., 3]
        sin_t, cos_t = torch.sin(th), torch.cos(th)
        thdot_sc = thdot / s[3]
        uval = u[..., 0]
        sc = self.params["nn_out_scale"]
        rf = self.params["residual_forces"]
        pre = "%s/~init_residual_networks/res_forces/~" % self.module_name
        if not self.params.get("side_info", False):
            uu = uval.expand(sin_t.shape)
            out = mlp_apply(self.nn, pre, torch.stack([sin_t, cos_t, thdot_sc, uu], dim=-1), rf["activation_fn"])
            return torch.stack([xdot, float(sc[0]) * out[..., 0], thdot, float(sc[1]) * out[..., 1]], dim=-1)  # :71-74
        out = mlp_apply(self.nn, pre, torch.stack([sin_t, cos_t, thdot_sc], dim=-1), rf["activation_fn"])
        cn = self.params["control_nn"]
        prec = "%s/~init_residual_networks/control_nn/~" % self.module_name
        ce = mlp_apply(self.nn, prec, torch.stack([cos_t, cos_t ** 2], dim=-1), cn["activation_fn"]) * uval[..., None]
        xdd = float(sc[0]) * (out[..., 0] + ce[..., 0])
        tdd = float(sc[1]) * (out[..., 1] + ce[..., 1])
        return torch.stack([xdot, xdd, thdot, tdd], dim=-1)  # :77-83


# -- quaternion helpers (rotor_uav/utils.py:13-63), [w,x,y,z] -----------------
def quatmult(a, b):
    w1, x1, y1, z1 = a.unbind(-1)
    w2, x2, y2, z2 = b.unbind(-1)
    return torch.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
                        w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                        w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                        w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], dim=-1)


def quatinv(a):
    return a * torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=a.dtype)


def _vec2quat(v):
    return torch.cat([torch.zeros_like(v[..., :1]), v], dim=-1)


def quat_rotatevector(q, v):
    return quatmult(quatmult(q, _vec2quat(v)), quatinv(q))[..., 1:]


def quat_rotatevectorinv(q, v):
    return quatmult(quatinv(q), quatmult(_vec2quat(v), q))[..., 1:]


QUAD_MIXER = np.array([[-0.707, 0.707, 0.707, -0.707],
                       [-0.707, 0.707, -0.707, 0.707],
                       [-1.0, -1.0, 1.0, 1.0],
                       [1.0, 1.0, 1.0, 1.0]])
HEXA_MIXER = np.array([[-1.0, 1.0, 0.5, -0.5, -0.5, 0.5],
                       [0, 0, -0.866, 0.866, -0.866, 0.866],
                       [1.0, -1.0, 1.0, -1.0, -1.0, 1.0],
                       [1.0, 1.0, 1.0, 1.0, 1.0, 1.0]])


class RotorOracle(OracleSDE):
    """sde_rotor_model.py:121-458 (``control_augmented_state`` not restated: unused
    by every shipped config)."""
    module_name = "sde_rotor_model"

    def __init__(self, params_model, nn_params, dtype=torch.float32):
        super().__init__(params_model, nn_params, dtype)
        assert not params_model.get("control_augmented_state", False)
        self.init_params = params_model["init_params"]
        self.state_scaling = list(params_model.get("state_scaling", [1.0] * self.n_x))

    def get_param(self, name):
        """:146-158."""
        init_value = self.init_params.get(name, 1e-3)
        if name in self.init_params and self.params.get("fixed_init_params", False):
            return float(init_value)
        v = self.scalar(name)
        return v if isinstance(v, torch.Tensor) else float(v)  # tensors keep their autograd history (training loss)

    def reduced_state(self, x):
        if "state_scaling" in self.params:
            return x[..., :13] / max(self.state_scaling)  # :139-144
        return x

    def projection_fn(self, x):
        q = x[..., 6:10]
        qn = q / torch.linalg.norm(q, dim=-1, keepdim=True)  # :160-171
        return torch.cat([x[..., :6], qn, x[..., 10:]], dim=-1)

    def fm_model(self, u):
        """motor_models.py:31-95."""
        fm = self.init_params["fm_model"]
        order = fm.get("order", 1)
        const = fm.get("use_constant_term", False)
        assume_sym = fm.get("assume_sym", True)
        assert not fm.get("learn_mixer", False), "learn_mixer not restated (unused by shipped configs)"
        monos = [u ** k for k in range(order, 0, -1)]
        if const:
            monos.append(torch.ones_like(u))
        coeffs = [self.get_param("coeffMot%d" % i) for i in range(len(monos))]
        thrust_per_motor = sum(c * m for c, m in zip(coeffs, monos))
        mx = self.get_param("Lmx")
        my = mx if assume_sym else self.get_param("Lmy")
        mz = self.get_param("CoeffDrag")
        mixer = _t((QUAD_MIXER if u.shape[-1] == 4 else HEXA_MIXER).copy(), u)
        scale = torch.stack([torch.as_tensor(v, dtype=u.dtype) for v in (mx, my, mz, 1.0)])  # rows x Lmx, Lmy, CoeffDrag
        tm = thrust_per_motor @ (mixer * scale[:, None]).T
        return tm[..., 3], tm[..., :3]

    def _nn_in(self, x, v_b):
        s = self.state_scaling
        iv = torch.cat([v_b, x[..., 10:13]], dim=-1)
        return iv / _t([s[3], s[4], s[5], s[10], s[11], s[12]], x)  # :311-315

    def compute_force_residual(self, x, v_b):
        if not ("residual_forces" in self.params or "aero_drag_effect" in self.params):
            return None
        Fres = torch.zeros_like(v_b)
        if "residual_forces" in self.params:
            rf = self.params["residual_forces"]
            iv = self._nn_in(x, v_b)
            ai = rf.get("active_indx", None)
            if ai is not None:
                iv = iv[..., list(ai)]
            pre = "%s/~init_residual_networks/res_forces/~" % self.module_name
            Fres = Fres + mlp_apply(self.nn, pre, iv, rf["activation_fn"])
        if self.params.get("aero_drag_effect", False):
            kdx, kdy, kdz, kh = (self.scalar(n) if isinstance(self.scalar(n), torch.Tensor) else float(self.scalar(n))
                                 for n in ("aero_kdx", "aero_kdy", "aero_kdz", "aero_kh"))
            drag = torch.stack([-kdx * v_b[..., 0], -kdy * v_b[..., 1],
                                -kdz * v_b[..., 2] + kh * (v_b[..., 0] ** 2 + v_b[..., 1] ** 2)], dim=-1)  # :277
            Fres = Fres + drag
        return Fres

    def compute_moment_residual(self, x, v_b):
        if "residual_moments" not in self.params:
            return None
        rm = self.params["residual_moments"]
        iv = self._nn_in(x, v_b)
        ai = rm.get("active_indx", None)
        if ai is not None:
            iv = iv[..., list(ai)]
        pre = "%s/~init_residual_networks/res_moments/~" % self.module_name
        return mlp_apply(self.nn, pre, iv, rm["activation_fn"])

    def compositional_drift(self, x, u):
        q = x[..., 6:10]
        v = x[..., 3:6]
        w = x[..., 10:13]
        v_b = quat_rotatevectorinv(q, v)  # :439
        Fres = self.compute_force_residual(x, v_b)
        Mres = self.compute_moment_residual(x, v_b)
        thrust, moment = self.fm_model(u)
        thrust = thrust.expand(x.shape[:-1])
        moment = moment.expand(*x.shape[:-1], 3)
        # translational_dynamics :173-210
        m = self.get_param("mass")
        g = self.init_params.get("gravity", 9.81)
        F = torch.stack([torch.zeros_like(thrust), torch.zeros_like(thrust), thrust], dim=-1)
        if Fres is not None:
            F = F + Fres
        F = quat_rotatevector(q, F)
        a = F / m + _t([0.0, 0.0, -g], x)
        # rotational_dynamics :212-250
        I = torch.stack([torch.as_tensor(self.get_param(n), dtype=x.dtype) for n in ("Ixx", "Iyy", "Izz")])
        M = moment
        if Mres is not None:
            M = M + Mres
        M = M - torch.linalg.cross(w, I * w, dim=-1)
        ang_acc = M / I
        quat_dot = 0.5 * quatmult(q, _vec2quat(w))
        return torch.cat([v, a, quat_dot, ang_acc], dim=-1)  # :391


MODEL_CLASSES = {"mass_spring_damper": MassSpringDamperOracle,
                 "cart_pole_sde": CartPoleOracle,
                 "sde_rotor_model": RotorOracle}


def make_model(kind, params_model, nn_params, dtype=torch.float32):
    return MODEL_CLASSES[kind](params_model, nn_params, dtype)


# ---------------------------------------------------------------------------
# euler_maruyama (sde_solver.py:243-317), identity obs maps, array controls
# ---------------------------------------------------------------------------
def euler_maruyama(model, time_step, y0, us, dw):
    """One rollout batch.  ``y0[..., n_x]``, ``us[H, n_u]`` (or ``[..., H, n_u]``),
    ``dw[..., H, n_x]`` standard normals (NO sqrt(dt), sde_solver.py:299-301).
    Returns ``xevol[..., H+1, n_x]``."""
    H = len(time_step)
    if us.dim() == 1:
        us = us[None]  # :264-265
    assert us.shape[-2] == H
    x = y0.expand(*dw.shape[:-2], y0.shape[-1]) if y0.dim() == 1 else y0
    xs = [x]
    for t in range(H):
        u = us[..., t, :]
        drift = model.compositional_drift(x, u)
        diff = model.diffusion(x, u)
        xn = x + drift * float(time_step[t]) + diff * dw[..., t, :]  # :301
        x = model.projection_fn(xn)  # :304
        xs.append(x)
    return torch.stack(xs, dim=-2)


def sample_rollouts(model, y0, us, rng, num_particles, mode=tf.MODE_ORIGINAL, noise=None):
    """``create_sampling_fn.multi_sampling`` (nsde.py:1024-1028): N particles, shared
    y0/us.  Returns ``(xevol[N,H+1,n_x], dw[N,H,n_x])``."""
    ts = compute_timesteps(model.params)
    H = len(ts)
    if noise is None:
        noise = tf.rollout_noise(rng, num_particles, H, model.n_x, mode)
    dw = torch.as_tensor(noise).to(model.dtype)
    y0 = torch.as_tensor(np.asarray(y0)).to(model.dtype)
    us = torch.as_tensor(np.asarray(us)).to(model.dtype)
    return euler_maruyama(model, ts, y0, us, dw), dw


# ---------------------------------------------------------------------------
# MPC cost (nsde.py:1485-1527) + rotor stage cost (sde_mpc_design.py:73-167)
# ---------------------------------------------------------------------------
def rotor_one_step_cost(x, u, xref, cp):
    """``one_step_cost_f`` (sde_mpc_design.py:73-132) without ``urate_err``
    (control-augmented state only)."""
    def shaped(err, key):
        w = _t(cp[key], x)
        e = err * w
        if key + "_sig" in cp:
            e = torch.sigmoid(e - 7.0) * _t(cp[key + "_sig"], x)
        return e.sum(-1)
    tot = 0.0
    if "qerr" in cp:
        qe = quatmult(xref[..., 6:10].expand(x[..., 6:10].shape), quatinv(x[..., 6:10]))[..., 1:] ** 2
        tot = tot + shaped(qe, "qerr")
    if "perr" in cp:
        tot = tot + shaped((x[..., :3] - xref[..., :3]) ** 2, "perr")
    if "verr" in cp:
        tot = tot + shaped((x[..., 3:6] - xref[..., 3:6]) ** 2, "verr")
    if "werr" in cp:
        tot = tot + shaped((x[..., 10:13] - xref[..., 10:13]) ** 2, "werr")
    if "uerr" in cp:
        ue = (u - _t(cp["uref"], x)) ** 2
        tot = tot + shaped(ue, "uerr").expand(x.shape[:-1])
    if "res_sig" in cp:
        lo, mult = cp["res_sig"]
        return torch.tanh(tot - lo) * mult
    return tot * cp.get("res_mult", 1.0)


def quadratic_one_step_cost(x, u, xref, cp):
    """Generic tracking cost used for the non-rotor models (the reference has no
    cartpole / MSD MPC script): optional feature map then weighted
    squares.  ``cp = {'feat': 'identity'|'cartpole', 'xerr': [...], 'uerr': [...],
    'uref': [...], 'res_mult': r}``; features of cartpole are
    ``[x, xdot, sin th, cos th, thdot]`` (cartpole_sde.py:49-53)."""
    def feat(z):
        if cp.get("feat", "identity") == "cartpole":
            return torch.stack([z[..., 0], z[..., 1], torch.sin(z[..., 2]), torch.cos(z[..., 2]), z[..., 3]], dim=-1)
        return z
    e = ((feat(x) - feat(xref)) ** 2 * _t(cp["xerr"], x)).sum(-1)
    ue = ((u - _t(cp.get("uref", [0.0] * u.shape[-1]), x)) ** 2 * _t(cp["uerr"], x)).sum(-1)
    return (e + ue) * cp.get("res_mult", 1.0)


def constraint_setup(n_x, n_u, params_mpc):
    """``initialize_problem_constraints`` (nsde.py:75-130) in numpy."""
    has_ub = "input_constr" in params_mpc
    lb = ub = None
    if has_ub:
        lb = np.full(n_u, -np.inf, np.float32)
        ub = np.full(n_u, np.inf, np.float32)
        for idx, (l, h) in zip(params_mpc["input_constr"]["input_id"], params_mpc["input_constr"]["input_bound"]):
            lb[idx], ub[idx] = l, h
    has_xb = "state_constr" in params_mpc
    sc = dict(has_xbound=has_xb, slack_proximal=False, state_idx=None, penalty=None, state_lb=None,
              state_ub=None, weight=1.0, slack_scaling=1.0)
    if has_xb:
        d = params_mpc["state_constr"]
        sc["weight"] = d.get("constr_pen", 1)
        sc["state_idx"] = list(d["state_id"])
        sc["slack_proximal"] = d["slack_proximal"]
        sc["penalty"] = np.asarray(d["state_penalty"], np.float32)
        ss = np.asarray(d.get("slack_scaling", 1), np.float32) if sc["slack_proximal"] else np.float32(1)
        sc["slack_scaling"] = ss
        sc["state_lb"] = np.asarray([b[0] for b in d["state_bound"]], np.float32) / ss
        sc["state_ub"] = np.asarray([b[1] for b in d["state_bound"]], np.float32) / ss
    return (has_ub, lb, ub), sc


def constr_cost(x, slack, sc):
    """``constr_cost_noprox`` / ``constr_cost_withprox`` (nsde.py:1400-1419)."""
    if not sc["has_xbound"]:
        return 0.0
    xs = x[..., sc["state_idx"]]
    pen = _t(sc["penalty"], x)
    if sc["slack_proximal"]:
        d = xs - slack * _t(sc["slack_scaling"], x)
        return (d ** 2 * pen).sum(-1) * sc["weight"]
    diff = torch.cat([xs - _t(sc["state_ub"], x), _t(sc["state_lb"], x) - xs], dim=-1)
    pen2 = torch.cat([pen, pen])
    return (torch.where(diff > 0, diff ** 2, torch.zeros_like(diff)) * pen2).sum(-1) * sc["weight"]


def compute_cost(model, y0, opt, dw, stage_cost, xref, sc=None, terminal_cost=None):
    """``compute_cost_function`` vmapped over particles and averaged
    (nsde.py:1485-1527, 1576-1586).  ``opt[H, n_u + n_slack]``, ``dw[N,H,n_x]``,
    ``xref[H+1, n_x]``.  Returns ``(mean cost, xtraj[N,H+1,n_x+1])`` where the last
    column is the STAGE cost (nsde.py:1524-1525)."""
    ts = torch.as_tensor(compute_timesteps(model.params)).to(opt.dtype)
    n_u = model.n_u
    useq = opt[:, :n_u]
    xevol = euler_maruyama(model, ts.numpy(), y0, useq, dw)  # [N,H+1,n_x]
    useq_aug = torch.cat([useq, useq[-1:]], dim=0)  # :1509
    has_slack = sc is not None and sc["has_xbound"] and sc["slack_proximal"]
    total = stage_cost(xevol, useq_aug, xref)  # [N,H+1]
    if sc is not None and sc["has_xbound"]:
        if has_slack:
            slack = opt[:, n_u:]
            s0 = xevol[..., 0, sc["state_idx"]]
            slack_aug = torch.cat([s0[..., None, :], slack.expand(*s0.shape[:-1], *slack.shape)], dim=-2)  # :1510
            total = total + constr_cost(xevol, slack_aug, sc)
        else:
            total = total + constr_cost(xevol, None, sc)
    cost = (0.5 * ts * (total[..., 1:] + total[..., :-1])).sum(-1)  # :1515
    if terminal_cost is not None:
        cost = cost + terminal_cost(xevol[..., -1, :], xref[-1])
    xtraj = torch.cat([xevol, total[..., None]], dim=-1)
    return cost.mean(), xtraj


def u_slew_cost(opt, time_steps, cp, n_u):
    """Particle-independent slew term of ``augmented_cost`` (sde_mpc_design.py:141-148)."""
    if "u_slew_coeff" not in cp:
        return opt.new_zeros(())
    ts = torch.as_tensor(np.asarray(time_steps)).to(opt.dtype)[:-1].reshape(-1, 1)
    slew = (opt[1:, :n_u] - opt[:-1, :n_u]) / ts
    coeff = _t(np.broadcast_to(np.asarray(cp["u_slew_coeff"], np.float64), (n_u,)), opt).reshape(1, -1)
    return (slew ** 2 * coeff).sum() * cp.get("res_mult", 1.0)


# ---------------------------------------------------------------------------
# Training loss, data term (nsde.py:683-804 sample_for_loss_computation, :1185-1203 loss_fn)
# ---------------------------------------------------------------------------
def loss_noise(rng, B, N, H, n_x, mode=tf.MODE_ORIGINAL):
    """dw[B, N, H, n_x] with the training key chain: k_b = split(rng, B)[b] (nsde.py:1196),
    k_bn = split(k_b, N)[n] (:1187), rng_brownian = split(k_bn, 3)[0] (:709), then the solver's own
    split(rng_brownian)[0] (sde_solver.py:276) and normal(., (H, n_x)) (:283)."""
    kb = tf.split(rng, B, mode)
    out = np.zeros((B, N, H, n_x), np.float32)
    for b in range(B):
        kbn = tf.split(kb[b], N, mode)
        for n in range(N):
            r1 = tf.split(kbn[n], 3, mode)[0]
            r2 = tf.split(r1, 2, mode)[0]
            out[b, n] = tf.normal(r2, (H, n_x), mode)
    return out


def state_transform_for_loss(kind, x):
    """cartpole_sde.py:49-53; identity for the other models (nsde.py:530-542)."""
    if kind == "cart_pole_sde":
        return torch.stack([x[..., 0], x[..., 1], torch.sin(x[..., 2]), torch.cos(x[..., 2]), x[..., 3]], dim=-1)
    return x


def loss_step_indices(rng, B, N, H, m, mode=tf.MODE_ORIGINAL):
    """idx[B, N, m]: time indexes ``num_sample2consider`` keeps for sample (b, n) (nsde.py:753-760):
    ``choice(split(k_bn, 3)[1], H, (m,), replace=False)`` with k_bn as in ``loss_noise``."""
    kb = tf.split(rng, B, mode)
    out = np.zeros((B, N, m), np.int64)
    for b in range(B):
        kbn = tf.split(kb[b], N, mode)
        for n in range(N):
            out[b, n] = tf.choice_without_replacement(tf.split(kbn[n], 3, mode)[1], H, m, mode)
    return out


def loss_u_indices(rng, B, N, H, num_steps2data, mode=tf.MODE_ORIGINAL):
    """idx[B, N, H]: the data sub-step whose control sample (b, n) uses at solver step t under
    ``u_sampling_strategy: random`` (nsde.py:65, :735): ``randint(split(k_bn, 3)[1], (H,), 0, num_steps2data)`` with
    k_bn as in ``loss_noise`` (the same ``rng_sample2consider`` key that ``num_sample2consider`` reuses at :755)."""
    kb = tf.split(rng, B, mode)
    out = np.zeros((B, N, H), np.int64)
    for b in range(B):
        kbn = tf.split(kb[b], N, mode)
        for n in range(N):
            out[b, n] = tf.randint(tf.split(kbn[n], 3, mode)[1], (H,), 0, num_steps2data, mode)
    return out


def data_loss(model, kind, ymeas, uval, dw, obs_weights=1.0, num_steps2data=1, step_idx=None, u_idx=None):
    """mean over (batch, particles) of sum_t sum_i ((T(y_meas) - T(y_pred)) / w)^2 (nsde.py:766, :1203) with
    u_sampling_strategy 'first', or 'random' when ``u_idx[B, N, H]`` (``loss_u_indices``) is given.
    ymeas[B, Hd+1, n_y], uval[B, Hd, n_u], dw[B, N, H, n_x] (H = Hd / num_steps2data);
    ``step_idx[B, N, m]`` (``loss_step_indices``) restricts the sum to the kept time indexes."""
    ts = compute_timesteps(model.params)
    H = len(ts)
    B = ymeas.shape[0]
    y_values = ymeas[:, ::num_steps2data]
    uu = uval.reshape(B, H, num_steps2data, -1)
    if u_idx is None:
        u_values = uu[:, None, :, 0]
    else:  # arr[arange(H), rnd_indx] per sample (nsde.py:66)
        ii = torch.as_tensor(u_idx)[..., None, None].expand(B, dw.shape[1], H, 1, uu.shape[-1])
        u_values = torch.gather(uu[:, None].expand(B, dw.shape[1], H, num_steps2data, uu.shape[-1]), 3, ii)[:, :, :, 0]
    xs = euler_maruyama(model, ts, y_values[:, None, 0, :].expand(B, dw.shape[1], -1), u_values, dw)
    pred = state_transform_for_loss(kind, xs[:, :, 1:])
    meas = state_transform_for_loss(kind, y_values[:, None, 1:])
    w = _t(np.broadcast_to(np.asarray(obs_weights, np.float64), (pred.shape[-1],)), pred)
    e2 = (((meas - pred) / w) ** 2).sum(dim=-1)  # [B, N, H]
    if step_idx is not None:
        e2 = torch.gather(e2, 2, torch.as_tensor(step_idx))
    err = e2.sum(dim=-1)  # [B, N]
    return err.mean(dim=1).mean(), err


def weight_penalty(tree, nominal, coeffs):
    """sum_leaves coeff * ||p - p_nominal||^2 (nsde.py:1216-1219)."""
    tot = 0.0
    for mod in tree:
        for name in tree[mod]:
            tot = tot + ((tree[mod][name] - nominal[mod][name]) ** 2).sum() * coeffs[mod][name]
    return tot


# ---------------------------------------------------------------------------------------------
# distance-aware regulariser of the training loss
# ---------------------------------------------------------------------------------------------
def density_function(model, z):
    """sigmoid(density_net(z) - 7)   (sde4mbrl/nsde.py:395)."""
    return torch.sigmoid(model.density_net(z) - 7.0)


def mu_coeff(model, z, dl):
    """``mu_coeff_nn`` (nsde.py:657-680): constant, learned global scale, or exp(MLP) * mu_coeff."""
    lm = dl["learn_mucoeff"]
    mu0 = dl["mu_coeff"]
    if lm["type"] == "constant":
        return torch.as_tensor(mu0, dtype=z.dtype).expand(z.shape[:-1])
    if lm["type"] == "global":
        return (mu0 * torch.exp(_t(model.scalar("mu_coeff"), z))).expand(z.shape[:-1])
    pre = "%s/~mu_coeff_nn/mu_coeff/~" % model.module_name
    return torch.exp(mlp_apply(model.nn, pre, z, lm["activation_fn"])[..., 0]) * mu0


def density_loss_terms(model, ymeas, uval, rng, N, dl, mode=tf.MODE_ORIGINAL, step_idx=None, num_steps2data=1):
    """``density_loss`` (nsde.py:597-654) over the data points as ``sample_for_loss_computation`` calls it
    (:771-804): for minibatch element b and particle n the key is split(split(split(rng,B)[b],N)[n],3)[2]
    (:1196,:1187,:709), split over the Hd data steps (:781); each step splits once more and uses the SECOND key for
    the ball (:615).  ymeas[B, Hd+1, n_y], uval[B, Hd, n_u] torch tensors.
    ``step_idx[B, N, m]`` (``density_on_partial`` with ``num_sample2consider``, :794-797): the terms are taken at
    ``y_values[:-1][idx]``, ``u_values[idx]`` (strategy 'first') with the keys ``split(., Hd)[idx]``.
    Returns grad_norm[B,N] (mean over t), sconvex[B,N] (sum over t), mu[B,N] (mean), density[B,N] (mean)."""
    B, Hd = ymeas.shape[0], ymeas.shape[1] - 1
    ns = int(dl["ball_nsamples"])
    kb = tf.split(rng, B, mode)
    out = [[None] * N for _ in range(B)]
    if step_idx is not None:
        y_values = ymeas[:, ::num_steps2data]
        u_values = uval.reshape(B, -1, num_steps2data, uval.shape[-1])[:, :, 0]
        for b in range(B):
            kbn = tf.split(kb[b], N, mode)
            for n in range(N):
                idx = torch.as_tensor(step_idx[b, n])
                z = model.noise_aug_state_ctrl(y_values[b, :-1][idx], u_values[b][idx]).detach().requires_grad_(True)
                dim = z.shape[-1]
                den = density_function(model, z)
                (gz,) = torch.autograd.grad(den.sum(), z, create_graph=True)
                mu = mu_coeff(model, z, dl)
                radius = _t(np.broadcast_to(np.asarray(dl["ball_radius"], np.float64), (dim,)).copy(), z)
                kt = tf.split(tf.split(kbn[n], 3, mode)[2], Hd, mode)
                ball = torch.stack([_t(tf.normal(tf.split(kt[int(t)], 2, mode)[1], (ns, dim), mode), z) for t in idx]) * radius
                den_ball = density_function(model, z[:, None, :] + ball)
                cond = den_ball - den[:, None] - (gz[:, None, :] * ball).sum(-1) - 0.5 * mu[:, None] * (ball ** 2).sum(-1)
                sconvex = (torch.clamp(cond, max=0.0) ** 2).sum(-1)
                out[b][n] = ((gz ** 2).sum(-1).mean(), sconvex.sum(), mu.mean(), den.mean())
        stack = lambda i: torch.stack([torch.stack([out[b][n][i] for n in range(N)]) for b in range(B)])
        return stack(0), stack(1), stack(2), stack(3)
    for b in range(B):
        z = model.noise_aug_state_ctrl(ymeas[b, :-1], uval[b]).detach().requires_grad_(True)  # [Hd, dim]
        dim = z.shape[-1]
        den = density_function(model, z)
        (gz,) = torch.autograd.grad(den.sum(), z, create_graph=True)
        mu = mu_coeff(model, z, dl)
        radius = _t(np.broadcast_to(np.asarray(dl["ball_radius"], np.float64), (dim,)).copy(), z)
        kbn = tf.split(kb[b], N, mode)
        for n in range(N):
            kt = tf.split(tf.split(kbn[n], 3, mode)[2], Hd, mode)
            ball = torch.stack([_t(tf.normal(tf.split(kt[t], 2, mode)[1], (ns, dim), mode), z) for t in range(Hd)]) * radius
            den_ball = density_function(model, z[:, None, :] + ball)  # [Hd, ns]
            cond = den_ball - den[:, None] - (gz[:, None, :] * ball).sum(-1) - 0.5 * mu[:, None] * (ball ** 2).sum(-1)
            sconvex = (torch.clamp(cond, max=0.0) ** 2).sum(-1)
            out[b][n] = ((gz ** 2).sum(-1).mean(), sconvex.sum(), mu.mean(), den.mean())
    stack = lambda i: torch.stack([torch.stack([out[b][n][i] for n in range(N)]) for b in range(B)])
    return stack(0), stack(1), stack(2), stack(3)


def training_total(loss_params, dl, loss_data, grad_norm, sconvex, mu, w_loss):
    """``total_sum`` of ``create_model_loss_fn.loss_fn`` (nsde.py:1203-1254)."""
    lgd, lsc, mum = grad_norm.mean(1).mean(), sconvex.mean(1).mean(), mu.mean(1).mean()
    mult = loss_params.get("pen_scvex_mult", 1.0)
    tot = 0.0
    if loss_params.get("pen_data", 0) > 0:
        tot = tot + loss_data * loss_params["pen_data"]
    if loss_params.get("pen_grad_density", 0) > 0:
        p_mu = dl["mu_coeff"]
        tot = tot + loss_params["pen_grad_density"] * p_mu * dl.get("grad_scaler", 1.0 / p_mu) * lgd * mult
    if loss_params.get("pen_density_scvex", 0) > 0:
        tot = tot + loss_params["pen_density_scvex"] * lsc * mult
    lmu = 0.0
    if loss_params.get("pen_mu_coeff", 0) > 0:
        kind = loss_params.get("pen_mu_type", "quad_inv")
        lmu = {"quad_inv": lambda: 1.0 / mum ** 2, "lin_inv": lambda: 1.0 / mum,
               "exp_inv": lambda: torch.exp(-mum * loss_params["pen_mu_temp"])}[kind]()
        tot = tot + lmu * loss_params["pen_mu_coeff"] * mult
    if loss_params.get("pen_weights", 0) > 0:
        tot = tot + loss_params["pen_weights"] * w_loss
    return tot, {"gradDensity": lgd, "densitySCVEX": lsc, "muCoeff": mum, "lossMuCoeff": lmu}


# ---------------------------------------------------------------------------------------------
# the reference's commented-out Stratonovich schemes (never registered, sde_solver.py:321-325): restated for the
# rollout-only extension of the CUDA kernel.  xi[..., H, n_x] are the standard normals of the rollout.
# ---------------------------------------------------------------------------------------------
def stratonovich_heun(model, time_step, y0, us, xi):
    """sde_solver.py:9-57: dW = sqrt(dt) xi (:37); predictor with (f0, g0), corrector with the averages (:44-49);
    projection after the step (:50)."""
    H = len(time_step)
    x = y0.expand(*xi.shape[:-2], y0.shape[-1]) if y0.dim() == 1 else y0
    xs = [x]
    for t in range(H):
        dt = float(time_step[t])
        u = us[..., t, :]
        dw = np.sqrt(np.float32(dt)).item() * xi[..., t, :]
        f0, g0 = model.compositional_drift(x, u), model.diffusion(x, u)
        xb = x + f0 * dt + g0 * dw
        f1, g1 = model.compositional_drift(xb, u), model.diffusion(xb, u)
        x = model.projection_fn(x + 0.5 * (f0 + f1) * dt + 0.5 * (g0 + g1) * dw)
        xs.append(x)
    return torch.stack(xs, dim=-2)


def stratonovich_milstein(model, time_step, y0, us, xi):
    """sde_solver.py:114-177 (derivative-free): xbar = x + f0 dt + g0 sqrt(dt) (:158-161);
    x+ = x + f0 dt + g0 sqrt(dt) xi + (0.5/sqrt(dt)) (g(xbar) - g0) xi^2 (:165); projection (:166)."""
    H = len(time_step)
    x = y0.expand(*xi.shape[:-2], y0.shape[-1]) if y0.dim() == 1 else y0
    xs = [x]
    for t in range(H):
        dt = float(time_step[t])
        sq = np.sqrt(np.float32(dt)).item()
        u = us[..., t, :]
        z = xi[..., t, :]
        f0, g0 = model.compositional_drift(x, u), model.diffusion(x, u)
        xt = x + f0 * dt
        g1 = model.diffusion(xt + g0 * sq, u)
        x = model.projection_fn(xt + g0 * (sq * z) + (0.5 / sq) * (g1 - g0) * z ** 2)
        xs.append(x)
    return torch.stack(xs, dim=-2)

"""The UNMODIFIED reference (wuwushrek/sde4mbrl) driven through its own public API on the host CPU.

Used by ``tests/test_ref_crosscheck.py`` (pins the oracle to the reference itself) and by ``bench.py --impl reference``
(the jit-compiled JAX arm the >= 100x target is defined on).  Everything here needs ``jax`` + ``dm-haiku`` and a
copy of the reference on disk; neither is present in the build image (no network), so every
entry point reports ``available() == (False, reason)`` there and the callers skip / fall back.  Nothing in
``sde4mbrl_b200/`` imports this file.

Reference call sites exercised:
  sde4mbrl/nsde.py:987-1030     create_sampling_fn -> multi_sampling
  sde4mbrl/nsde.py:1341-1597    create_online_cost_sampling_fn -> multi_sampling, jax.value_and_grad of it
  sde4mbrlExamples/rotor_uav/sde_mpc_design.py:73-167   one_step_cost_f, augmented_cost
  sde4mbrl/apg.py:53-161        init_apg, apg
"""
import copy
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_STATE = {"ready": False}


def reference_root():
    for cand in (os.environ.get("SDE4_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "sde4mbrl")) and os.path.isdir(os.path.join(cand, "sde4mbrlExamples")):
            return cand
    return None


def available():
    """(True, "") when the reference can be imported and run here, else (False, one-line reason)."""
    try:
        import jax  # noqa: F401
        import haiku  # noqa: F401
    except Exception as ex:  # ModuleNotFoundError in the build image
        return False, "jax / dm-haiku not importable (%s)" % (ex,)
    if reference_root() is None:
        return False, "no copy of the reference found (SDE4_REFERENCE_ROOT, baseline/_ref, /root/reference)"
    return True, ""


def _setup():
    if _STATE["ready"]:
        return
    os.environ.setdefault("JAX_PLATFORM_NAME", "cpu")  # as the reference's own MPC scripts do (hexa_ahg/test_mpc.py:2-4)
    root = reference_root()
    if root not in sys.path:
        sys.path.insert(0, root)
    _STATE["ready"] = True


def rng_mode():
    """0 / 1 = the jax.random stream layout of the installed jax (jax_threefry_partitionable False / True)."""
    _setup()
    import jax
    return int(bool(jax.config.jax_threefry_partitionable))


def model_class(kind):
    _setup()
    if kind in ("hexa", "rotor", "sde_rotor_model"):
        from sde4mbrlExamples.rotor_uav.sde_rotor_model import SDERotorModel
        return SDERotorModel
    if kind in ("cartpole", "cart_pole_sde"):
        from sde4mbrlExamples.cartpole.cartpole_sde import CartPoleSDE
        return CartPoleSDE
    if kind in ("msd", "mass_spring_damper"):
        from sde4mbrlExamples.mass_spring_damper.mass_spring_model import MassSpringDamper
        return MassSpringDamper
    raise KeyError(kind)


def _tree(nn_params):
    import jax.numpy as jnp
    return {m: {k: jnp.asarray(v) for k, v in leaves.items()} for m, leaves in nn_params.items()}


def rollout(kind, params_model, nn_params, y0, u, key, num_samples):
    """jit(create_sampling_fn(...).multi_sampling)(nn_params, y0, u, key) -> numpy xevol [N, H+1, n_x]."""
    _setup()
    import jax
    import jax.numpy as jnp
    import numpy as np
    from sde4mbrl.nsde import create_sampling_fn
    pm = copy.deepcopy(params_model)
    _, multi_sampling = create_sampling_fn(pm, sde_constr=model_class(kind), num_samples=num_samples)
    fn = jax.jit(multi_sampling)
    xevol, _, _ = fn(_tree(nn_params), jnp.asarray(y0), jnp.asarray(u), jnp.asarray(np.asarray(key, np.uint32)))
    return np.asarray(xevol)


def rotor_cost_fn(params_model, params_mpc, cost_params, nn_params, y0, key, xref):
    """The reference's ``cost_xt`` (sde_mpc_design.py:193-217) for one multirotor problem, as a jax function of
    ``opt[H, n_opt]`` returning ``(cost, xtraj)``; also returns the proximal function and the time steps."""
    _setup()
    import jax.numpy as jnp
    import numpy as np
    from sde4mbrl.nsde import compute_timesteps, create_online_cost_sampling_fn
    from sde4mbrlExamples.rotor_uav.sde_mpc_design import augmented_cost, one_step_cost_f
    pm, mpc = copy.deepcopy(params_model), copy.deepcopy(params_mpc)
    _, multi_cost, vmapped_prox, _, _ = create_online_cost_sampling_fn(pm, mpc, sde_constr=model_class("hexa"))
    nn = _tree(nn_params)
    y = jnp.asarray(y0)
    k = jnp.asarray(np.asarray(key, np.uint32))
    xr = jnp.asarray(xref)
    cp = copy.deepcopy(cost_params)
    red_cost_f = lambda _x, _u, extra_args: one_step_cost_f(_x, _u, extra_args, cp)
    ts = jnp.array(compute_timesteps(pm))
    n_u = pm["n_u"]
    state = {}

    def inner(u_params):
        c, xtraj = multi_cost(nn, y, u_params, k, red_cost_f, _terminal_cost=None, extra_dyn_args=None,
                              extra_cost_args=xr, extra_cost_terminal_args=xr[-1])
        state["xtraj"] = xtraj
        return c, xtraj

    def cost_xt(u_params):
        return augmented_cost(u_params, inner, cp, ts, n_u)
    return cost_xt, vmapped_prox, np.asarray(ts)


def rotor_cost_value_and_grad(params_model, params_mpc, cost_params, nn_params, y0, opt, key, xref):
    """jit(jax.value_and_grad(cost_xt))(opt) -> (cost, grad[H, n_opt]) as numpy (apg.py:82,224)."""
    _setup()
    import jax
    import jax.numpy as jnp
    import numpy as np
    cost_xt, _, _ = rotor_cost_fn(params_model, params_mpc, cost_params, nn_params, y0, key, xref)
    c, g = jax.jit(jax.value_and_grad(cost_xt))(jnp.asarray(opt))
    return float(c), np.asarray(g)


def rotor_apg_solve(params_model, params_mpc, cost_params, apg_params, nn_params, y0, opt0, key, xref):
    """init_apg + apg (apg.py:53-161) on the reference's cost; returns the final APGState as a dict of numpy values."""
    _setup()
    import jax
    import jax.numpy as jnp
    import numpy as np
    from sde4mbrl.apg import apg, init_apg
    cost_xt, vmapped_prox, _ = rotor_cost_fn(params_model, params_mpc, cost_params, nn_params, y0, key, xref)
    prox = None if vmapped_prox is None else (lambda x, stepsize=None: vmapped_prox(x))
    cfg = copy.deepcopy(apg_params)

    def solve(x0):
        st = init_apg(x0, cost_xt, cfg)
        return apg(st, cost_xt, cfg, proximal_fn=prox)
    st = jax.jit(solve)(jnp.asarray(opt0))
    return {f: np.asarray(getattr(st, f)) for f in st._fields}


def time_rotor_value_and_grad(params_model, params_mpc, cost_params, nn_params, y0, opt, key, xref, steps, warmup):
    """Seconds per jit-compiled value+gradient evaluation (block_until_ready), after ``warmup`` untimed calls."""
    _setup()
    import jax
    import jax.numpy as jnp
    cost_xt, _, _ = rotor_cost_fn(params_model, params_mpc, cost_params, nn_params, y0, key, xref)
    fn = jax.jit(jax.value_and_grad(cost_xt))
    o = jnp.asarray(opt)
    for _ in range(max(1, warmup)):
        jax.block_until_ready(fn(o))
    out = []
    for _ in range(steps):
        t0 = time.perf_counter()
        jax.block_until_ready(fn(o))
        out.append(time.perf_counter() - t0)
    return out

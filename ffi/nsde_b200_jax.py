"""jax.ffi binding of libsde4b200.so -- the Python half of the XLA custom-call layer (C++ half: ffi/sde4_xla_ffi.cc).

What a maintainer of the reference adds to run its own jitted host code on the CUDA kernels: the three factories
below return functions with the signatures of the reference's closures (sde4mbrl/nsde.py:1024-1028, :1576-1586,
:1185-1203) that can be called under ``jax.jit`` / differentiated with ``jax.grad`` -- the rollout, the MPC cost with
a ``jax.custom_vjp`` whose backward pass is the gradient the fused kernel already produced (apg.py:82,224), and the
training-loss data term with a custom VJP w.r.t. the haiku parameter tree (train_sde.py:248-262).

Needs jax (with ``jax.ffi``) and the shim built against jaxlib's headers
(``python -c 'import __graft_entry__ as g; g.build()'`` does it when ``jax.ffi.include_dir()`` exists).  jax is not
installed in this repository's build image, so this module is exercised there only as far as it does not need jax
(``TreePacker``, tests/test_ffi_shim.py); the tested binding in that image is ctypes (sde4mbrl_b200/_lib.py).
"""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_PATH = os.path.join(ROOT, "sde4mbrl_b200", "libsde4_xla_ffi.so")
TARGETS = {"sde4_rollout": "Sde4Rollout", "sde4_cost_grad": "Sde4CostGrad", "sde4_loss_grad": "Sde4LossGrad"}
_registered = False


def _bytes_attr(pod):
    """A POD descriptor (ctypes.Structure) as the uint8 array attribute the handlers decode."""
    return np.frombuffer(bytes(pod), dtype=np.uint8)


class TreePacker:
    """haiku parameter tree <-> the packed float32 block of ``sde4_param_layout``, as ONE gather.

    ``model.pack`` is affine in the leaves (weights are copied / transposed / zero padded, a few model scalars are
    duplicated or fixed by the YAML), so probing it once with distinct leaf values gives, for every slot of the block,
    the flat leaf element it holds (or a constant).  ``pack`` is then ``where(src >= 0, flat[src], const)`` in any array
    namespace -- numpy here, ``jax.numpy`` under ``jit``, where autodiff of the gather also gives the scatter-add that
    turns the kernel's packed parameter gradient back into a tree."""

    def __init__(self, model, example_tree):
        self.model = model
        self.desc = model.describe()
        self.keys = [(m, k) for m in sorted(example_tree) for k in sorted(example_tree[m])]
        self.shapes = [np.shape(example_tree[m][k]) for m, k in self.keys]
        sizes = [int(np.prod(s)) if len(s) else 1 for s in self.shapes]
        self.offsets = np.concatenate([[0], np.cumsum(sizes)])
        probe, zero = {}, {}
        for (m, k), shp, o, n in zip(self.keys, self.shapes, self.offsets[:-1], sizes):
            # distinct positive values, exactly representable in float32 for the block sizes at hand (< 2^24)
            probe.setdefault(m, {})[k] = (np.arange(n, dtype=np.float64) + o + 1).astype(np.float32).reshape(shp)
            zero.setdefault(m, {})[k] = np.zeros(shp, np.float32)
        assert self.offsets[-1] < (1 << 24)
        self.const = np.asarray(model.pack(zero, self.desc), np.float32)
        marked = np.asarray(model.pack(probe, self.desc), np.float32) - self.const
        self.src = np.where(marked != 0, np.rint(marked).astype(np.int64) - 1, -1)

    def flatten(self, tree, xp=np):
        return xp.concatenate([xp.reshape(xp.asarray(tree[m][k], dtype=xp.float32), (-1,)) for m, k in self.keys])

    def pack(self, tree, xp=np):
        flat = self.flatten(tree, xp)
        return xp.where(self.src >= 0, flat[np.maximum(self.src, 0)], self.const)

    def unflatten(self, flat):
        out = {}
        for (m, k), shp, o, e in zip(self.keys, self.shapes, self.offsets[:-1], self.offsets[1:]):
            out.setdefault(m, {})[k] = flat[int(o):int(e)].reshape(shp)
        return out


def register(shim_path=SHIM_PATH):
    """Load the shim (and libsde4b200.so under it) and register its handlers as XLA FFI targets on the CUDA platform."""
    global _registered
    if _registered:
        return
    import jax
    from sde4mbrl_b200 import _lib as L
    ctypes.CDLL(L.LIB_PATH, mode=ctypes.RTLD_GLOBAL)  # the shim resolves sde4_* against it
    shim = ctypes.CDLL(shim_path)
    shim.sde4_xla_ffi_available.restype = ctypes.c_int
    if shim.sde4_xla_ffi_available() != 1:
        raise RuntimeError("libsde4_xla_ffi.so was built WITHOUT the XLA FFI headers (stub); rebuild it where "
                           "jax.ffi.include_dir() exists")
    for target, symbol in TARGETS.items():
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(shim, symbol)), platform="CUDA")
    _registered = True


def _rng_mode():
    import jax
    return int(bool(jax.config.jax_threefry_partitionable))


def _key_data(rng):
    import jax
    import jax.numpy as jnp
    if jnp.issubdtype(rng.dtype, jax.dtypes.prng_key):  # new-style typed keys
        rng = jax.random.key_data(rng)
    return rng.astype(jnp.uint32).reshape(-1, 2)


def create_sampling_fn(params_model, sde_constr, num_samples=None):
    """``create_sampling_fn`` (nsde.py:987-1030): returns ``multi_sampling(nn_params, y, u, rng) -> (xevol, yevol, uevol)``
    for use under ``jax.jit``.  ``sde_constr`` is one of sde4mbrl_b200.sde_models' classes (the static description of the
    reference's model class: which kernel specialisation to launch)."""
    import jax
    import jax.numpy as jnp
    from sde4mbrl_b200 import _lib as L
    from sde4mbrl_b200.nsde import compute_timesteps
    register()
    model = sde_constr(params_model)
    desc = model.describe()
    N = int(params_model["num_particles"] if num_samples is None else num_samples)
    dt = np.asarray(compute_timesteps(params_model), np.float32)
    H, n_x, n_u = dt.shape[0], model.n_x, model.n_u
    noise_mode = L.NOISE_THREEFRY if model.is_noisy() else L.NOISE_NONE
    tp = TreePacker(model, _example_tree(model))

    def multi_sampling(nn_params, y, u, rng, extra_args=None):
        assert extra_args is None, "extra_args must be None (no shipped model uses them)"
        params = tp.pack(nn_params, jnp)
        u = jnp.broadcast_to(jnp.asarray(u, jnp.float32).reshape(-1, n_u), (H, n_u)) if jnp.ndim(u) == 1 else jnp.asarray(u, jnp.float32)
        call = jax.ffi.ffi_call("sde4_rollout", jax.ShapeDtypeStruct((H + 1, n_x, N), jnp.float32), vmap_method="sequential")
        xbuf = call(params, jnp.asarray(y, jnp.float32).reshape(1, n_x), u, jnp.asarray(dt), _key_data(rng),
                    model_desc=_bytes_attr(desc), num_particles=np.int32(N), rng_mode=np.int32(_rng_mode()),
                    noise_mode=np.int32(noise_mode), solver=np.int32(0))
        xevol = jnp.transpose(xbuf, (2, 0, 1))  # [N, H + 1, n_x]
        return xevol, xevol, jnp.broadcast_to(u, (N, H, n_u))  # identity observation map (sde_solver.py:307)
    return model.init_params(0), multi_sampling


def _example_tree(model):
    return model.init_params(0)


def create_online_cost_sampling_fn(params_model, params_mpc, sde_constr, stage_cost, terminal_cost=None):
    """``create_online_cost_sampling_fn`` (nsde.py:1341-1597) for the cost families the kernel knows (``costs.StageCost``
    descriptors; the reference passes Python callables here).  Returns
    ``multi_sampling(nn_params, y, opt_params, rng, xref) -> (mean_cost, xtraj)``, differentiable in ``opt_params``
    through a ``jax.custom_vjp``: the fused kernel returns the gradient with the value, the backward pass only scales it."""
    import jax
    import jax.numpy as jnp
    from sde4mbrl_b200 import _lib as L
    from sde4mbrl_b200 import nsde as host
    register()
    pm = dict(params_model)
    pm["num_particles"] = params_mpc.get("num_particles", pm.get("num_particles", 1))
    pm["horizon"] = params_mpc.get("horizon", pm.get("horizon", 1))
    for k in ("num_short_dt", "short_step_dt", "long_step_dt"):
        pm[k] = params_mpc[k]
    model = sde_constr(pm)
    desc = model.describe()
    N = int(pm["num_particles"])
    dt = np.asarray(host.compute_timesteps(pm), np.float32)
    H, n_x = dt.shape[0], model.n_x
    n_opt = model.n_u + stage_cost.desc.n_slack
    noise_mode = L.NOISE_THREEFRY if model.is_noisy() else L.NOISE_NONE
    lib = L.load()
    ws_bytes = int(lib.sde4_cost_workspace_bytes(ctypes.byref(desc), 1, N, H, stage_cost.desc.n_slack, 1, 1))
    term_attr = _bytes_attr(terminal_cost.desc) if terminal_cost is not None else np.zeros(0, np.uint8)
    tp = TreePacker(model, _example_tree(model))

    def _call(params, y, opt, xref, key):
        out_types = (jax.ShapeDtypeStruct((1,), jnp.float32), jax.ShapeDtypeStruct((1, H, n_opt), jnp.float32),
                     jax.ShapeDtypeStruct((H + 1, n_x + 1, N), jnp.float32), jax.ShapeDtypeStruct((ws_bytes,), jnp.uint8))
        call = jax.ffi.ffi_call("sde4_cost_grad", out_types, vmap_method="sequential")
        cost, grad, xtraj, _ = call(params, y.reshape(1, n_x), opt.reshape(1, H, n_opt), jnp.asarray(dt),
                                    xref.reshape(1, H + 1, n_x), key, model_desc=_bytes_attr(desc),
                                    stage_desc=_bytes_attr(stage_cost.desc), terminal_desc=term_attr,
                                    num_particles=np.int32(N), rng_mode=np.int32(_rng_mode()),
                                    noise_mode=np.int32(noise_mode))
        return cost[0], grad[0], jnp.transpose(xtraj, (2, 0, 1))  # xtraj [N, H + 1, n_x + 1]

    @jax.custom_vjp
    def cost_fn(params, y, opt, xref, key):
        c, _, xtraj = _call(params, y, opt, xref, key)
        return c, xtraj

    def cost_fwd(params, y, opt, xref, key):
        c, g, xtraj = _call(params, y, opt, xref, key)
        return (c, xtraj), (g, params, y, xref, key)

    def cost_bwd(res, cts):
        g, params, y, xref, key = res
        ct_cost, _ = cts  # the trajectories are returned for inspection, not differentiated (apg.py differentiates the cost)
        zeros = lambda a: jnp.zeros_like(a)
        return zeros(params), zeros(y), ct_cost * g, zeros(xref), np.zeros(key.shape, jax.dtypes.float0)
    cost_fn.defvjp(cost_fwd, cost_bwd)

    def multi_sampling(nn_params, y, opt_params, rng, xref):
        assert opt_params.ndim == 2 and opt_params.shape == (H, n_opt), "Shape of the opt params do not match"
        return cost_fn(tp.pack(nn_params, jnp), jnp.asarray(y, jnp.float32), jnp.asarray(opt_params, jnp.float32),
                       jnp.asarray(xref, jnp.float32), _key_data(rng))
    return model.init_params(0), multi_sampling


def create_model_loss_fn(params_model, loss_params, sde_constr):
    """Data term of ``create_model_loss_fn(...).loss_fn`` (nsde.py:1185-1203) with a custom VJP w.r.t. the haiku tree:
    ``loss_fn(nn_params, y[B, H+1, n_y], u[B, H, n_u], rng) -> mean_b loss_b``; ``jax.grad(loss_fn)`` returns a tree."""
    import jax
    import jax.numpy as jnp
    from sde4mbrl_b200 import _lib as L
    from sde4mbrl_b200 import costs
    register()
    pm = dict(params_model)
    pm["horizon"] = int(loss_params["horizon"])
    pm["num_particles"] = int(loss_params["num_particles"])
    model = sde_constr(pm)
    desc = model.describe()
    N, H, n_x = pm["num_particles"], pm["horizon"], model.n_x
    dt = np.full(H, float(loss_params.get("stepsize", pm["stepsize"])), np.float32)
    ow = np.asarray(loss_params["obs_weights"], np.float64)
    feat = "cartpole" if len(ow) == n_x + 1 else "identity"  # cartpole_sde.py:49-53 transforms theta into (sin, cos)
    data_cost = costs.QuadraticCost(list(1.0 / ow ** 2), [0.0] * model.n_u, [0.0] * model.n_u, feat=feat)
    tp = TreePacker(model, _example_tree(model))
    lib = L.load()
    noise_mode = L.NOISE_THREEFRY if model.is_noisy() else L.NOISE_NONE

    def _call(params, y, u, keys):
        B = y.shape[0]
        ws = int(lib.sde4_loss_workspace_bytes(ctypes.byref(desc), B, N, H)) + ((B * n_x * 4 + 255) & ~255) + 256
        out_types = (jax.ShapeDtypeStruct((B,), jnp.float32), jax.ShapeDtypeStruct(params.shape, jnp.float32),
                     jax.ShapeDtypeStruct((ws,), jnp.uint8))
        call = jax.ffi.ffi_call("sde4_loss_grad", out_types, vmap_method="sequential")
        loss, gpar, _ = call(params, y, u, jnp.asarray(dt), keys, model_desc=_bytes_attr(desc),
                             data_cost_desc=_bytes_attr(data_cost.desc), num_particles=np.int32(N),
                             rng_mode=np.int32(_rng_mode()), noise_mode=np.int32(noise_mode))
        return jnp.mean(loss), gpar

    @jax.custom_vjp
    def packed_loss(params, y, u, keys):
        return _call(params, y, u, keys)[0]

    def fwd(params, y, u, keys):
        v, gpar = _call(params, y, u, keys)
        return v, (gpar, y, u, keys)

    def bwd(res, ct):
        gpar, y, u, keys = res
        return ct * gpar, jnp.zeros_like(y), jnp.zeros_like(u), np.zeros(keys.shape, jax.dtypes.float0)
    packed_loss.defvjp(fwd, bwd)

    def loss_fn(nn_params, y, u, rng):
        keys = _key_data(jax.random.split(rng, y.shape[0]))  # nsde.py:1196
        return packed_loss(tp.pack(nn_params, jnp), jnp.asarray(y, jnp.float32), jnp.asarray(u, jnp.float32), keys)
    return model.init_params(0), loss_fn

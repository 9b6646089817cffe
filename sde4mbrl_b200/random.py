"""``jax.random`` key plumbing on the device (threefry2x32, bit-exact with JAX).

Keys are jax-layout ``uint32[2]`` values stored as int32 CUDA tensors (same bits).  Splitting
runs in ``libsde4b200.so`` (``sde4_rng_split``) so no host round trip is needed; the rollout
kernels derive their per-particle Brownian keys themselves (``nsde.py:1026`` ->
``sde_solver.py:276``).  Reference call sites: ``sde4mbrl/sde_solver.py:276,277,283``,
``sde4mbrl/nsde.py:1026,1187,1196,1581``, ``rotor_uav/sde_mpc_design.py:227``.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .engine import as_key, _dev, _ptr, _stream

_partitionable = False  # jax_threefry_partitionable (False was JAX's default when the reference was written)


def set_threefry_partitionable(flag):
    """Select the jax.random stream layout (``jax.config.jax_threefry_partitionable``)."""
    global _partitionable
    _partitionable = bool(flag)


def rng_mode():
    return L.RNG_PARTITIONABLE if _partitionable else L.RNG_ORIGINAL


def PRNGKey(seed):
    """``jax.random.PRNGKey(seed)`` -> int32 CUDA tensor [2] with bits [hi32(seed), lo32(seed)]."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return as_key(np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=np.uint32))


def split(key, num=2):
    """``jax.random.split(key, num)`` -> int32 CUDA tensor [num, 2]."""
    key = as_key(key).reshape(2)
    out = torch.empty((num, 2), dtype=torch.int32, device=key.device)
    L.check(L.load().sde4_rng_split(_ptr(key), int(num), rng_mode(), _ptr(out), _stream()))
    return out


def key_data(key):
    """uint32 numpy view of a key tensor (host copy; for printing / tests)."""
    return as_key(key).cpu().numpy().view(np.uint32)


def bits(key, shape):
    """``jax.random.bits(key, shape)`` uint32 (as int32 bits) -- RNG self-test entry."""
    kd = key_data(key).reshape(2)
    n = int(np.prod(shape)) if len(shape) else 1
    out = torch.empty(n, dtype=torch.int32, device=_dev())
    k = (C.c_uint32 * 2)(int(kd[0]), int(kd[1]))
    L.check(L.load().sde4_rng_bits(k, n, 0, n, rng_mode(), _ptr(out), _stream()))
    return out.reshape(shape)


def normal(key, shape):
    """``jax.random.normal(key, shape)`` float32."""
    kd = key_data(key).reshape(2)
    n = int(np.prod(shape)) if len(shape) else 1
    out = torch.empty(n, dtype=torch.float32, device=_dev())
    k = (C.c_uint32 * 2)(int(kd[0]), int(kd[1]))
    L.check(L.load().sde4_rng_normal(k, n, 0, n, rng_mode(), _ptr(out), _stream()))
    return out.reshape(shape)


def split_many(keys, num=2):
    """``jax.vmap(lambda k: jax.random.split(k, num))(keys)``: keys[..., 2] -> int32 CUDA tensor [..., num, 2]."""
    keys = as_key(keys)
    lead = tuple(keys.shape[:-1])
    flat = keys.reshape(-1, 2).contiguous()
    out = torch.empty((flat.shape[0], num, 2), dtype=torch.int32, device=flat.device)
    L.check(L.load().sde4_rng_split_many(_ptr(flat), flat.shape[0], int(num), rng_mode(), _ptr(out), _stream()))
    return out.reshape(lead + (num, 2))


def normal_many(keys, shape):
    """``jax.vmap(lambda k: jax.random.normal(k, shape))(keys)``: keys[..., 2] -> float32 [..., *shape]."""
    keys = as_key(keys)
    lead = tuple(keys.shape[:-1])
    flat = keys.reshape(-1, 2).contiguous()
    n = int(np.prod(shape)) if len(shape) else 1
    out = torch.empty((flat.shape[0], n), dtype=torch.float32, device=flat.device)
    L.check(L.load().sde4_rng_normal_many(_ptr(flat), flat.shape[0], n, rng_mode(), _ptr(out), _stream()))
    return out.reshape(lead + tuple(shape))


def choice_mask_many(keys, n, m):
    """0/1 indicator of ``jax.vmap(lambda k: jax.random.choice(k, n, shape=(m,), replace=False))(keys)``:
    keys[K, 2] -> float32 [n, K] (time-major, the layout ``sde4_cost_args.step_weight`` reads)."""
    keys = as_key(keys)
    flat = keys.reshape(-1, 2).contiguous()
    out = torch.empty((int(n), flat.shape[0]), dtype=torch.float32, device=flat.device)
    L.check(L.load().sde4_rng_choice_mask(_ptr(flat), flat.shape[0], int(n), int(m), rng_mode(), _ptr(out), _stream()))
    return out


def randint_many(keys, n, minval, maxval):
    """``jax.vmap(lambda k: jax.random.randint(k, (n,), minval, maxval))(keys)``: keys[..., 2] -> int32 CUDA tensor
    [..., n] (the per-step data index of ``u_sampling_strategy: random``, sde4mbrl/nsde.py:65, :1308)."""
    keys = as_key(keys)
    lead = tuple(keys.shape[:-1])
    flat = keys.reshape(-1, 2).contiguous()
    out = torch.empty((flat.shape[0], int(n)), dtype=torch.int32, device=flat.device)
    L.check(L.load().sde4_rng_randint_many(_ptr(flat), flat.shape[0], int(n), int(minval), int(maxval), rng_mode(), _ptr(out),
                                           _stream()))
    return out.reshape(lead + (int(n),))


def randint(key, shape, minval, maxval):
    """``jax.random.randint(key, shape, minval, maxval)`` int32."""
    n = int(np.prod(shape)) if len(shape) else 1
    return randint_many(as_key(key).reshape(1, 2), n, minval, maxval).reshape(shape)

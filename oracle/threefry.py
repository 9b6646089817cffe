"""NumPy restatement of ``jax.random`` (threefry2x32) -- test infrastructure only.

The reference draws all of its randomness through ``jax.random``
(call sites: ``sde4mbrl/sde_solver.py:276,277,283``, ``sde4mbrl/nsde.py:1026,
1187,1196,1581``, ``sde4mbrlExamples/rotor_uav/sde_mpc_design.py:227``).  jax is
an un-vendored, un-pinned third-party dependency (``setup.py:21``) that is not
installed here, so this file restates its *published* algorithm:

* Threefry-2x32, 20 rounds (Salmon et al., "Parallel random numbers: as easy as
  1, 2, 3", SC'11; Random123 ``threefry2x32_R20``).
* ``jax._src.prng``: ``threefry_seed``, ``threefry_split``,
  ``threefry_random_bits`` in both stream layouts
  (``jax_threefry_partitionable`` False = MODE_ORIGINAL, True = MODE_PARTITIONABLE;
  the default flipped in JAX 0.5.0).
* ``jax.random.uniform`` / ``jax.random.normal`` (float32) with XLA's
  single-precision ``erf_inv`` (Giles' polynomial).

Only the uint32 bit stream is required to be bit exact; normals are compared
within a few ulp.
"""
import numpy as np

MODE_ORIGINAL = 0       # jax_threefry_partitionable = False (JAX < 0.5.0 default)
MODE_PARTITIONABLE = 1  # jax_threefry_partitionable = True  (JAX >= 0.5.0 default)

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_M32 = np.uint64(0xFFFFFFFF)


def _rotl(x, r):
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))).astype(np.uint32)


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32-20 on uint32 arrays (broadcasting).  Returns (y0, y1)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=np.uint32)
        k1 = np.asarray(k1, dtype=np.uint32)
        x0 = np.asarray(x0, dtype=np.uint32).copy()
        x1 = np.asarray(x1, dtype=np.uint32).copy()
        ks = (k0, k1, (k0 ^ k1 ^ np.uint32(0x1BD11BDA)).astype(np.uint32))
        x0 = (x0 + ks[0]).astype(np.uint32)
        x1 = (x1 + ks[1]).astype(np.uint32)
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = (x0 + x1).astype(np.uint32)
                x1 = _rotl(x1, r)
                x1 = (x1 ^ x0).astype(np.uint32)
            x0 = (x0 + ks[(g + 1) % 3]).astype(np.uint32)
            x1 = (x1 + ks[(g + 2) % 3] + np.uint32(g + 1)).astype(np.uint32)
    return x0, x1


def PRNGKey(seed):
    """``jax.random.PRNGKey`` for the threefry impl: [hi32(seed), lo32(seed)]."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=np.uint32)


def _threefry_2x32_flat(key, count):
    """``jax._src.prng.threefry_2x32``: hash a flat count vector by halves."""
    count = np.asarray(count, dtype=np.uint32).ravel()
    odd = count.size % 2
    if odd:
        count = np.concatenate([count, np.zeros(1, np.uint32)])
    half = count.size // 2
    y0, y1 = threefry2x32(key[0], key[1], count[:half], count[half:])
    out = np.concatenate([y0, y1])
    return out[:-1] if odd else out


def split(key, num=2, mode=MODE_ORIGINAL):
    """``jax.random.split(key, num)`` -> uint32[num, 2]."""
    key = np.asarray(key, dtype=np.uint32)
    if mode == MODE_ORIGINAL:
        return _threefry_2x32_flat(key, np.arange(2 * num, dtype=np.uint32)).reshape(num, 2)
    idx = np.arange(num, dtype=np.uint64)
    y0, y1 = threefry2x32(key[0], key[1], (idx >> np.uint64(32)).astype(np.uint32),
                          (idx & _M32).astype(np.uint32))
    return np.stack([y0, y1], axis=1)


def random_bits(key, shape, mode=MODE_ORIGINAL):
    """32-bit ``jax._src.prng.threefry_random_bits`` -> uint32[shape]."""
    key = np.asarray(key, dtype=np.uint32)
    size = int(np.prod(shape)) if len(shape) else 1
    if mode == MODE_ORIGINAL:
        return _threefry_2x32_flat(key, np.arange(size, dtype=np.uint32)).reshape(shape)
    idx = np.arange(size, dtype=np.uint64)
    y0, y1 = threefry2x32(key[0], key[1], (idx >> np.uint64(32)).astype(np.uint32),
                          (idx & _M32).astype(np.uint32))
    return (y0 ^ y1).reshape(shape)


def bits_to_unit_float(bits):
    """uint32 -> float32 in [0, 1): mantissa trick of ``jax.random.uniform``."""
    fb = (np.asarray(bits, np.uint32) >> np.uint32(9)) | np.uint32(0x3F800000)
    return fb.view(np.float32) - np.float32(1.0)


def uniform(key, shape, minval=0.0, maxval=1.0, mode=MODE_ORIGINAL):
    u = bits_to_unit_float(random_bits(key, shape, mode))
    minval = np.float32(minval)
    maxval = np.float32(maxval)
    return np.maximum(minval, (u * (maxval - minval) + minval).astype(np.float32))


_ERFINV_SMALL = np.array([2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
                          0.00021858087, -0.00125372503, -0.00417768164, 0.246640727,
                          1.50140941], dtype=np.float32)
_ERFINV_LARGE = np.array([-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
                          0.00573950773, -0.0076224613, 0.00943887047, 1.00167406,
                          2.83297682], dtype=np.float32)


def erf_inv_f32(x):
    """XLA's float32 ``erf_inv`` (Giles 2012, single precision), in float32."""
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = (-np.log1p((-x * x).astype(np.float32))).astype(np.float32)
        small = w < np.float32(5.0)
        ws = (w - np.float32(2.5)).astype(np.float32)
        wl = (np.sqrt(np.maximum(w, np.float32(0))) - np.float32(3.0)).astype(np.float32)
        ww = np.where(small, ws, wl).astype(np.float32)
        p = np.where(small, _ERFINV_SMALL[0], _ERFINV_LARGE[0]).astype(np.float32)
        for i in range(1, 9):
            c = np.where(small, _ERFINV_SMALL[i], _ERFINV_LARGE[i]).astype(np.float32)
            p = (c + p * ww).astype(np.float32)
        out = (p * x).astype(np.float32)
    return np.where(np.abs(x) == 1, np.copysign(np.float32(np.inf), x), out).astype(np.float32)


_LO = np.nextafter(np.float32(-1.0), np.float32(0.0))


def normal(key, shape, mode=MODE_ORIGINAL):
    """``jax.random.normal(key, shape)`` float32: sqrt(2)*erf_inv(uniform(lo,1))."""
    u = uniform(key, shape, _LO, np.float32(1.0), mode)
    return (np.float32(np.sqrt(2)) * erf_inv_f32(u)).astype(np.float32)


# ---------------------------------------------------------------------------
# Key-derivation chains of the hot path
# ---------------------------------------------------------------------------
def choice_without_replacement(key, n, m, mode=MODE_ORIGINAL):
    """``jax.random.choice(key, n, shape=(m,), replace=False)`` (sde4mbrl/nsde.py:755): jax draws
    ``permutation(key, n)[:m]``; ``permutation`` -> ``_shuffle(key, arange(n))`` = ``ceil(3 ln n / ln(2^32-1))`` rounds
    of ``key, subkey = split(key); x = stable_sort_by(random_bits(subkey, 32, (n,)), x)`` (jax/_src/random.py, restated
    from the published algorithm -- jax is not installable here, parity unpinned)."""
    x = np.arange(n)
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
    for _ in range(rounds):
        key, sub = split(key, 2, mode)
        x = x[np.argsort(random_bits(sub, (n,), mode), kind="stable")]
    return x[:m]


def randint(key, shape, minval, maxval, mode=MODE_ORIGINAL):
    """``jax.random.randint(key, shape, minval, maxval)`` for int32 (sde4mbrl/nsde.py:65, :1308): ``k1, k2 = split(key)``,
    ``hi, lo = random_bits(k1), random_bits(k2)``, ``span = maxval - minval`` (1 if not positive),
    ``mult = ((2^16 % span)^2) % span`` (= 2^32 % span), ``offset = ((hi % span) * mult + lo % span) % span`` in wrapping
    uint32 arithmetic (jax/_src/random.py ``_randint``, restated from the published algorithm -- parity unpinned)."""
    ks = split(key, 2, mode)
    hi = random_bits(ks[0], shape, mode).astype(np.uint64)
    lo = random_bits(ks[1], shape, mode).astype(np.uint64)
    span = np.uint64(1 if maxval <= minval else int(maxval) - int(minval))
    m32 = np.uint64(0xFFFFFFFF)
    mult = np.uint64(65536) % span
    mult = ((mult * mult) & m32) % span
    off = ((((hi % span) * mult) & m32) + (lo % span)) & m32
    off = off % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32)


def rollout_particle_brownian_keys(rng, num_particles, mode=MODE_ORIGINAL):
    """Per-particle Brownian keys: ``split(rng, N)[p]`` (nsde.py:1026) followed
    by ``rng_brownian, _ = split(k_p)`` (sde_solver.py:276)."""
    kp = split(rng, num_particles, mode)
    return np.stack([split(kp[p], 2, mode)[0] for p in range(num_particles)])


def rollout_noise(rng, num_particles, horizon, n_x, mode=MODE_ORIGINAL):
    """dw[N, H, n_x] exactly as ``multi_sampling`` -> ``euler_maruyama`` draws it
    (nsde.py:1026 -> sde_solver.py:276,283)."""
    kb = rollout_particle_brownian_keys(rng, num_particles, mode)
    return np.stack([normal(kb[p], (horizon, n_x), mode) for p in range(num_particles)])

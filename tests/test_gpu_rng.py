"""K0: the in-kernel threefry2x32 stream is BIT-EXACT with jax.random (both layouts); normals
within a few ulp of XLA's float32 erf_inv restatement."""
import numpy as np
import pytest

from oracle import threefry as tf

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("total", [1, 2, 5, 1300, 4097])
def test_bits_bit_exact(cuda, mode, total):
    from sde4mbrl_b200 import random as R
    R.set_threefry_partitionable(bool(mode))
    try:
        for seed in (0, 1, 2050, (5 << 32) + 17):
            got = R.bits(R.PRNGKey(seed), (total,)).cpu().numpy().view(np.uint32)
            want = tf.random_bits(tf.PRNGKey(seed), (total,), mode)
            np.testing.assert_array_equal(got, want)
    finally:
        R.set_threefry_partitionable(False)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("num", [1, 2, 3, 200, 4096])
def test_split_bit_exact(cuda, mode, num):
    from sde4mbrl_b200 import random as R
    R.set_threefry_partitionable(bool(mode))
    try:
        got = R.key_data(R.split(R.PRNGKey(2050), num))
        np.testing.assert_array_equal(got, tf.split(tf.PRNGKey(2050), num, mode))
    finally:
        R.set_threefry_partitionable(False)


def test_jax_doc_values(cuda):
    from sde4mbrl_b200 import random as R
    k = R.PRNGKey(42)
    assert abs(float(R.normal(k, (1,))[0]) - (-0.18471177)) < 5e-7
    ks = R.key_data(R.split(k))
    assert ks.tolist() == [[2465931498, 3679230171], [255383827, 267815257]]
    assert abs(float(R.normal(R.split(k)[1], (1,))[0]) - 1.3694694) < 5e-7


@pytest.mark.parametrize("mode", [0, 1])
def test_normal_close(cuda, mode):
    from sde4mbrl_b200 import random as R
    R.set_threefry_partitionable(bool(mode))
    try:
        got = R.normal(R.PRNGKey(9), (100000,)).cpu().numpy()
        want = tf.normal(tf.PRNGKey(9), (100000,), mode)
        # float32 tolerance: 4 ulp relative, 1e-6 absolute near zero
        assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) < 1e-6
        assert abs(got.mean()) < 0.02 and abs(got.std() - 1.0) < 0.02
    finally:
        R.set_threefry_partitionable(False)


@pytest.mark.parametrize("mode", [0, 1])
def test_batched_split_and_normal(cuda, mode):
    """sde4_rng_split_many / sde4_rng_normal_many == vmap of jax.random.split / normal over a batch of keys."""
    from sde4mbrl_b200 import random as R
    R.set_threefry_partitionable(bool(mode))
    try:
        keys = tf.split(tf.PRNGKey(77), 9, mode)
        got = R.split_many(keys, 5).cpu().numpy().view(np.uint32)
        for k in range(9):
            np.testing.assert_array_equal(got[k], tf.split(keys[k], 5, mode))
        deep = R.split_many(keys.reshape(3, 3, 2), 2).cpu().numpy().view(np.uint32)
        assert deep.shape == (3, 3, 2, 2)
        np.testing.assert_array_equal(deep[1, 2], tf.split(keys[5], 2, mode))
        z = R.normal_many(keys, (7, 3)).cpu().numpy()
        for k in range(9):
            want = tf.normal(keys[k], (7, 3), mode)
            assert np.max(np.abs(z[k] - want) / np.maximum(np.abs(want), 1.0)) < 1e-6
    finally:
        R.set_threefry_partitionable(False)


@pytest.mark.parametrize("mode", [0, 1])
def test_choice_mask(cuda, mode):
    """sde4_rng_choice_mask == indicator of jax.random.choice(key, n, (m,), replace=False) for a batch of keys."""
    from sde4mbrl_b200 import random as R
    R.set_threefry_partitionable(bool(mode))
    try:
        keys = tf.split(tf.PRNGKey(5), 33, mode)
        for n, m in ((7, 3), (30, 29), (1, 1), (50, 0), (64, 64)):
            got = R.choice_mask_many(keys, n, m).cpu().numpy()
            assert got.shape == (n, 33) and (got.sum(0) == m).all()
            for k in range(33):
                want = np.zeros(n, np.float32)
                want[tf.choice_without_replacement(keys[k], n, m, mode)] = 1.0
                np.testing.assert_array_equal(got[:, k], want)
    finally:
        R.set_threefry_partitionable(False)


@pytest.mark.parametrize("mode", [0, 1])
def test_randint_many(cuda, mode):
    """sde4_rng_randint_many == jax.random.randint per key, bit for bit (odd / even counts, spans that do and do not
    divide 2^32, a span beyond 2^16 where the uint32 multiplier wraps, the degenerate maxval <= minval)."""
    from sde4mbrl_b200 import random as R
    R.set_threefry_partitionable(bool(mode))
    try:
        keys = tf.split(tf.PRNGKey(9), 17, mode)
        for n, lo, hi in ((1, 0, 2), (7, 0, 3), (30, -4, 9), (50, 0, 4), (13, 5, 100008), (6, 2, 2), (9, 0, 1 << 30)):
            got = R.randint_many(keys, n, lo, hi).cpu().numpy()
            assert got.shape == (17, n) and got.dtype == np.int32
            for k in range(17):
                np.testing.assert_array_equal(got[k], tf.randint(keys[k], (n,), lo, hi, mode))
        np.testing.assert_array_equal(R.randint(keys[3], (4, 5), 0, 7).cpu().numpy(), tf.randint(keys[3], (4, 5), 0, 7, mode))
    finally:
        R.set_threefry_partitionable(False)

"""Pins the oracle's RNG: Random123 known-answer vectors for threefry2x32 and values printed
in the public JAX documentation for split / normal (both stream layouts)."""
import numpy as np

from oracle import threefry as tf


def test_threefry_random123_kat():
    # Random123 kat_vectors, threefry2x32 20 rounds
    assert [hex(int(v)) for v in tf.threefry2x32(0, 0, 0, 0)] == ["0x6b200159", "0x99ba4efe"]
    assert [hex(int(v)) for v in tf.threefry2x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)] == ["0x1cb996fc", "0xbb002be7"]
    assert [hex(int(v)) for v in tf.threefry2x32(0x13198A2E, 0x03707344, 0x243F6A88, 0x85A308D3)] == ["0xc4923a9c", "0x483df7a0"]


def test_prngkey_layout():
    assert tf.PRNGKey(0).tolist() == [0, 0]
    assert tf.PRNGKey(42).tolist() == [0, 42]
    assert tf.PRNGKey((7 << 32) + 5).tolist() == [7, 5]


def test_split_matches_jax_docs_original_layout():
    # jax.random.split(jax.random.PRNGKey(0)) with jax_threefry_partitionable=False
    assert tf.split(tf.PRNGKey(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    # "Pseudo random numbers in JAX": key = PRNGKey(42); new_key, subkey = split(key)
    assert tf.split(tf.PRNGKey(42)).tolist() == [[2465931498, 3679230171], [255383827, 267815257]]


def test_split_partitionable_layout():
    assert tf.split(tf.PRNGKey(0), mode=tf.MODE_PARTITIONABLE).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]


def test_normal_matches_jax_docs():
    # random.normal(PRNGKey(42)) -> -0.18471177 ; after split: random.normal(subkey) -> 1.3694694
    assert abs(float(tf.normal(tf.PRNGKey(42), ())) - (-0.18471177)) < 2e-7
    sub = tf.split(tf.PRNGKey(42))[1]
    assert abs(float(tf.normal(sub, ())) - 1.3694694) < 2e-7
    # JAX quickstart: random.normal(PRNGKey(0), (3,))
    np.testing.assert_allclose(tf.normal(tf.PRNGKey(0), (3,)), [1.8160863, -0.48262316, 0.33988908], rtol=0, atol=3e-7)
    # partitionable layout (JAX >= 0.5 docs): random.normal(key(42)) -> -0.028304616
    assert abs(float(tf.normal(tf.PRNGKey(42), (), mode=tf.MODE_PARTITIONABLE)) - (-0.028304616)) < 2e-7


def test_odd_size_padding_and_halves():
    key = tf.PRNGKey(7)
    b5 = tf.random_bits(key, (5,))
    # element j < half uses out0 of block (j, j+half) with the pad slot reading counter 0
    y0, y1 = tf.threefry2x32(key[0], key[1], np.array([0, 1, 2], np.uint32), np.array([3, 4, 0], np.uint32))
    assert b5.tolist() == [int(y0[0]), int(y0[1]), int(y0[2]), int(y1[0]), int(y1[1])]


def test_erfinv_against_scipy():
    import scipy.special as sp
    x = np.linspace(-0.9999, 0.9999, 4001).astype(np.float32)
    ref = sp.erfinv(x.astype(np.float64))
    got = tf.erf_inv_f32(x).astype(np.float64)
    # float32 conditioning of 1 - x*x near |x| -> 1 limits the tails to a few 1e-6
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3)) < 5e-6
    core = np.abs(x) < 0.99
    assert np.max(np.abs(got - ref)[core] / np.maximum(np.abs(ref[core]), 1e-3)) < 1e-6


def test_rollout_noise_chain():
    rng = tf.PRNGKey(3)
    dw = tf.rollout_noise(rng, 4, 5, 3)
    kp = tf.split(rng, 4)
    kb = tf.split(kp[2], 2)[0]
    np.testing.assert_array_equal(dw[2], tf.normal(kb, (5, 3)))


def test_choice_without_replacement_is_a_stable_sort_of_the_bit_stream():
    """jax.random.choice(key, n, (m,), replace=False) = permutation(key, n)[:m]: one stable sort of arange(n) by
    random_bits(split(key)[1], (n,)) for n <= 1625."""
    for mode in (tf.MODE_ORIGINAL, tf.MODE_PARTITIONABLE):
        key = tf.PRNGKey(12)
        full = tf.choice_without_replacement(key, 40, 40, mode)
        assert sorted(full.tolist()) == list(range(40))
        bits = tf.random_bits(tf.split(key, 2, mode)[1], (40,), mode)
        assert (np.diff(bits[full].astype(np.int64)) >= 0).all()
        np.testing.assert_array_equal(tf.choice_without_replacement(key, 40, 7, mode), full[:7])
        assert tf.choice_without_replacement(key, 1, 1, mode).tolist() == [0]


def test_randint_restatement_properties():
    """``jax.random.randint`` restatement (oracle/threefry.py, used by ``u_sampling_strategy: random``): for a span
    that divides 2^32 the multiplier 2^32 % span is 0 and the draw is the low word's remainder; span <= 0 returns
    minval; values stay in range and both stream layouts differ."""
    from oracle import threefry as tf
    import numpy as np
    key = tf.PRNGKey(42)
    for mode in (tf.MODE_ORIGINAL, tf.MODE_PARTITIONABLE):
        lo = tf.random_bits(tf.split(key, 2, mode)[1], (37,), mode)
        for span in (1, 2, 8, 1 << 16):
            got = tf.randint(key, (37,), 3, 3 + span, mode)
            np.testing.assert_array_equal(got, 3 + (lo % np.uint32(span)).astype(np.int32))
        np.testing.assert_array_equal(tf.randint(key, (5,), 7, 7, mode), np.full(5, 7, np.int32))
        for span in (3, 5, 7, 1000, 100003):
            hi = tf.random_bits(tf.split(key, 2, mode)[0], (37,), mode).astype(object)
            want = np.array([(int(h) * (1 << 32) + int(l)) % span for h, l in zip(hi, lo.astype(object))])
            got = tf.randint(key, (37,), -2, -2 + span, mode)
            assert got.min() >= -2 and got.max() < -2 + span
            # exact whenever (hi % span) * (2^32 % span) does not wrap uint32 -- true for these spans (< 2^16)
            if span < (1 << 16):
                np.testing.assert_array_equal(got, -2 + want)
    assert not np.array_equal(tf.randint(key, (64,), 0, 5, tf.MODE_ORIGINAL), tf.randint(key, (64,), 0, 5, tf.MODE_PARTITIONABLE))

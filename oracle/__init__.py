"""CPU oracle for the sde4mbrl hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's
data-parallel hot path (batched Euler-Maruyama particle rollouts of the
controlled neural SDE, the MPC cost, its gradient and the APG solver).  It is
the checker that the CUDA path in ``sde4mbrl_b200`` is compared against.

Rules (enforced by tests/test_layout.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import this package;
  * nothing under ``sde4mbrl_b200/`` imports it -- the product path fails
    loudly when the CUDA library is missing, it never falls back to this code.

PARITY STATUS: **parity unpinned against the reference itself.**  The
reference (pure Python/JAX) cannot be imported here (no jax / haiku in the
image, no network) and its own ``tests/`` hold no golden vector, known-answer
test or fixture for this path (``tests/opt_test.py`` is an unrunnable script
without assertions).  What *is* pinned:
  * threefry2x32 against the published Random123 known-answer vectors, and the
    jax.random split / bits / normal layouts against values printed in the
    public JAX documentation (see ``tests/test_oracle_rng.py``);
  * the model math against the reference's shipped learned weights (pickles)
    by physical sanity checks (hover equilibrium etc.) and against
    finite differences / float64 autograd for the gradients.
Every function cites the reference ``file:line`` it follows.
"""

"""CPU oracle for the chiron particle hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the reference algorithm (choderalab/chiron,
`chiron/neighbors.py`, `potential.py`, `integrators.py`, `mcmc.py`, `states.py`, `utils.py`)
plus the third-party arithmetic it calls (`jax.random` legacy threefry, XLA fp32 `erf_inv`,
`jnp.mod`).  It exists so that `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` can check and time the reference algorithm on the CPU.

Nothing under `chiron_b200/` may import it: the product path is CUDA-only and fails loudly
when the extension is missing.

Parity pin status: PINNED against the reference's own golden vectors
(`chiron/tests/test_pairs.py`, `test_mcmc.py:81-84`, `test_mcmc.py:451-452`,
`test_utils.py:86-113`, `test_potential.py:59-93`), see `tests/test_oracle_golden.py` and
`tests/golden/`.  The reference itself cannot be imported here (jax/openmm are not installed).
"""

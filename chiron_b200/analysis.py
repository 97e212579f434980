"""`MBAREstimator` with the surface of `chiron/analysis.py`, without pymbar (absent in this image).

The reference hands `pymbar.MBAR` the reported energies in the legacy `u_kln` layout
(`analysis.py:27-30`: `(n_iterations, n_states, n_replicas)` transposed to
`(n_replicas, n_states, n_iterations)`, i.e. sample n of the replica that started in state k,
evaluated at state l) together with `N_k`.  This class solves the same MBAR self-consistent
equations (Shirts & Chodera 2008, eq. 11) by fixed-point iteration followed by Newton steps, in
NumPy on the host: post-hoc analysis, not part of the device hot path.
"""
import numpy as np


def _logsumexp(a, axis):
    m = np.max(a, axis=axis, keepdims=True)
    return (m + np.log(np.sum(np.exp(a - m), axis=axis, keepdims=True))).squeeze(axis)


def solve_mbar(u_ln: np.ndarray, N_l: np.ndarray, tol: float = 1e-10, max_iter: int = 10000) -> np.ndarray:
    """Free energies f_l (f_0 = 0) from reduced potentials u_ln[l, n] of all pooled samples n
    evaluated at every state l, with N_l samples drawn from state l."""
    u_ln = np.asarray(u_ln, dtype=np.float64)
    N_l = np.asarray(N_l, dtype=np.float64)
    K = u_ln.shape[0]
    f = np.zeros(K)
    sampled = N_l > 0
    logN = np.where(sampled, np.log(np.where(sampled, N_l, 1.0)), -np.inf)
    for _ in range(max_iter):
        # log of the mixture denominator per sample: log sum_k N_k exp(f_k - u_kn)
        log_den = _logsumexp((logN + f)[:, None] - u_ln, axis=0)
        f_new = -_logsumexp(-u_ln - log_den[None, :], axis=1)
        f_new -= f_new[0]
        if np.max(np.abs(f_new - f)) < tol:
            f = f_new
            break
        f = f_new
    return f


class MBAREstimator:
    def __init__(self) -> None:
        self.mbar_f_k = None
        self.mbar = None

    def initialize(self, u_kn: np.ndarray, N_k: np.ndarray):
        """u_kn: (n_iterations, n_states, n_replicas) as reported by MultiStateSampler; N_k samples per
        state (the reference passes `[iteration] * n_states`, `multistate.py:690`)."""
        u = np.transpose(np.asarray(u_kn, dtype=np.float64), (2, 1, 0))   # (replica k, state l, iteration n)
        N_k = np.asarray(N_k, dtype=int)
        K_from, K_at, _ = u.shape
        cols = [u[k, :, :N_k[k]] for k in range(K_from)]                  # pymbar's u_kln convention
        u_ln = np.concatenate(cols, axis=1)
        self.mbar_f_k = solve_mbar(u_ln, N_k[:K_at] if K_from == K_at else np.bincount(np.arange(K_from), N_k, K_at))
        self.mbar = self

    @property
    def f_k(self):
        return self.mbar_f_k

    def get_free_energy_difference(self):
        return self.mbar_f_k[-1]

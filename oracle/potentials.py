"""NumPy restatement of chiron's potentials (energy and the force jax.grad produces).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

    LJ masked energy        chiron/potential.py:193-213, 276-279
    LJ no-list N^2 path     chiron/potential.py:26-63, 235-258
    force = -grad(energy)   chiron/potential.py:21-24  (analytic equivalent: potential.py:322-326)
    harmonic oscillator     chiron/potential.py:413-418
"""
import numpy as np
from . import pairs


def lj_pair_energy(d, sigma, epsilon, dtype=np.float32):
    """4 eps ((s/d)^12 - (s/d)^6) with integer powers, potential.py:208-212."""
    with np.errstate(divide="ignore", invalid="ignore"):
        q = (dtype(sigma) / d).astype(dtype)
        q2 = (q * q).astype(dtype)
        q6 = (q2 * q2 * q2).astype(dtype)
        q12 = (q6 * q6).astype(dtype)
        return ((dtype(4.0) * dtype(epsilon)) * (q12 - q6)).astype(dtype)


def lj_pair_force_scalar(d, sigma, epsilon, dtype=np.float32):
    """f such that F_i += f * r_ij: 24 eps / d^2 (2 (s/d)^12 - (s/d)^6), potential.py:322-326."""
    with np.errstate(divide="ignore", invalid="ignore"):
        q = (dtype(sigma) / d).astype(dtype)
        q2 = (q * q).astype(dtype)
        q6 = (q2 * q2 * q2).astype(dtype)
        q12 = (q6 * q6).astype(dtype)
        return (dtype(24.0) * (dtype(epsilon) / (d * d)) * (dtype(2.0) * q12 - q6)).astype(dtype)


def lj_energy_nlist(x, box, sigma, epsilon, cutoff, neighbor_list, neighbor_mask, periodic=True,
                    dtype=np.float32, accumulate=np.float64):
    """LJPotential.compute_energy with a NeighborListNsqrd (potential.py:272-279).
    Masked-out entries are skipped (the reference multiplies by 0; identical unless d == 0)."""
    _, _, mask, d, _ = pairs.calculate_neighborlist(x, box, cutoff, neighbor_list, neighbor_mask,
                                                    periodic, dtype)
    e = lj_pair_energy(d, sigma, epsilon, dtype)
    return np.where(mask != 0, e, 0).sum(dtype=accumulate)


def lj_force_nlist(x, box, sigma, epsilon, cutoff, neighbor_list, neighbor_mask, periodic=True,
                   dtype=np.float32, accumulate=np.float64):
    """-grad of lj_energy_nlist: scatter-add of f*r_ij to i and -f*r_ij to j."""
    _, nl, mask, d, r = pairs.calculate_neighborlist(x, box, cutoff, neighbor_list, neighbor_mask,
                                                     periodic, dtype)
    f = np.where(mask != 0, lj_pair_force_scalar(d, sigma, epsilon, dtype), 0).astype(accumulate)
    fv = f[..., None] * r.astype(accumulate)
    F = fv.sum(axis=1)
    np.subtract.at(F, np.asarray(nl).astype(np.int64).reshape(-1), fv.reshape(-1, 3))
    return F.astype(dtype)


def lj_energy_force_bruteforce(x, box, sigma, epsilon, cutoff, periodic=True, dtype=np.float32,
                               chunk=1024, want_force=True):
    """All pairs i<j with d < cutoff, chunked (the pair set a valid Verlet list reduces to).
    Returns (energy f64, force (N,3) dtype or None, n_interacting_pairs)."""
    x = np.asarray(x, dtype=dtype)
    n = x.shape[0]
    F = np.zeros((n, 3), dtype=np.float64)
    e_tot, n_int = 0.0, 0
    for s in range(0, n, chunk):
        ii = np.arange(s, min(n, s + chunk))
        r, d = pairs.displacement(x[ii][:, None, :], x[None, :, :], box, periodic, dtype)
        m = (d < dtype(cutoff)) & (ii[:, None] != np.arange(n)[None, :])
        e = np.where(m, lj_pair_energy(d, sigma, epsilon, dtype), 0)
        e_tot += 0.5 * e.sum(dtype=np.float64)
        n_int += int(m.sum())
        if want_force:
            f = np.where(m, lj_pair_force_scalar(d, sigma, epsilon, dtype), 0).astype(np.float64)
            F[ii] += (f[..., None] * r.astype(np.float64)).sum(axis=1)
    return e_tot, (F.astype(dtype) if want_force else None), n_int // 2


def lj_energy_nopbc(x, sigma, epsilon, cutoff, dtype=np.float32):
    """LJPotential.compute_energy(positions, nbr_list=None): non-periodic N^2 pair list
    (potential.py:26-63, 235-258)."""
    e, _, _ = lj_energy_force_bruteforce(x, None, sigma, epsilon, cutoff, periodic=False,
                                         dtype=dtype, want_force=False)
    return e


def ho_energy(x, x0, k, U0, dtype=np.float32):
    """0.5 k sum (x-x0)^2 + U0 (potential.py:415-417)."""
    dx = (np.asarray(x, dtype=dtype) - np.asarray(x0, dtype=dtype)).astype(dtype)
    return dtype(dtype(0.5) * dtype(k) * np.sum((dx * dx).astype(dtype), dtype=dtype) + dtype(U0))


def ho_force(x, x0, k, dtype=np.float32):
    dx = (np.asarray(x, dtype=dtype) - np.asarray(x0, dtype=dtype)).astype(dtype)
    return (-(dtype(k)) * dx).astype(dtype)


def lj_mixture_energy_force_nlist(x, box, sigma_i, epsilon_i, cutoff, neighbor_list, neighbor_mask, shift=False,
                                  periodic=True, dtype=np.float64, switch_distance=0.0):
    """Generalisation of `lj_energy_nlist` / `lj_force_nlist` to per-particle parameters with Lorentz-Berthelot
    mixing (sigma_ij = (sigma_i + sigma_j)/2, eps_ij = sqrt(eps_i eps_j)), an optional energy shift at the
    cutoff and an optional quintic switching function.  The reference has no such path (potential.py:131-137 takes one sigma / epsilon);
    the pair formula is potential.py:208-212.  Evaluated in float64: the checker of chiron_b200.LJMixturePotential."""
    _, nl, mask, d, r = pairs.calculate_neighborlist(x, box, cutoff, neighbor_list, neighbor_mask, periodic, dtype)
    nl = np.asarray(nl).astype(np.int64)
    sig = np.asarray(sigma_i, dtype=dtype)
    eps = np.asarray(epsilon_i, dtype=dtype)
    sij = 0.5 * (sig[:, None] + sig[nl])
    eij = np.sqrt(eps[:, None] * eps[nl])
    m = mask != 0
    with np.errstate(divide="ignore", invalid="ignore"):
        q6 = (sij / d) ** 6
        e = 4.0 * eij * (q6 * q6 - q6)
        if shift:
            qc6 = (sij / dtype(cutoff)) ** 6
            e = e - 4.0 * eij * (qc6 * qc6 - qc6)
        f = 24.0 * (eij / (d * d)) * (2.0 * q6 * q6 - q6)
        if switch_distance > 0.0:
            # OpenMM NonbondedForce switching function between switch_distance and the cutoff
            w = 1.0 / (dtype(cutoff) - dtype(switch_distance))
            t = np.clip((d - dtype(switch_distance)) * w, 0.0, 1.0)
            S = 1.0 + t ** 3 * (-10.0 + t * (15.0 - 6.0 * t))
            dS = -30.0 * t * t * (1.0 - t) ** 2 * w
            f = f * S - e * dS / d
            e = e * S
    energy = np.where(m, e, 0.0).sum(dtype=np.float64)
    fv = np.where(m, f, 0.0)[..., None] * r
    F = fv.sum(axis=1)
    np.subtract.at(F, nl.reshape(-1), fv.reshape(-1, 3))
    return float(energy), F

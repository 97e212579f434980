"""NumPy restatement of chiron's BAOAB Langevin integrator and Metropolis moves.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

    LangevinIntegrator.run      chiron/integrators.py:110-218
    initialize_velocities       chiron/utils.py:116-144
    SamplerState.new_PRNG_key   chiron/states.py:150-154
    reduced potential           chiron/states.py:302-325
    MCMove._step / accept       chiron/mcmc.py:357-463, 531-548
    displacement move           chiron/mcmc.py:680-787
    barostat move               chiron/mcmc.py:913-1009
"""
import numpy as np
from . import jax_random as jr
from . import pairs

KB_J_PER_K = 1.380649e-23            # openmm.unit.BOLTZMANN_CONSTANT_kB
NA_PER_MOL = 6.02214076e23           # openmm.unit.AVOGADRO_CONSTANT_NA
R_KJ_PER_MOL_K = KB_J_PER_K * NA_PER_MOL * 1e-3
ATM_NM3_IN_KJ = 101325.0 * 1e-27 * 1e-3

f32 = np.float32


class KeyedState:
    """The PRNG part of chiron's SamplerState (states.py:150-154)."""

    def __init__(self, key):
        self.key = np.asarray(key, dtype=np.uint32)

    def new_key(self):
        k = jr.split(self.key)
        self.key = k[0]
        return k[1]


def prng_stream(seed):
    """chiron.utils.PRNG: set_seed + get_random_key (utils.py:29-38) as a generator."""
    key = jr.PRNGKey(seed)
    while True:
        k = jr.split(key)
        key = k[0]
        yield k[1]


def maxwell_boltzmann(key, masses, temperature_K):
    """utils.initialize_velocities: sqrt(kT/m) * normal(key, (N,3)) in fp32."""
    kT = R_KJ_PER_MOL_K * float(temperature_K)
    m = np.asarray(masses, dtype=f32)[:, None]
    sigma_v = np.sqrt(f32(kT) / m).astype(f32)
    return (sigma_v * jr.normal(key, (m.shape[0], 3))).astype(f32)


def langevin_run(x, v, masses, temperature_K, dt_ps, gamma_per_ps, state: KeyedState, nsteps,
                 force_fn, energy_fn=None, report_interval=100, refresh_velocities=False,
                 nbr=None, trace=None):
    """LangevinIntegrator.run.  `nbr` (optional) is an object with wrap(x), check(x), build(x);
    returns (x, v, key_after_loop, reported_energies)."""
    key = state.new_key()
    kT = R_KJ_PER_MOL_K * float(temperature_K)
    m = np.asarray(masses, dtype=f32)[:, None]
    sigma_v = np.sqrt(f32(kT) / m).astype(f32)
    # exp evaluated in fp64 on the fp32-rounded argument, then rounded once: a correctly rounded
    # fp32 exp (NumPy's SIMD fp32 exp is 1 ulp off at -0.004, which moves b by 7e-6).
    a = f32(np.exp(np.float64(f32(-gamma_per_ps * dt_ps))))
    b = np.sqrt(f32(1.0) - f32(np.exp(np.float64(f32(-2.0 * gamma_per_ps * dt_ps))))).astype(f32)
    x = np.array(x, dtype=f32)
    if refresh_velocities or v is None or np.shape(v)[0] != x.shape[0]:
        v = (sigma_v * jr.normal(key, x.shape)).astype(f32)
    v = np.array(v, dtype=f32)
    hdt = f32(dt_ps * 0.5)
    if nbr is not None:
        nbr.build(x)
    F = np.asarray(force_fn(x), dtype=f32)
    energies = []
    for step in range(int(nsteps)):
        k = jr.split(key)
        key, sub = k[0], k[1]
        v = (v + ((hdt * F).astype(f32) / m).astype(f32)).astype(f32)
        x = (x + (hdt * v).astype(f32)).astype(f32)
        xi = jr.normal(sub, x.shape)
        v = ((a * v).astype(f32) + (((b * sigma_v).astype(f32)) * xi).astype(f32)).astype(f32)
        x = (x + (hdt * v).astype(f32)).astype(f32)
        if nbr is not None:
            x = nbr.wrap(x)
            if nbr.check(x):
                nbr.build(x)
        F = np.asarray(force_fn(x), dtype=f32)
        v = (v + ((hdt * F).astype(f32) / m).astype(f32)).astype(f32)
        if step % report_interval == 0 and energy_fn is not None:
            energies.append(float(energy_fn(x)))
        if trace is not None:
            trace.append((x.copy(), v.copy()))
    return x, v, key, energies


class OracleNeighborList:
    """Minimal NeighborListNsqrd (neighbors.py:446-907) for the oracle integrator / moves."""

    def __init__(self, box, cutoff, skin, n_max_neighbors=200, periodic=True):
        self.box = None if box is None else np.asarray(box, dtype=f32)
        self.cutoff, self.skin, self.M, self.periodic = cutoff, skin, n_max_neighbors, periodic
        self.n_builds = 0

    def wrap(self, x):
        return pairs.wrap(x, self.box, self.periodic)

    def build(self, x, box=None):
        if box is not None:
            self.box = np.asarray(box, dtype=f32)
        self.ref = np.array(x, dtype=f32)
        out = pairs.build_neighborlist(x, self.box, self.cutoff, self.skin, self.M, self.periodic)
        self.neighbor_list, self.neighbor_mask = out["neighbor_list"], out["neighbor_mask"]
        self.n_neighbors, self.M = out["n_neighbors"], out["n_max_neighbors"]
        self.n_builds += 1

    def check(self, x):
        return pairs.check_neighborlist(x, self.ref, self.box, self.skin, self.periodic)


def reduced_potential(U_kj_mol, temperature_K, pressure_atm=None, volume_nm3=None):
    """ThermodynamicState.get_reduced_potential (states.py:302-325) following the reference's
    fp32 op chain: beta * (U / N_A + P V)."""
    beta_per_J = 1.0 / (KB_J_PER_K * float(temperature_K))
    rp = f32(f32(U_kj_mol) / f32(NA_PER_MOL))                       # kJ per particle
    if pressure_atm is not None:
        pv = f32(f32(pressure_atm) * f32(volume_nm3))                # atm nm^3
        rp = f32(rp + f32(pv * f32(ATM_NM3_IN_KJ)))
    return f32(f32(f32(beta_per_J) * rp) * f32(1000.0))


def metropolis_accept(log_ratio, key):
    """MCMove._accept_or_reject (mcmc.py:541-548)."""
    u = jr.uniform(key)
    return bool((-log_ratio <= 0.0) or (u < np.exp(f32(log_ratio))))


def mc_displacement_step(x, box, state: KeyedState, sigma_disp, u_current, reduced_fn, nbr=None,
                         subset_mask=None):
    """One MonteCarloDisplacementMove step (mcmc.py:733-787 + 396-463).
    reduced_fn(x, box) -> reduced potential.  Returns (x_new, u_new, accepted)."""
    key = state.new_key()
    disp = (jr.normal(key, np.shape(x)) * f32(sigma_disp)).astype(f32)
    if subset_mask is not None:
        disp = (disp * np.asarray(subset_mask, dtype=f32)[:, None]).astype(f32)
    xp = (np.asarray(x, dtype=f32) + disp).astype(f32)
    rebuilt = False
    if nbr is not None:
        xp = nbr.wrap(xp)
        if nbr.check(xp):
            saved = (nbr.ref, nbr.neighbor_list, nbr.neighbor_mask, nbr.n_neighbors, nbr.M)
            nbr.build(xp)
            rebuilt = True
    u_new = reduced_fn(xp, box)
    log_ratio = f32(-u_new + u_current)
    if np.isnan(u_new):
        decision = False
    else:
        decision = metropolis_accept(log_ratio, state.new_key())
    if decision:
        return xp, u_new, True
    if rebuilt:
        nbr.ref, nbr.neighbor_list, nbr.neighbor_mask, nbr.n_neighbors, nbr.M = saved
    return x, u_current, False


def mc_barostat_step(x, box, state: KeyedState, volume_max_scale, u_current, reduced_fn, nbr=None):
    """One MonteCarloBarostatMove step (mcmc.py:956-1009 + 396-463).
    Returns (x_new, box_new, u_new, accepted)."""
    key = state.new_key()
    box = np.asarray(box, dtype=f32)
    n = np.shape(x)[0]
    V0 = f32(f32(box[0, 0] * box[1, 1]) * box[2, 2])
    dV_max = f32(f32(volume_max_scale) * V0)
    dV = f32(jr.uniform(key, (), -1.0, 1.0) * dV_max)
    V1 = f32(V0 + dV)
    s = np.power(f32(V1 / V0), f32(1.0 / 3.0)).astype(f32)
    xp = (np.asarray(x, dtype=f32) * s).astype(f32)
    boxp = (box * s).astype(f32)
    saved = None
    if nbr is not None:
        saved = (nbr.ref, nbr.neighbor_list, nbr.neighbor_mask, nbr.n_neighbors, nbr.M, nbr.box)
        nbr.build(xp, boxp)
    u_new = reduced_fn(xp, boxp)
    log_ratio = f32(f32(-(f32(u_new - u_current))) + f32(f32(n) * np.log(f32(V1 / V0))))
    if np.isnan(u_new):
        decision = False
    else:
        decision = metropolis_accept(log_ratio, state.new_key())
    if decision:
        return xp, boxp, u_new, True
    if saved is not None:
        nbr.ref, nbr.neighbor_list, nbr.neighbor_mask, nbr.n_neighbors, nbr.M, nbr.box = saved
    return x, box, u_current, False

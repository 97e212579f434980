"""x64 variants (`chx_*_f64`, chiron_b200.x64) against the oracle evaluated in float64 -- the reference with
`jax_enable_x64`.  Bar (BASELINE.json north_star): displacements, lists, masks bit-exact; energies and forces within
rel 1e-10; a BAOAB step with the noise handed in to 1e-14."""
import numpy as np
import pytest
import torch

from oracle import pairs, potentials as pot

pytestmark = pytest.mark.gpu
f64 = np.float64


def _np(t):
    return t.detach().cpu().numpy()


def _lj(n_side, rho_star, seed):
    from chiron_b200 import unit
    from chiron_b200.testsystems import LennardJonesFluid
    lj = LennardJonesFluid(nparticles=n_side ** 3, reduced_density=rho_star, seed=seed)
    x = np.asarray(lj.positions.value_in_unit(unit.nanometer), dtype=f64)
    box = np.asarray(lj.box_vectors.value_in_unit(unit.nanometer), dtype=f64)
    rng = np.random.default_rng(seed)
    x = x + rng.normal(size=x.shape) * 0.013            # genuine float64 coordinates, not fp32-representable
    return x, box


def test_displacement_and_wrap_f64_bit_exact(cuda_device):
    from chiron_b200 import x64
    rng = np.random.default_rng(3)
    box = np.diag([3.1, 4.7, 2.9])
    a, b = rng.normal(size=(5000, 3)) * 6, rng.normal(size=(5000, 3)) * 6
    r, d = x64.displacement(a, b, box)
    ro, do = pairs.displacement(a, b, box, dtype=f64)
    assert _np(r).dtype == f64 and np.array_equal(_np(r), ro) and np.array_equal(_np(d), do)
    r, d = x64.displacement(a, b, None, periodic=False)
    ro, do = pairs.displacement(a, b, None, periodic=False, dtype=f64)
    assert np.array_equal(_np(r), ro) and np.array_equal(_np(d), do)
    assert np.array_equal(_np(x64.wrap(a, box)), pairs.wrap(a, box, dtype=f64))
    # the float64 path is not the fp32 path in disguise
    r32, _ = pairs.displacement(a.astype(np.float32), b.astype(np.float32), box.astype(np.float32))
    assert np.abs(ro - r32).max() > 1e-9


@pytest.mark.parametrize("n_side,rho", [(8, 0.8), (10, 0.1)])
def test_neighborlist_f64_build_calculate_check(cuda_device, n_side, rho):
    from chiron_b200 import x64
    x, box = _lj(n_side, rho, seed=21)
    x[::7] += np.diag(box) * 1.5                          # a few particles outside the box (the reference does not wrap)
    rc, skin = 1.02, 0.4
    nl = x64.build_neighborlist(x, box, rc, skin, 30)     # 30: forces the growth loop at rho* = 0.8
    ref = pairs.build_neighborlist(x, box, rc, skin, nl["n_max_neighbors"], dtype=f64)
    assert np.array_equal(_np(nl["n_neighbors"]), ref["n_neighbors"])
    assert np.array_equal(_np(nl["neighbor_list"]).astype(np.uint32), ref["neighbor_list"])
    assert np.array_equal(_np(nl["neighbor_mask"]), ref["neighbor_mask"].astype(np.int32))
    xm = x + np.random.default_rng(5).normal(size=x.shape) * 0.05
    n, _, mask, d, r = x64.calculate(xm, box, rc, nl["neighbor_list"], nl["neighbor_mask"])
    no, _, mo, do, ro = pairs.calculate_neighborlist(xm, box, rc, ref["neighbor_list"], ref["neighbor_mask"], dtype=f64)
    assert np.array_equal(_np(n), no) and np.array_equal(_np(mask), mo)
    assert np.array_equal(_np(d), do) and np.array_equal(_np(r), ro)
    assert x64.check(xm, x, box, 0.02) is pairs.check_neighborlist(xm, x, box, 0.02, dtype=f64) is True
    assert x64.check(x + 1e-4, x, box, skin) is pairs.check_neighborlist(x + 1e-4, x, box, skin, dtype=f64) is False


def test_lj_energy_force_f64_rel_1e10(cuda_device):
    from chiron_b200 import x64
    sigma, eps, rc, skin = 0.34, 0.238 * 4.184, 1.02, 0.4
    x, box = _lj(9, 0.8, seed=33)
    nl = x64.build_neighborlist(x, box, rc, skin, 200)
    ref = pairs.build_neighborlist(x, box, rc, skin, nl["n_max_neighbors"], dtype=f64)
    e, F = x64.lj_energy_force(x, box, sigma, eps, rc, nl["neighbor_list"], nl["neighbor_mask"])
    e_ref = pot.lj_energy_nlist(x, box, sigma, eps, rc, ref["neighbor_list"], ref["neighbor_mask"], dtype=f64)
    F_ref = pot.lj_force_nlist(x, box, sigma, eps, rc, ref["neighbor_list"], ref["neighbor_mask"], dtype=f64)
    assert abs(float(e) - float(e_ref)) <= 1e-10 * abs(float(e_ref))
    assert np.abs(_np(F) - F_ref).max() <= 1e-10 * np.abs(F_ref).max()
    # and the fp32 path of the same configuration agrees with it to fp32 accuracy only
    from chiron_b200 import unit
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.testsystems import _topology
    p32 = LJPotential(_topology(x.shape[0]), sigma * unit.nanometer, 0.238 * unit.kilocalories_per_mole, rc * unit.nanometer)
    nl32 = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=rc * unit.nanometer, skin=skin * unit.nanometer, n_max_neighbors=200)
    nl32.build(x.astype(np.float32), box.astype(np.float32))
    e32 = float(p32.compute_energy(x.astype(np.float32), nl32))
    assert 1e-9 < abs(e32 - float(e_ref)) / abs(float(e_ref)) < 1e-5


def test_baoab_step_f64_with_injected_noise(cuda_device):
    from chiron_b200 import x64
    rng = np.random.default_rng(9)
    n = 777
    box = np.diag([3.0, 3.5, 4.0])
    x, v, F = rng.random((n, 3)) * 3, rng.normal(size=(n, 3)) * 0.3, rng.normal(size=(n, 3)) * 500
    mass, xi = rng.uniform(10, 50, n), rng.normal(size=(n, 3))
    h, kT = 0.0005, 2.494
    a, b = np.exp(-0.001), np.sqrt(1 - np.exp(-0.002))
    dev = cuda_device
    xd, vd = torch.as_tensor(x, device=dev).clone(), torch.as_tensor(v, device=dev).clone()
    Fd, md, nd = torch.as_tensor(F, device=dev), torch.as_tensor(mass, device=dev), torch.as_tensor(xi, device=dev)
    x64.baoab_update(xd, vd, Fd, md, nd, h, a, b, kT, box)
    m = mass[:, None]
    vv = v + (h * F) / m
    xx = x + h * vv
    vv = a * vv + (b * np.sqrt(kT / m)) * xi
    xx = xx + h * vv
    L = np.diag(box)
    xx = xx - np.floor(xx / L) * L
    assert np.abs(_np(xd) - xx).max() < 1e-14 * 4.0 and np.abs(_np(vd) - vv).max() < 1e-14 * np.abs(vv).max()
    x64.kick(vd, Fd, md, h)
    assert np.abs(_np(vd) - (vv + (h * F) / m)).max() < 1e-14 * np.abs(vv).max()

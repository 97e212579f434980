"""`minimize_energy` with the interface of `chiron/minimze.py` (the module name is the reference's).

The reference wraps `jaxopt.GradientDescent` around `jax.value_and_grad(potential_fn)`
(`minimze.py:5-45`); jaxopt and JAX are not available here, and the potentials of this package
provide analytic forces from the same kernel that evaluates the energy, so the minimiser is a
steepest-descent loop with a backtracking / growing step on `compute_energy_and_force`.  Like
jaxopt's `OptStep`, the result exposes the optimised coordinates as `.params`.
"""
from typing import Callable, NamedTuple

import torch

from . import _lib


class OptStep(NamedTuple):
    params: torch.Tensor
    state: dict


def _energy_force_fn(potential_fn: Callable, nbr_list):
    owner = getattr(potential_fn, "__self__", None)
    if owner is None or not hasattr(owner, "compute_force"):
        raise TypeError("minimize_energy needs a bound `compute_energy` of a potential that provides "
                        "`compute_force` (no autodiff in chiron_b200)")

    def fn(x):
        if hasattr(owner, "compute_energy_and_force"):
            try:
                e, f = owner.compute_energy_and_force(x, nbr_list)
            except TypeError:
                e, f = owner.compute_energy_and_force(x)
        else:
            e, f = owner.compute_energy(x, nbr_list), owner.compute_force(x, nbr_list)
        if not isinstance(f, torch.Tensor):
            f = torch.zeros_like(x)
        return float(e), f
    return fn


def minimize_energy(coordinates, potential_fn: Callable, nbr_list=None, maxiter: int = 1000,
                    tolerance: float = 1e-4, initial_step: float = 1e-4) -> OptStep:
    """Minimise `potential_fn(coordinates, nbr_list)`; stops after `maxiter` force evaluations or when
    the largest force component drops below `tolerance` (kJ/mol/nm)."""
    x = _lib.as_device_f32(getattr(coordinates, "_value", coordinates)).clone()
    fn = _energy_force_fn(potential_fn, nbr_list)
    if maxiter <= 0:
        return OptStep(params=x, state={"iter_num": 0, "error": float("nan")})
    e, f = fn(x)
    step, it = float(initial_step), 0
    fmax = float(f.abs().max())
    while it < maxiter and fmax >= tolerance and step > 1e-14:
        x_try = x + step * f
        e_try, f_try = fn(x_try)
        it += 1
        if e_try <= e:
            x, e, f = x_try, e_try, f_try
            fmax = float(f.abs().max())
            step *= 1.5
        else:
            step *= 0.25
    return OptStep(params=x, state={"iter_num": it, "error": fmax, "value": e})

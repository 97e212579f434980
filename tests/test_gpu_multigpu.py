"""Replica exchange sharded over 2 GPUs (NCCL) against the same run in one process: identical state indices and
swap statistics, energy matrices equal to rounding (chiron/multistate.py:497-531,
563-599 is the loop that is sharded).  Needs 2 visible GPUs: `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_multigpu.py -m gpu`; skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_REP, SWEEPS, STEPS = 8, 4, 30


def _run(rank, world, port, out_dir):
    import torch.distributed as dist
    from loguru import logger
    logger.remove()
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        torch.cuda.set_device(0)
    from chiron_b200 import random as crandom, unit
    from chiron_b200.mcmc import LangevinDynamicsMove, MCMCSampler, MoveSchedule
    from chiron_b200.multistate import MultiStateSampler
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.reporters import MultistateReporter
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.testsystems import LennardJonesFluid
    ljs = [LennardJonesFluid(nparticles=1000, reduced_density=0.8, seed=5 + k) for k in range(N_REP)]
    potential = LJPotential(ljs[0].topology, ljs[0].sigma, ljs[0].epsilon, 1.02 * unit.nanometer)
    temps = [300.0 * 1.5 ** (k / (N_REP - 1.0)) for k in range(N_REP)]
    thermo = [ThermodynamicState(potential, temperature=t * unit.kelvin) for t in temps]
    keys = crandom.split(crandom.PRNGKey(99), N_REP)
    states = [SamplerState(ljs[k].positions, keys[k], box_vectors=ljs[0].box_vectors) for k in range(N_REP)]
    nbrs = [NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.02 * unit.nanometer, skin=0.5 * unit.nanometer,
                              n_max_neighbors=400, builder="cell") for _ in range(N_REP)]
    move = LangevinDynamicsMove(timestep=1.0 * unit.femtosecond, number_of_steps=STEPS)
    ms = MultiStateSampler(MCMCSampler(MoveSchedule([("LangevinDynamicsMove", move)])), MultistateReporter(),
                           exchange="neighbors", exchange_seed=17, mcmc_iterations_per_sweep=1)
    ms.create(thermo, states, nbrs)
    ms._offline_estimator = None
    ms._online_estimator = type("NoAnalysis", (), {"update": lambda self: None, "f_k": None})()
    ms._reporter._default_properties = ["u_kn", "state_index"]
    ms.run(SWEEPS)
    assert bool(ms._batched)
    xyz = ms._report_positions()["positions"]            # all-gathered: every rank sees every replica
    if rank == 0:
        np.savez(os.path.join(out_dir, f"world{world}.npz"), states=np.asarray(ms._replica_thermodynamic_states),
                 u=np.asarray(ms._energy_thermodynamic_states), xyz=xyz,
                 box=np.asarray(ljs[0].box_vectors.value_in_unit(unit.nanometer), dtype=np.float64).diagonal(),
                 hist=np.asarray(ms._reporter.get_property("state_index")))
    if world > 1:
        dist.destroy_process_group()


def test_remd_two_gpus_match_one_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_run, args=(1, port, str(tmp_path)), nprocs=1, join=True)
    mp.spawn(_run, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "world1.npz"), np.load(tmp_path / "world2.npz")
    assert np.array_equal(a["hist"], b["hist"])                    # the same swaps in every sweep
    assert np.array_equal(a["states"], b["states"]) and not np.array_equal(a["states"], np.arange(N_REP))
    assert np.allclose(a["u"], b["u"], rtol=1e-6)
    L = a["box"]
    dx = a["xyz"] - b["xyz"]
    dx -= L * np.round(dx / L)
    assert np.abs(dx).max() < 1e-4 * float(L.max())              # trajectories equal to fp32 summation order
    assert np.abs(b["xyz"]).sum(axis=(1, 2)).min() > 0             # no replica left at zero (positions all-gathered)

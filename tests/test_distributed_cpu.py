"""world_size-2 tests of the multi-rank driver (anisotropicelastoplasticity_b200/distributed.py) on CPU.

The driver is backend-agnostic; here each rank's "engine" is the CPU oracle restricted to its slab
(tests/oracle_slab_backend.py), so the real SlabSolver code -- halo exchange of the 3 shared node planes, the 4-byte
all-reduce(max) for dt, particle migration with global ids -- runs under torch.distributed/gloo and must reproduce the
single-process oracle (decomposition invariance)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import relerr


def _scene():
    from anisotropicelastoplasticity_b200 import scenes as sc
    s = sc.small_block(material=sc.SAND, res=16, cells=3, seed=21, lo=(0.3, 0.26, 0.3))
    s.particles.v[:, 1] += 4.0          # drive particles across the slab boundary (y)
    return s


def _reference(nsteps):
    from oracle.oracle_py import Oracle
    o = Oracle(_scene()); o.init()
    dts = [o.substep() for _ in range(nsteps)]
    return o.particles(), dts


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, nsteps, bounds, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from anisotropicelastoplasticity_b200.distributed import SlabPlan, SlabSolver
    from oracle_slab_backend import OracleSlabBackend
    plan = SlabPlan(1, bounds)
    b = OracleSlabBackend(_scene(), plan, rank)
    n0 = len(b.ids)
    s = SlabSolver(b, plan, rank)
    s.init()
    dts = []
    for _ in range(nsteps):
        s.substep(); dts.append(b.dt)
    parts = b.particles_local(); parts["dts"] = dts; parts["n0"] = n0; parts["migrated"] = s.stats["migrated"]
    gathered = [None] * world
    dist.gather_object(parts, gathered if rank == 0 else None, dst=0)
    if rank == 0:
        torch.save(gathered, out)
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.parametrize("bounds", [[0, 6, 16], [0, 5, 9, 16]])
def test_slab_driver_matches_single_process(tmp_path, bounds):
    nsteps = 6; world = len(bounds) - 1
    out = str(tmp_path / "gathered.pt")
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    mp.spawn(_worker, args=(world, _free_port(), nsteps, bounds, out), nprocs=world, join=True)
    gathered = torch.load(out, weights_only=False)
    ref, ref_dts = _reference(nsteps)
    ids = np.concatenate([g["ids"] for g in gathered]); order = np.argsort(ids)
    assert (np.sort(ids) == np.arange(ref["x"].shape[0])).all()                       # nobody lost or duplicated
    assert sum(g["migrated"] for g in gathered) > 0                                   # particles really crossed the boundary
    assert any(len(g["ids"]) != g["n0"] for g in gathered)
    for g in gathered:
        assert np.allclose(g["dts"], ref_dts, rtol=1e-12)                             # same dt on every rank (all-reduce max)
    for k in ("x", "v", "B", "FE", "FP", "vol", "q"):
        got = np.concatenate([g[k] for g in gathered], axis=0)[order]
        assert relerr(got, ref[k]) < 1e-11, k


def test_local_group_matches_single_process():
    """Same check through LocalSlabGroup (in-process lockstep, direct buffer swaps)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from anisotropicelastoplasticity_b200.distributed import LocalSlabGroup, SlabPlan
    from oracle_slab_backend import OracleSlabBackend
    plan = SlabPlan(1, [0, 6, 16])
    grp = LocalSlabGroup([OracleSlabBackend(_scene(), plan, r) for r in range(2)], plan)
    grp.init(); grp.run(6)
    got = grp.gather_particles(); ref, _ = _reference(6)
    for k in ("x", "v", "FE", "FP", "q"):
        assert relerr(got[k], ref[k]) < 1e-11, k


def test_slab_plan():
    from anisotropicelastoplasticity_b200.distributed import SlabPlan
    p = SlabPlan.uniform(512, 8)
    assert p.bounds == [0, 64, 128, 192, 256, 320, 384, 448, 512] and p.slab(3) == (1, 192, 256)
    cells = np.concatenate([np.full(1000, 10), np.full(1000, 200), np.full(2000, 400)])
    q = SlabPlan.balanced(cells, 512, 4)
    assert q.bounds[0] == 0 and q.bounds[-1] == 512 and all(b1 - b0 >= 4 for b0, b1 in zip(q.bounds, q.bounds[1:]))
    assert (p.owner_of_cells(np.array([0, 63, 64, 511])) == [0, 0, 1, 7]).all()
    with pytest.raises(ValueError):
        SlabPlan(1, [0, 2, 16])

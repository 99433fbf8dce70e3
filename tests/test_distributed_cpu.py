"""CPU, gloo, world_size 2: the multi-rank drivers (row-slab halo exchange, sharded replica exchange) give
bit-identical lattices to the single-process oracle run.  The engine is the oracle-backed stand-in."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slab_worker(rank, world, port, rows, cols, periodic, n_sweeps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_engine import OracleEngine
    from tsu_emulator_b200.distributed import SlabShardedIsing2D

    fac = lambda lr, r0: OracleEngine(lr, cols, n_replicas=2, temperature=2.269, periodic=periodic, seed=11, row0=r0,
                                      global_rows=rows).init_random()
    drv = SlabShardedIsing2D(rows, cols, fac, periodic=periodic)
    drv.sweep(n_sweeps)
    obs = drv.observables()
    q.put((rank, np.stack([drv.engine.bits(r) for r in range(2)]), obs.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("rows,cols,periodic", [(8, 12, True), (6, 10, False), (12, 64, True)])
def test_row_slabs_two_ranks_match_single_lattice(rows, cols, periodic):
    from oracle import ising2d_oracle as O

    world, n_sweeps = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, rows, cols, periodic, n_sweeps, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    [p.join(timeout=60) for p in procs]
    got = np.concatenate([r[1] for r in res], axis=1)  # [replica, rows, cols]
    for rep in range(2):
        start = O.init_bits(11, rep, rows, cols)
        want = O.checkerboard_sweeps_philox(start, 11, rep, 0, n_sweeps, 1.0, 0.0, 2.269, periodic)
        assert (got[rep] == want).all()
        up, anti = res[0][2][rep]
        assert up == want.sum()
        e = -(O.energy(want, 1.0, 0.0, periodic))
        from tsu_emulator_b200.lattice import lattice_bond_count
        nb = lattice_bond_count(rows, cols, periodic and rows > 2, periodic and cols > 2)
        assert nb - 2 * anti == e
    assert (res[0][2] == res[1][2]).all()  # all-reduced observables agree on both ranks


def test_single_rank_slab_driver_equals_plain_engine():
    from oracle import ising2d_oracle as O
    from oracle_engine import OracleEngine
    from tsu_emulator_b200.distributed import SlabShardedIsing2D

    fac = lambda lr, r0: OracleEngine(lr, 10, temperature=1.5, periodic=True, seed=3, row0=r0, global_rows=8).init_random()
    drv = SlabShardedIsing2D(8, 10, fac, periodic=True).sweep(2)
    want = O.checkerboard_sweeps_philox(O.init_bits(3, 0, 8, 10), 3, 0, 0, 2, 1.0, 0.0, 1.5, True)
    assert (drv.engine.bits(0) == want).all()


def _pt_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q.put((rank,) + _run_pt())
    dist.barrier()
    dist.destroy_process_group()


def _run_pt():
    from oracle_engine import OracleEngine, cpu_swap
    from tsu_emulator_b200.distributed import LatticeTempering

    temps = [1.0, 2.0, 2.5, 4.0]
    fac = lambda n, r0, T: OracleEngine(6, 8, n_replicas=n, temperature=T, periodic=True, seed=5, replica0=r0).init_random()
    pt = LatticeTempering(temps, n_ladders=3, engine_factory=fac, swap_fn=cpu_swap(77), n_sweeps=2, swap_interval=2, seed=77)
    for _ in range(6):
        pt.step()
    m, e = pt.observables_by_slot()
    return m, e, pt.slot_replica.numpy().copy(), pt.lut_index.numpy().copy()


def test_sharded_tempering_two_ranks_equals_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_pt_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda x: x[0])
    [p.join(timeout=60) for p in procs]
    m1, e1, sr1, li1 = _run_pt()          # world size 1, same seeds
    for r in res:
        assert np.array_equal(r[1], m1) and np.array_equal(r[2], e1)
        assert np.array_equal(r[3], sr1) and np.array_equal(r[4], li1)
    assert not np.array_equal(sr1, np.arange(12).reshape(3, 4))  # some swaps were accepted


def test_replica_shard_partitions():
    from tsu_emulator_b200.distributed import replica_shard

    for n, w in [(10, 3), (4096, 8), (5, 8), (50, 8)]:
        parts = [replica_shard(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1

"""GPU (single rank): the drivers of distributed.py on the real engine - halo path of the kernel and the
device-side replica-exchange pass - against the oracle."""
import numpy as np
import pytest

from oracle import ising2d_oracle as O

pytestmark = pytest.mark.gpu


def test_slab_driver_single_rank_uses_halo_path():
    from tsu_emulator_b200 import Ising2DEngine
    from tsu_emulator_b200.distributed import SlabShardedIsing2D

    for rows, cols in [(16, 256), (10, 12)]:
        fac = lambda lr, r0: Ising2DEngine(lr, cols, n_replicas=2, temperature=2.269, periodic=True, seed=9, row0=r0,
                                           global_rows=rows).init_random()
        drv = SlabShardedIsing2D(rows, cols, fac, periodic=True).sweep(3)
        got = drv.engine.get_spins(pm1=False)
        for r in range(2):
            want = O.checkerboard_sweeps_philox(O.init_bits(9, r, rows, cols), 9, r, 0, 3, 1.0, 0.0, 2.269, True)
            assert (got[r] == want).all()
        obs = drv.observables().cpu().numpy()
        assert obs[0, 0] == got[0].sum()


def test_two_slabs_emulated_on_one_gpu_match_single_lattice():
    """two slab engines on one device, halos copied by hand: exercises row0 / halo_top / halo_bot of the kernels"""
    import torch
    from tsu_emulator_b200 import Ising2DEngine

    rows, cols, seed, T = 24, 512, 21, 2.0
    for periodic in (True, False):
        slabs = [Ising2DEngine(rows // 2, cols, temperature=T, periodic=periodic, seed=seed, row0=r0, global_rows=rows).init_random()
                 for r0 in (0, rows // 2)]
        start = np.concatenate([s.get_spins(pm1=False)[0] for s in slabs])
        # get_spins of a slab with odd/even row0 must agree with the oracle's global init
        assert (start == O.init_bits(seed, 0, rows, cols)).all()
        for sweep in range(3):
            for colour in (0, 1):
                opp = 1 - colour
                tops = [slabs[1].state[:, opp, -1, :].contiguous() if periodic else None, slabs[0].state[:, opp, -1, :].contiguous()]
                bots = [slabs[1].state[:, opp, 0, :].contiguous(), slabs[0].state[:, opp, 0, :].contiguous() if periodic else None]
                for k, s in enumerate(slabs):
                    s.half_sweep(colour, halo_top=tops[k], halo_bot=bots[k])
            for s in slabs:
                s.sweep_index += 1
        got = np.concatenate([s.get_spins(pm1=False)[0] for s in slabs])
        want = O.checkerboard_sweeps_philox(start, seed, 0, 0, 3, 1.0, 0.0, T, periodic)
        assert (got == want).all()


def test_lattice_tempering_device_swap_matches_host_pass():
    import torch
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import cpu_swap
    from tsu_emulator_b200 import Ising2DEngine
    from tsu_emulator_b200.distributed import LatticeTempering

    temps = [2.0, 2.2, 2.4, 2.6, 2.8]

    def run(swap_fn):
        fac = lambda n, r0, T: Ising2DEngine(8, 16, n_replicas=n, temperature=T, periodic=True, seed=5, replica0=r0).init_random()
        pt = LatticeTempering(temps, n_ladders=4, engine_factory=fac, swap_fn=swap_fn, n_sweeps=3, swap_interval=2, seed=123)
        for _ in range(60):
            pt.step()
        m, e = pt.observables_by_slot()
        return m, e, pt.slot_replica.cpu().numpy(), pt.stats.cpu().numpy()

    m_dev, e_dev, sr_dev, stats = run(None)  # tsu_pt_swap kernel

    def host_swap(energy, T_slot, slot_replica, lut_index, K, R, step):
        e, t, sr, li = energy.cpu(), T_slot.cpu(), slot_replica.cpu(), lut_index.cpu()
        cpu_swap(123)(e, t, sr, li, K, R, step)
        slot_replica.copy_(sr)
        lut_index.copy_(li)

    m_host, e_host, sr_host, _ = run(host_swap)
    assert np.array_equal(sr_dev, sr_host) and np.array_equal(m_dev, m_host) and np.array_equal(e_dev, e_host)
    assert stats[0] == 30 * 4 * 4 and 0 < stats[1] <= stats[0]         # 30 passes x 4 ladders x 4 pairs
    assert e_dev.mean(0)[0] < e_dev.mean(0)[-1]                        # colder slots sit at lower energy


def _slab_rank(rank, world, port, rows, cols, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from tsu_emulator_b200 import Ising2DEngine
    from tsu_emulator_b200.distributed import SlabShardedIsing2D

    ok = True
    whole = Ising2DEngine(rows, cols, n_replicas=2, temperature=2.269, periodic=True, seed=7).init_random().sweep(3)
    lr = rows // world
    for overlap, transport in ((True, "p2p"), (True, "nccl"), (False, "nccl")):
        fac = lambda lr_, r0: Ising2DEngine(lr_, cols, n_replicas=2, temperature=2.269, periodic=True, seed=7, row0=r0,
                                            global_rows=rows).init_random()
        drv = SlabShardedIsing2D(rows, cols, fac, periodic=True, overlap=overlap, transport=transport)
        ok = ok and drv.transport() == transport
        drv.sweep(2).sweep(1)
        ok = ok and torch.equal(drv.engine.state, whole.state[:, :, rank * lr:(rank + 1) * lr, :])
        ok = ok and torch.equal(drv.observables(), whole.observables_tensor())
        drv.close()
    # open boundaries: the first and the last rank have one neighbour only
    whole = Ising2DEngine(rows, cols, temperature=2.0, periodic=False, seed=8).init_random().sweep(2)
    fac = lambda lr_, r0: Ising2DEngine(lr_, cols, temperature=2.0, periodic=False, seed=8, row0=r0, global_rows=rows).init_random()
    drv = SlabShardedIsing2D(rows, cols, fac, periodic=False, transport="p2p").sweep(2)
    ok = ok and torch.equal(drv.engine.state, whole.state[:, :, rank * lr:(rank + 1) * lr, :])
    drv.close()
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_row_slabs_on_all_gpus_match_single_gpu():
    """one rank per visible GPU (skipped on a one-GPU box): peer-mapped halos (one C-ABI call per sweep batch), NCCL
    send/recv overlapped with the interior update and the serial exchange all give the bits and the observables of the
    unsharded lattice, periodic and open"""
    import torch
    import torch.multiprocessing as mp
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if world < 4 else (4 if world < 8 else 8)
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_slab_rank, args=(r, world, port, 2048, 4096, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=300) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(ok for _, ok in res), res

"""
Multi-GPU drivers (one process per GPU, torch.distributed; NCCL over NVLink on the box, gloo in CPU tests).

  * SlabShardedIsing2D - ONE large lattice (or a batch of them) split into contiguous row slabs.  Per
    half-sweep every rank needs the opposite-colour row just above and just below its slab: the rows that
    were updated in the previous half-sweep are sent to the ring neighbours (wpr words per replica and
    side: 8 KiB for 131072 columns).  Philox counters use global row indices, so the bits are identical to
    the single-GPU run for any number of ranks.
  * replica_shard - independent replicas / chains / ladders: contiguous index ranges, no collective.
  * LatticeTempering - K temperature ladders x R temperatures of one lattice size with replica exchange
    (tsu/gibbs.py:238-338 semantics on the lattice kernels).  Lattices never move: the swap permutes the
    replica -> temperature-table index.  With a process group the replicas are sharded across ranks, the
    per-replica energies are all-gathered (K*R float64) and every rank evaluates the same deterministic
    swap pass.

The drivers only talk to an "engine" object (Ising2DEngine on GPUs); the CPU tests plug in an
oracle-backed stand-in to exercise the partitioning and exchange logic under gloo with world_size 2.
"""

from typing import Optional, Sequence

import numpy as np


def replica_shard(n_total: int, rank: int, world: int):
    """contiguous [start, stop) of `n_total` independent units owned by `rank`"""
    base, extra = divmod(n_total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist

    return dist


class SlabShardedIsing2D:
    """row-slab decomposition of a rows x cols lattice over the ranks of `group`.

    engine_factory(local_rows, row0) must return an engine exposing
        .state                      tensor [n_replicas, 2, local_rows, wpr] (int32 words)
        .half_sweep(colour, halo_top=..., halo_bot=...)   halos: [n_replicas, wpr] or None
        .sweep_index                incremented by the driver
        .observables_tensor(next_rows=...) -> int64 [n_replicas, 2]
    """

    def __init__(self, rows: int, cols: int, engine_factory, periodic: bool = True, group=None):
        dist = _dist()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if rows % self.world:
            raise ValueError("rows must be divisible by the number of ranks")
        self.rows, self.cols, self.periodic = rows, cols, periodic
        self.local_rows = rows // self.world
        self.row0 = self.rank * self.local_rows
        self.engine = engine_factory(self.local_rows, self.row0)
        import torch

        st = self.engine.state
        self.n_replicas, self.wpr = st.shape[0], st.shape[3]
        self.halo_top = torch.zeros((self.n_replicas, self.wpr), dtype=st.dtype, device=st.device)
        self.halo_bot = torch.zeros_like(self.halo_top)
        self._next_rows = torch.zeros((self.n_replicas, 2, self.wpr), dtype=st.dtype, device=st.device)
        self.up = (self.rank - 1) % self.world     # owns the rows above mine
        self.down = (self.rank + 1) % self.world   # owns the rows below mine
        self.has_up = self.periodic or self.rank > 0
        self.has_down = self.periodic or self.rank < self.world - 1

    # -- halo exchange ---------------------------------------------------------------------------
    def exchange(self, colour: int):
        """make rows (row0-1) and (row0+local_rows) of `colour` available as halo_top / halo_bot"""
        st = self.engine.state
        if self.world == 1:
            if self.periodic:
                self.halo_top.copy_(st[:, colour, -1, :])
                self.halo_bot.copy_(st[:, colour, 0, :])
            return
        dist = _dist()
        first = st[:, colour, 0, :].contiguous()
        last = st[:, colour, -1, :].contiguous()
        if self.world == 2 and self.periodic:
            # both neighbours are the same peer: order the two messages explicitly
            peer = self._peer(self.up)
            if self.rank == 0:
                dist.send(first, peer, self.group)
                dist.recv(self.halo_bot, peer, self.group)
                dist.send(last, peer, self.group)
                dist.recv(self.halo_top, peer, self.group)
            else:
                dist.recv(self.halo_bot, peer, self.group)
                dist.send(first, peer, self.group)
                dist.recv(self.halo_top, peer, self.group)
                dist.send(last, peer, self.group)
            return
        ops = []
        if self.has_up:
            ops.append(dist.P2POp(dist.isend, first, self._peer(self.up), self.group))
            ops.append(dist.P2POp(dist.irecv, self.halo_top, self._peer(self.up), self.group))
        if self.has_down:
            ops.append(dist.P2POp(dist.isend, last, self._peer(self.down), self.group))
            ops.append(dist.P2POp(dist.irecv, self.halo_bot, self._peer(self.down), self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def _peer(self, group_rank: int) -> int:
        dist = _dist()
        if self.group is None:
            return group_rank
        return dist.get_global_rank(self.group, group_rank)

    # -- updates -----------------------------------------------------------------------------------
    def half_sweep(self, colour: int):
        self.exchange(1 - colour)
        self.engine.half_sweep(colour, halo_top=self.halo_top if self.has_up else None,
                               halo_bot=self.halo_bot if self.has_down else None)

    def sweep(self, n_sweeps: int = 1):
        for _ in range(n_sweeps):
            self.half_sweep(0)
            self.half_sweep(1)
            self.engine.sweep_index += 1
        return self

    # -- observables -------------------------------------------------------------------------------
    def observables(self):
        """global (# up spins, # anti-aligned bonds) per replica, summed over the slabs"""
        import torch

        nxt = None
        if self.periodic or self.world > 1:
            for colour in (0, 1):          # collective: every rank takes part even if it has no lower neighbour
                self.exchange(colour)
                self._next_rows[:, colour, :] = self.halo_bot
            if self.has_down:
                nxt = self._next_rows
        obs = self.engine.observables_tensor(next_rows=nxt).clone()
        if self.world > 1:
            _dist().all_reduce(obs, group=self.group)
        return obs


class LatticeTempering:
    """K ladders x R temperatures of rows x cols lattices with replica exchange on the lattice kernels.

    Replica g = ladder * R + j starts at temperature slot j.  Every iteration = n_sweeps sweeps of all
    replicas; every `swap_interval` iterations one exchange pass per ladder (pairs i = 0..R-2 in order,
    Metropolis rule of tsu/gibbs.py:308-323).  `engine_factory(n_local, replica0, temperatures_local)` builds
    the engine of this rank's replica range; swap_fn(energy, T_slot, slot_replica, lut_index, K, R, step)
    performs the pass in place (tsu_pt_swap on GPUs).  criterion="metropolis" (default) is the detailed-balance
    rule; "reference" reproduces the expression of gibbs.py:317 (see include/tsu_b200.h).
    """

    def __init__(self, temperatures: Sequence[float], n_ladders: int, engine_factory, swap_fn=None, group=None,
                 n_sweeps: int = 10, swap_interval: int = 10, seed: int = 0, criterion: str = "metropolis"):
        import torch

        dist = _dist()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.temps = np.asarray(list(temperatures), dtype=np.float64)
        self.R, self.K = len(self.temps), int(n_ladders)
        self.n_total = self.R * self.K
        self.start, self.stop = replica_shard(self.n_total, self.rank, self.world)
        self.n_sweeps, self.swap_interval, self.seed = int(n_sweeps), int(swap_interval), int(seed)
        slot0 = np.arange(self.n_total) % self.R
        self.engine = engine_factory(self.stop - self.start, self.start, self.temps[slot0[self.start:self.stop]])
        dev = self.engine.state.device
        self.T_slot = torch.from_numpy(self.temps).to(dev)
        self.slot_replica = torch.arange(self.n_total, dtype=torch.int32, device=dev).reshape(self.K, self.R).contiguous()
        self.lut_index = torch.from_numpy(slot0.astype(np.int32)).to(dev)   # replica -> temperature slot
        self.stats = torch.zeros(2, dtype=torch.int64, device=dev)
        self.iteration = 0
        if criterion not in ("metropolis", "reference"):
            raise ValueError("criterion must be 'metropolis' or 'reference'")
        self.criterion = 1 if criterion == "metropolis" else 0
        self.swap_fn = swap_fn or self._swap_cuda
        # the engine's LUT tables must be ordered by slot: one table per temperature of the ladder
        self.engine.set_temperature_tables(self.temps, self.lut_index[self.start:self.stop])

    def _swap_cuda(self, energy, T_slot, slot_replica, lut_index, K, R, step):
        from . import _lib
        from ._lib import ptr

        _lib.call("tsu_pt_swap", ptr(energy), ptr(T_slot), ptr(slot_replica), ptr(lut_index), K, R, self.seed,
                  step & 0xFFFFFFFF, ptr(self.stats), None, self.criterion, _lib.current_stream())

    def gather_energies(self):
        import torch

        e_local = self.engine.energy_tensor()
        if self.world == 1:
            return e_local
        dist = _dist()
        sizes = [replica_shard(self.n_total, r, self.world) for r in range(self.world)]
        parts = [torch.empty(b - a, dtype=e_local.dtype, device=e_local.device) for a, b in sizes]
        dist.all_gather(parts, e_local.contiguous(), group=self.group)
        return torch.cat(parts)

    def step(self):
        """one iteration: n_sweeps sweeps, then (every swap_interval iterations) the exchange pass"""
        self.engine.sweep(self.n_sweeps)
        self.iteration += 1
        if self.iteration % self.swap_interval == 0:
            energy = self.gather_energies()
            self.swap_fn(energy, self.T_slot, self.slot_replica, self.lut_index, self.K, self.R, self.iteration)
            self.engine.set_lut_index(self.lut_index[self.start:self.stop])
        return self

    def observables_by_slot(self):
        """(magnetisation, energy) arrays of shape [K, R] ordered by temperature slot (host numpy)"""
        import torch

        obs = self.engine.observables_tensor().to(torch.float64)
        full = obs
        if self.world > 1:
            dist = _dist()
            sizes = [replica_shard(self.n_total, r, self.world) for r in range(self.world)]
            parts = [torch.empty((b - a, 2), dtype=obs.dtype, device=obs.device) for a, b in sizes]
            dist.all_gather(parts, obs.contiguous(), group=self.group)
            full = torch.cat(parts)
        sr = self.slot_replica.long()
        per_slot = full[sr.reshape(-1)].reshape(self.K, self.R, 2).cpu().numpy()
        n = self.engine.n_sites
        m = (2.0 * per_slot[..., 0] - n) / n
        e = -self.engine.coupling * (self.engine.n_bonds - 2.0 * per_slot[..., 1]) - self.engine.field * (2.0 * per_slot[..., 0] - n)
        return m, e

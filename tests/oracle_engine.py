"""CPU stand-in for Ising2DEngine backed by the oracle (tests only): same state layout and driver-facing
methods, so the multi-rank drivers of tsu_emulator_b200/distributed.py can run under gloo without a GPU."""
import numpy as np
import torch

from oracle import dense_oracle as D
from oracle import ising2d_oracle as O


class OracleEngine:
    def __init__(self, rows, cols, n_replicas=1, coupling=1.0, field=0.0, temperature=1.0, periodic=True, seed=0,
                 replica0=0, row0=0, global_rows=None):
        self.rows, self.cols, self.n_replicas = rows, cols, n_replicas
        self.coupling, self.field = coupling, field
        self.periodic, self.seed, self.replica0, self.row0 = periodic, seed, replica0, row0
        self.global_rows = global_rows if global_rows is not None else rows
        self.is_slab = self.global_rows != rows
        self.wrap_cols = periodic and cols > 2
        self.wrap_rows = periodic and self.global_rows > 2
        self.sweep_index = 0
        self.wpr = O.words_per_row(cols)
        self.temps = np.broadcast_to(np.asarray(temperature, dtype=np.float64), (n_replicas,)).copy()
        self.state = torch.zeros((n_replicas, 2, rows, self.wpr), dtype=torch.int32)
        self.n_sites = rows * cols

    # -- helpers
    def bits(self, r):
        return O.unpack_spins(self.state[r].numpy().view(np.uint32), self.rows, self.cols, self.row0)

    def set_bits(self, r, bits):
        self.state[r] = torch.from_numpy(O.pack_spins(bits, self.row0).view(np.int32))

    def init_random(self):
        for r in range(self.n_replicas):
            self.set_bits(r, O.init_bits(self.seed, self.replica0 + r, self.rows, self.cols, self.row0))
        return self

    def _row_from_words(self, words, row_g, colour):
        """full-width row with only the sites of `colour` filled in from packed words"""
        row = np.zeros(self.cols, dtype=np.int64)
        p = (row_g + colour) & 1
        n = O.colour_count(self.cols, row_g, colour)
        k = np.arange(n)
        w = words.numpy().view(np.uint32)
        row[p::2] = (w[k >> 5] >> (k & 31).astype(np.uint32)) & 1
        return row

    def half_sweep(self, colour, halo_top=None, halo_bot=None, uniforms=None, rows=None):
        for r in range(self.n_replicas):
            b = self.bits(r)
            before = b.copy()
            top = bot = None
            if halo_top is not None:
                top = self._row_from_words(halo_top[r], self.row0 - 1, 1 - colour)
            elif self.wrap_rows and not self.is_slab:
                top = b[-1].copy()
            if halo_bot is not None:
                bot = self._row_from_words(halo_bot[r], self.row0 + self.rows, 1 - colour)
            elif self.wrap_rows and not self.is_slab:
                bot = b[0].copy()
            u = O.philox_uniform_field(self.seed, self.replica0 + r, self.sweep_index, self.rows, self.cols, self.row0)
            O.half_sweep_slab(b, top, bot, colour, u, self.coupling, self.field, float(self.temps[r]), self.wrap_cols,
                              self.row0)
            if rows is not None:  # only the rows [begin, end) are updated (a half-sweep reads the other colour only)
                keep = np.ones(self.rows, dtype=bool)
                keep[rows[0]:rows[1]] = False
                b[keep] = before[keep]
            self.set_bits(r, b)

    def sweep(self, n=1):
        for _ in range(n):
            self.half_sweep(0)
            self.half_sweep(1)
            self.sweep_index += 1
        return self

    def observables_tensor(self, next_rows=None):
        out = torch.zeros((self.n_replicas, 2), dtype=torch.int64)
        for r in range(self.n_replicas):
            b = self.bits(r)
            up = int(b.sum())
            anti = int((b[:, :-1] != b[:, 1:]).sum() + (b[:-1] != b[1:]).sum())
            if self.wrap_cols:
                anti += int((b[:, -1] != b[:, 0]).sum())
            if next_rows is not None:
                nxt = np.zeros(self.cols, dtype=np.int64)
                for colour in (0, 1):
                    row = self._row_from_words(next_rows[r, colour], self.row0 + self.rows, colour)
                    p = (self.row0 + self.rows + colour) & 1
                    nxt[p::2] = row[p::2]
                anti += int((b[-1] != nxt).sum())
            elif self.wrap_rows and not self.is_slab:
                anti += int((b[-1] != b[0]).sum())
            out[r, 0], out[r, 1] = up, anti
        return out

    # -- tempering hooks
    @property
    def n_bonds(self):
        from tsu_emulator_b200.lattice import lattice_bond_count
        return lattice_bond_count(self.rows, self.cols, self.wrap_rows, self.wrap_cols)

    def set_temperature_tables(self, temps, lut_index):
        self._tables = np.asarray(temps, dtype=np.float64)
        self.set_lut_index(lut_index)

    def set_lut_index(self, lut_index):
        self.temps = self._tables[np.asarray(lut_index)]

    def energy_tensor(self):
        obs = self.observables_tensor().numpy().astype(np.float64)
        e = -self.coupling * (self.n_bonds - 2.0 * obs[:, 1]) - self.field * (2.0 * obs[:, 0] - self.n_sites)
        return torch.from_numpy(e)


def cpu_swap(seed, criterion=1):
    """host replica-exchange pass with the kernel's semantics and Philox stream (csrc/dense_gibbs.cu: pt_swap_kernel)"""
    from oracle.philox_ref import philox4x32_10

    def fn(energy, T_slot, slot_replica, lut_index, K, R, step):
        e, T = energy.numpy(), T_slot.numpy()
        sr = slot_replica.numpy()
        for ladder in range(K):
            for i in range(R - 1):
                ra, rb = sr[ladder, i], sr[ladder, i + 1]
                delta = (1.0 / T[i] - 1.0 / T[i + 1]) * ((e[ra] - e[rb]) if criterion else (e[rb] - e[ra]))
                acc = delta >= 0
                if not acc:
                    o = philox4x32_10(i, ladder, step & 0xFFFFFFFF, D.STREAM_PT_SWAP, seed & 0xFFFFFFFF, seed >> 32)
                    m = ((int(o[0]) << 32) | int(o[1])) >> 11
                    acc = m * (1.0 / 9007199254740992.0) < np.exp(delta)
                if acc:
                    sr[ladder, i], sr[ladder, i + 1] = rb, ra
            for i in range(R):
                lut_index[sr[ladder, i]] = i
    return fn

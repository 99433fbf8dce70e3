"""
Batched 2-D Ising lattice engine on one B200: host side of the bit-packed checkerboard kernels
(csrc/ising2d.cu, C-ABI tsu_ising2d_* in include/tsu_b200.h).

It replaces, for nearest-neighbour lattices, the reference path
    IsingGrid.__init__ (dense N x N J)            tsu/models/ising.py:320-361
    IsingModel.sample -> GibbsSampler.gibbs_sweep  tsu/models/ising.py:150-181, tsu/gibbs.py:128-162
without ever materialising J: the acceptance probabilities sigmoid(h_bit/T) of the (at most 15)
distinct local environments are tabulated on the host in float64 with the reference's own
formula and handed to the kernel as integer thresholds.
"""

import math
from typing import Optional, Sequence, Union

import numpy as np

from . import _lib
from ._lib import c_void_p, ptr

LUT_WORDS = 32


def sigmoid_clamped(x: float) -> float:
    """tsu/gibbs.py:61-77: sigma(x) with x > 20 -> 1.0, x < -20 -> 0.0 (strict), float64."""
    if x > 20:
        return 1.0
    elif x < -20:
        return 0.0
    return 1.0 / (1.0 + np.exp(-x))


def build_lut(J: float, h: float, T: float, bias_mode: str = "physical") -> np.ndarray:
    """threshold table uint32[32] for one (J, h, T).

    Entry d*5+u: site with d neighbours, u of them up.  Bit model of tsu/models/ising.py:127-148:
    J_bit = 4J, local field = 4J*u + h_bit (tsu/gibbs.py:96-99), with
      bias_mode "physical":  h_bit =  2h - 2*rowsum(J)   (correct spin->bit transformation)
      bias_mode "reference": h_bit = -2h + 2*rowsum(J)   (what ising.py:148 returns)
    p = sigmoid_clamped(field / T); threshold = ceil(p * 2^32) so that for an integer uniform k,
    (k / 2^32 < p) == (k < threshold).  Entry 25: bit mask of classes with p == 1.0.
    """
    if T <= 0:
        raise ValueError("Temperature must be positive")
    lut = np.zeros(LUT_WORDS, dtype=np.uint32)
    always = 0
    for d in range(5):
        rowsum = J * d
        if bias_mode == "physical":
            bias = 2 * h - 2 * rowsum
        elif bias_mode == "reference":
            bias = -2 * h + 2 * rowsum
        else:
            raise ValueError("bias_mode must be 'physical' or 'reference'")
        for u in range(5):
            field = float(4 * J * u) + float(bias)
            p = sigmoid_clamped(field / T)
            t = int(math.ceil(p * 4294967296.0))
            if t >= 4294967296:
                always |= 1 << (d * 5 + u)
                t = 4294967295
            lut[d * 5 + u] = t
    lut[25] = always
    return lut


def lattice_bond_count(rows: int, cols: int, wrap_rows: bool, wrap_cols: bool) -> int:
    """number of distinct bonds wired by tsu/models/ising.py:343-361"""
    return rows * (cols - 1) + (rows if wrap_cols else 0) + (rows - 1) * cols + (cols if wrap_rows else 0)


class Ising2DEngine:
    """n_replicas independent rows x cols lattices (or one row-slab of each) resident in HBM.

    State: torch int32 tensor [n_replicas, 2, rows, wpr] (bit-packed, layout in include/tsu_b200.h).
    Replica r uses temperature temperatures[r] (scalar = shared).  All randomness is
    Philox(seed; replica0+r, sweep index, global row, word), so a lattice split into row slabs
    (row0/halos) or a replica batch split across GPUs (replica0) reproduces the single-GPU bits.
    """

    def __init__(
        self,
        rows: int,
        cols: int,
        n_replicas: int = 1,
        coupling: float = 1.0,
        field: float = 0.0,
        temperature: Union[float, Sequence[float]] = 1.0,
        periodic: bool = True,
        seed: int = 0,
        device=None,
        bias_mode: str = "physical",
        replica0: int = 0,
        row0: int = 0,
        global_rows: Optional[int] = None,
    ):
        torch = _lib.require_cuda()
        self._torch = torch
        self.lib = _lib.load()
        if rows <= 0 or cols <= 0 or n_replicas <= 0:
            raise ValueError("rows, cols and n_replicas must be positive")
        self.rows, self.cols, self.n_replicas = int(rows), int(cols), int(n_replicas)
        self.global_rows = int(global_rows) if global_rows is not None else self.rows
        self.row0 = int(row0)
        self.is_slab = self.global_rows != self.rows
        if periodic and (self.global_rows % 2 or cols % 2):
            raise ValueError("periodic checkerboard lattices need even rows and cols")
        self.periodic = bool(periodic)
        # a periodic dimension of size 2 adds no new bond in the reference (set_coupling assigns)
        self.wrap_rows = self.periodic and self.global_rows > 2
        self.wrap_cols = self.periodic and cols > 2
        self.coupling, self.field = float(coupling), float(field)
        self.bias_mode = bias_mode
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.replica0 = int(replica0)
        self.sweep_index = 0
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.wpr = int(self.lib.tsu_ising2d_words_per_row(self.cols))
        with torch.cuda.device(self.device):
            self.state = torch.zeros((self.n_replicas, 2, self.rows, self.wpr), dtype=torch.int32, device=self.device)
            self._obs = torch.zeros((self.n_replicas, 2), dtype=torch.int64, device=self.device)
        self.lut = None
        self.lut_index = None
        self.set_temperature(temperature)

    # ------------------------------------------------------------------ parameters
    def set_temperature(self, temperature):
        torch = self._torch
        temps = np.atleast_1d(np.asarray(temperature, dtype=np.float64))
        if temps.size not in (1, self.n_replicas):
            raise ValueError("temperature must be a scalar or one value per replica")
        if np.any(temps <= 0):
            raise ValueError("Temperature must be positive")
        self.temperatures = temps.copy()
        uniq, inv = np.unique(temps, return_inverse=True)
        luts = np.stack([build_lut(self.coupling, self.field, float(t), self.bias_mode) for t in uniq])
        self.lut = torch.from_numpy(luts.view(np.int32)).to(self.device)
        if temps.size == 1:
            self.lut_index = None
        else:
            self.lut_index = torch.from_numpy(inv.astype(np.int32)).to(self.device)
        self._lut_temps = uniq
        self._jit = self._jit_prepare(luts[0]) if temps.size == 1 and self._jit_worthwhile() else 0

    # NVRTC needs 1-2 s per new temperature and the specialised kernel is ~17 % faster, so it is compiled
    # automatically only where it can pay back within seconds: the wide full-word kernel on >= 2^28 sites.
    JIT_MIN_SITES = 1 << 28

    def _jit_worthwhile(self, force: bool = False) -> bool:
        import os

        if os.environ.get("TSU_B200_NO_JIT"):
            return False
        wide = (self.cols // 2) // 32 >= 4 and self.rows >= 3   # at least one 4-word group of full words per row
        if not wide:
            return False  # only the wide full-word kernel has a specialised form (open rims run the generic kernel)
        if force or os.environ.get("TSU_B200_JIT"):
            return True
        return self.n_replicas * self.rows * self.cols >= self.JIT_MIN_SITES

    def specialise(self) -> bool:
        """compile (or fetch from the per-process cache) the half-sweep kernel specialised for the current temperature
        now, whatever the lattice size; returns whether the specialised kernel is in use.  Results are bit-identical."""
        if self.lut_index is None and self._jit_worthwhile(force=True):
            lut_host = self.lut[0].cpu().numpy().view(np.uint32)
            self._jit = self._jit_prepare(lut_host)
        return self._jit > 0

    def _jit_prepare(self, lut_host: np.ndarray) -> int:
        """table-specialised build of the fast kernel (0 = unavailable: the prebuilt kernels are used)"""
        import ctypes
        import os

        if os.environ.get("TSU_B200_NO_JIT"):
            return 0
        src_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc").encode()
        log = ctypes.create_string_buffer(4096)
        arr = (ctypes.c_uint32 * LUT_WORDS)(*[int(x) for x in lut_host])
        with self._torch.cuda.device(self.device):
            handle = int(self.lib.tsu_ising2d_jit_prepare(arr, src_dir, log, 4096))
        self._jit_log = log.value.decode(errors="replace")
        return max(handle, 0)

    def set_temperature_tables(self, temperatures, lut_index):
        """one threshold table per entry of `temperatures` (kept in that order) and an explicit
        replica -> table map (int32 tensor [n_replicas]); used by replica exchange, which permutes the map"""
        torch = self._torch
        temps = np.asarray(temperatures, dtype=np.float64)
        if np.any(temps <= 0):
            raise ValueError("Temperature must be positive")
        luts = np.stack([build_lut(self.coupling, self.field, float(t), self.bias_mode) for t in temps])
        self.lut = torch.from_numpy(luts.view(np.int32)).to(self.device)
        self._lut_temps = temps
        self._jit = 0
        self.set_lut_index(lut_index)

    def set_lut_index(self, lut_index):
        torch = self._torch
        idx = torch.as_tensor(lut_index, dtype=torch.int32, device=self.device).contiguous()
        if idx.numel() != self.n_replicas:
            raise ValueError("one table index per replica")
        self.lut_index = idx.clone()
        self.temperatures = self._lut_temps[self.lut_index.cpu().numpy()] if hasattr(self, "_lut_temps") else self.temperatures

    def chunk_view(self, start: int, count: int):
        """engine over replicas [start, start+count) sharing this engine's HBM state (no copy)"""
        cache = self.__dict__.setdefault("_views", {})
        key = (int(start), int(count))
        v = cache.get(key)
        if v is None:
            if start < 0 or count <= 0 or start + count > self.n_replicas:
                raise ValueError("chunk out of range")
            v = object.__new__(Ising2DEngine)
            v.__dict__.update({k: val for k, val in self.__dict__.items() if k != "_views"})
            v.n_replicas = int(count)
            v.replica0 = self.replica0 + int(start)
            v.state = self.state[start:start + count]
            v._obs = self._obs[start:start + count]
            v.temperatures = self.temperatures if self.temperatures.size == 1 else self.temperatures[start:start + count]
            cache[key] = v
        v.lut = self.lut
        v.lut_index = None if self.lut_index is None else self.lut_index[start:start + count]
        return v

    # ------------------------------------------------------------------ state i/o
    def init_random(self):
        """iid Bernoulli(1/2) spins from the Philox init stream (np.random.randint of gibbs.py:201)."""
        with self._torch.cuda.device(self.device):
            _lib.call(
                "tsu_ising2d_init_random", ptr(self.state), self.n_replicas, self.rows, self.cols, self.seed,
                self.replica0, self.row0, _lib.current_stream(),
            )
        return self

    def set_spins(self, spins):
        """spins: array [n_replicas, rows, cols] (or [rows, cols]) in {-1,+1} or {0,1}; > 0 means up."""
        torch = self._torch
        a = np.asarray(spins)
        if a.ndim == 2:
            a = a[None]
        if a.shape != (self.n_replicas, self.rows, self.cols):
            raise ValueError(f"expected spins of shape {(self.n_replicas, self.rows, self.cols)}, got {a.shape}")
        host = torch.from_numpy(np.ascontiguousarray((a > 0).astype(np.int8)))
        with torch.cuda.device(self.device):
            dev = host.to(self.device)
            _lib.call("tsu_ising2d_pack", ptr(dev), ptr(self.state), self.n_replicas, self.rows, self.cols,
                      _lib.current_stream())
        return self

    def spins_tensor(self, pm1: bool = True):
        """int8 tensor [n_replicas, rows, cols] on the device"""
        torch = self._torch
        with torch.cuda.device(self.device):
            out = torch.empty((self.n_replicas, self.rows, self.cols), dtype=torch.int8, device=self.device)
            _lib.call("tsu_ising2d_unpack", ptr(self.state), ptr(out), self.n_replicas, self.rows, self.cols,
                      1 if pm1 else 0, _lib.current_stream())
        return out

    def get_spins(self, pm1: bool = True) -> np.ndarray:
        return self.spins_tensor(pm1).cpu().numpy().astype(np.int64)

    # ------------------------------------------------------------------ updates
    def half_sweep(self, colour: int, halo_top=None, halo_bot=None, uniforms=None, rows=None):
        """resample every site of `colour`; halos are [n_replicas, wpr] int32 rows of the other colour.
        rows=(begin, end) restricts the update to those local rows (same bits for any split of the rows)."""
        with self._torch.cuda.device(self.device):
            if rows is not None:
                if uniforms is not None:
                    raise ValueError("row ranges are not available in injected-uniform mode")
                _lib.call(
                    "tsu_ising2d_half_sweep_rows", int(getattr(self, "_jit", 0)) if self.lut_index is None else 0,
                    ptr(self.state), self.n_replicas, self.rows, self.cols,
                    int(self.wrap_rows and not self.is_slab), int(self.wrap_cols), int(colour), ptr(self.lut),
                    ptr(self.lut_index), self.seed, self.sweep_index & 0xFFFFFFFF, self.replica0, self.row0,
                    ptr(halo_top), ptr(halo_bot), int(rows[0]), int(rows[1]), _lib.current_stream(),
                )
            elif uniforms is None and self.lut_index is None and getattr(self, "_jit", 0) > 0:
                _lib.call(
                    "tsu_ising2d_half_sweep_jit", self._jit, ptr(self.state), self.n_replicas, self.rows, self.cols,
                    int(self.wrap_rows and not self.is_slab), int(self.wrap_cols), int(colour), ptr(self.lut),
                    self.seed, self.sweep_index & 0xFFFFFFFF, self.replica0, self.row0, ptr(halo_top), ptr(halo_bot),
                    _lib.current_stream(),
                )
            elif uniforms is None:
                _lib.call(
                    "tsu_ising2d_half_sweep", ptr(self.state), self.n_replicas, self.rows, self.cols,
                    int(self.wrap_rows and not self.is_slab), int(self.wrap_cols), int(colour), ptr(self.lut),
                    ptr(self.lut_index), self.seed, self.sweep_index & 0xFFFFFFFF, self.replica0, self.row0,
                    ptr(halo_top), ptr(halo_bot), _lib.current_stream(),
                )
            else:
                _lib.call(
                    "tsu_ising2d_half_sweep_injected", ptr(self.state), self.n_replicas, self.rows, self.cols,
                    int(self.wrap_rows and not self.is_slab), int(self.wrap_cols), int(colour), ptr(self.lut),
                    ptr(self.lut_index), ptr(uniforms), self.row0, ptr(halo_top), ptr(halo_bot),
                    _lib.current_stream(),
                )

    def sweep(self, n_sweeps: int = 1):
        """n full sweeps: black half-sweep then white half-sweep (one gibbs_update each)."""
        if self.is_slab:
            raise RuntimeError("row-slab engines are driven by ShardedIsing2D (halo exchange between half-sweeps)")
        with self._torch.cuda.device(self.device):
            if self.lut_index is None and getattr(self, "_jit", 0) > 0:
                _lib.call(
                    "tsu_ising2d_sweeps_jit", self._jit, ptr(self.state), self.n_replicas, self.rows, self.cols,
                    int(self.wrap_rows), int(self.wrap_cols), ptr(self.lut), self.seed, self.sweep_index & 0xFFFFFFFF,
                    int(n_sweeps), self.replica0, _lib.current_stream(),
                )
            else:
                _lib.call(
                    "tsu_ising2d_sweeps", ptr(self.state), self.n_replicas, self.rows, self.cols, int(self.wrap_rows),
                    int(self.wrap_cols), ptr(self.lut), ptr(self.lut_index), self.seed, self.sweep_index & 0xFFFFFFFF,
                    int(n_sweeps), self.replica0, _lib.current_stream(),
                )
        self.sweep_index += int(n_sweeps)
        return self

    def sweep_injected(self, uniforms_u32):
        """parity mode: uniforms_u32 [n_sweeps, n_replicas, rows, cols] integers k (meaning k / 2^32)."""
        torch = self._torch
        u = np.asarray(uniforms_u32)
        if u.ndim == 3:
            u = u[:, None]
        dev = torch.from_numpy(np.ascontiguousarray(u.astype(np.uint32)).view(np.int32)).to(self.device)
        for t in range(dev.shape[0]):
            for colour in (0, 1):
                self.half_sweep(colour, uniforms=dev[t])
            self.sweep_index += 1
        return self

    # ------------------------------------------------------------------ observables
    def observables_tensor(self, next_rows=None):
        """int64 tensor [n_replicas, 2]: (# up spins, # anti-aligned right+down bonds) of the local rows"""
        with self._torch.cuda.device(self.device):
            _lib.call(
                "tsu_ising2d_observables", ptr(self.state), self.n_replicas, self.rows, self.cols,
                int(self.wrap_rows and not self.is_slab), int(self.wrap_cols), self.row0, ptr(next_rows),
                ptr(self._obs), _lib.current_stream(),
            )
        return self._obs

    @property
    def n_sites(self) -> int:
        return self.rows * self.cols

    @property
    def n_bonds(self) -> int:
        return lattice_bond_count(self.rows, self.cols, self.wrap_rows, self.wrap_cols)

    def magnetization(self) -> np.ndarray:
        """signed magnetisation per spin of every replica (tsu/models/ising.py:183-193 on the current state)"""
        obs = self.observables_tensor().cpu().numpy()
        return (2.0 * obs[:, 0] - self.n_sites) / self.n_sites

    def energy(self) -> np.ndarray:
        """E = -J sum_<ij> s_i s_j - h sum_i s_i of every replica (tsu/models/ising.py:98-117)"""
        obs = self.observables_tensor().cpu().numpy().astype(np.float64)
        return -self.coupling * (self.n_bonds - 2.0 * obs[:, 1]) - self.field * (2.0 * obs[:, 0] - self.n_sites)

    def energy_tensor(self):
        """float64 device tensor of energies (no host sync) - feeds the replica-exchange kernel"""
        torch = self._torch
        obs = self.observables_tensor()
        with torch.cuda.device(self.device):
            e = torch.empty(self.n_replicas, dtype=torch.float64, device=self.device)
            _lib.call(
                "tsu_ising2d_energy_from_observables", ptr(obs), self.n_replicas, self.coupling, self.field,
                self.n_bonds, self.n_sites, ptr(e), _lib.current_stream(),
            )
        return e

"""
Drop-in mirror of the reference's tsu/models/ising.py (IsingConfig, IsingModel, IsingChain, IsingGrid,
demonstrate_phase_transition) plus README's IsingModel2D (README.md:116-131), with every sweep on the
B200.

  * IsingModel / IsingChain: arbitrary couplings -> dense-J kernel through GibbsSampler (as in the
    reference, ising.py:150-181).
  * IsingGrid / IsingModel2D: nearest-neighbour lattices -> bit-packed checkerboard kernel (lattice.py);
    the dense N x N matrix the reference builds (ising.py:343-361) is only materialised if `.J` is read.

Bias sign.  The reference's spin->bit bias is sign-flipped (ising.py:140-148 returns -2h + 2 rowsum(J);
the correct transformation of its own Hamiltonian is 2h - 2 rowsum(J)): with it a ferromagnet orders at
every temperature.  Default here is the physically correct bias; `compat_reference_bias=True`
reproduces the reference's numbers bit-for-bit (that mode is what the parity tests pin).
Update order: lattices use checkerboard order (black sites then white sites), the reference visits
sites in index order; both are valid heat-bath sweeps with the same single-site rule.
"""

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from ..gibbs import GibbsConfig, GibbsSampler
from ..lattice import Ising2DEngine


@dataclass
class IsingConfig:
    """tsu/models/ising.py:25-36"""

    temperature: float = 1.0
    external_field: float = 0.0
    n_burnin: int = 100
    n_sweeps: int = 10

    def __post_init__(self):
        if self.temperature <= 0:
            raise ValueError("Temperature must be positive")


def _resolve_compat_bias(flag: Optional[bool]) -> bool:
    """explicit argument, else the process-wide switch TSU_COMPAT_REFERENCE_BIAS=1 (run code written against the
    reference with the reference's spin-to-bit bias, ising.py:140-148, without touching it), else the correct bias"""
    if flag is not None:
        return bool(flag)
    import os

    return os.environ.get("TSU_COMPAT_REFERENCE_BIAS", "0") not in ("", "0")


class IsingModel:
    """General Ising model on an arbitrary graph (tsu/models/ising.py:39-262).

    Constructors: IsingModel(n_spins, config=None)                         (code, ising.py:52-75)
                  IsingModel(J=J, h=h, temperature=T)                      (README.md:136-143)
    """

    def __init__(self, n_spins: Optional[int] = None, config: Optional[IsingConfig] = None, *, J=None, h=None,
                 temperature: Optional[float] = None, compat_reference_bias: Optional[bool] = None,
                 seed: Optional[int] = None):
        if n_spins is None:
            if J is None:
                raise ValueError("give n_spins or a coupling matrix J")
            n_spins = int(np.asarray(J).shape[0])
        self.n_spins = int(n_spins)
        if config is None:
            config = IsingConfig(temperature=temperature) if temperature is not None else IsingConfig()
        self.config = config
        self.compat_reference_bias = _resolve_compat_bias(compat_reference_bias)
        # subclasses that know their wiring (IsingChain, IsingGrid) build the dense matrix only on demand
        self._J = None if getattr(self, "_lazy", False) else np.zeros((self.n_spins, self.n_spins))
        self.h = np.ones(self.n_spins) * self.config.external_field
        if J is not None:
            Jm = np.asarray(J, dtype=np.float64)
            if Jm.shape != (self.n_spins, self.n_spins):
                raise ValueError("Coupling matrix must be square")
            self._J = Jm.copy()
        if h is not None:
            self.set_external_field(np.asarray(h, dtype=np.float64))
        gibbs_config = GibbsConfig(
            temperature=self.config.temperature, n_burnin=self.config.n_burnin, n_sweeps=self.config.n_sweeps
        )
        self.sampler = GibbsSampler(gibbs_config, seed=seed)

    # J is a plain attribute here; IsingGrid overrides it with a lazily built matrix
    @property
    def J(self) -> np.ndarray:
        return self._J

    @J.setter
    def J(self, value):
        self._J = np.asarray(value, dtype=np.float64)

    def set_coupling(self, i: int, j: int, strength: float):
        """ising.py:77-86 (symmetric assignment)"""
        self.J[i, j] = strength
        self.J[j, i] = strength

    def set_external_field(self, field: np.ndarray):
        """ising.py:88-96"""
        if len(field) != self.n_spins:
            raise ValueError(f"Field must have length {self.n_spins}")
        self.h = np.array(field)

    def energy(self, state: np.ndarray) -> float:
        """ising.py:98-117: E = -1/2 s^T J s - h^T s on +-1 spins"""
        state = np.asarray(state)
        interaction_energy = -0.5 * state.dot(self.J).dot(state)
        field_energy = -self.h.dot(state)
        return interaction_energy + field_energy

    def _spins_to_bits(self, spins: np.ndarray) -> np.ndarray:
        return ((np.asarray(spins) + 1) // 2).astype(int)

    def _bits_to_spins(self, bits: np.ndarray) -> np.ndarray:
        return 2 * np.asarray(bits) - 1

    def _get_bit_coupling(self) -> np.ndarray:
        """ising.py:127-138: J_bit = 4 J"""
        return 4 * self.J

    def _get_bit_bias(self) -> np.ndarray:
        """bit bias.  Correct transformation: 2h - 2 rowsum(J); the reference returns the negative
        (ising.py:140-148), reproduced when compat_reference_bias is set."""
        if self.compat_reference_bias:
            return -2 * self.h + 2 * np.sum(self.J, axis=1)
        return 2 * self.h - 2 * np.sum(self.J, axis=1)

    def _sync_sampler_config(self):
        # callers mutate model.config / sampler.config between calls (ising.py:491-492): re-read every time
        self.sampler.config.temperature = self.config.temperature
        self.sampler.config.n_burnin = self.config.n_burnin
        self.sampler.config.n_sweeps = self.config.n_sweeps

    def _chromatic(self) -> bool:
        """mostly empty coupling matrices run on the chromatic CSR kernel (a sweep costs nnz, not N^2)"""
        from ..sparse import is_sparse_enough

        return is_sparse_enough(self.J)

    def sample(self, n_samples: int = 1000, initial_state: Optional[np.ndarray] = None) -> np.ndarray:
        """ising.py:150-181: (n_samples, n_spins) configurations in {-1,+1} from the Boltzmann distribution"""
        self._sync_sampler_config()
        initial_bits = self._spins_to_bits(initial_state) if initial_state is not None else None
        bit_samples = self.sampler.sample_boltzmann(
            self._get_bit_coupling(), bias=self._get_bit_bias(), n_samples=n_samples, initial_state=initial_bits,
            chromatic=self._chromatic(),
        )
        return self._bits_to_spins(bit_samples)

    def magnetization(self, samples: np.ndarray) -> float:
        """ising.py:183-193"""
        return np.mean(np.sum(samples, axis=1)) / self.n_spins

    def specific_heat(self, samples: np.ndarray) -> float:
        """ising.py:195-213: C = (<E^2> - <E>^2) / (T^2 N)"""
        energies = self._energies(samples)
        T = self.config.temperature
        return float((np.mean(energies**2) - np.mean(energies) ** 2) / (T**2 * self.n_spins))

    def _energies(self, samples: np.ndarray) -> np.ndarray:
        s = np.asarray(samples, dtype=np.float64)
        return -0.5 * np.einsum("ki,ij,kj->k", s, self.J, s) - s @ self.h

    def susceptibility(self, samples: np.ndarray) -> float:
        """ising.py:215-233: chi = (<m^2> - <m>^2) N / T with the signed magnetisation per spin"""
        magnetizations = np.sum(samples, axis=1) / self.n_spins
        T = self.config.temperature
        return (np.mean(magnetizations**2) - np.mean(magnetizations) ** 2) * self.n_spins / T

    def find_ground_state(self, n_steps: int = 1000) -> Tuple[np.ndarray, float]:
        """ising.py:235-262: simulated annealing from 10 T to 0.01 T"""
        self._sync_sampler_config()
        best_bits, _ = self.sampler.simulated_annealing(
            self._get_bit_coupling(), bias=self._get_bit_bias(), T_initial=10.0 * self.config.temperature,
            T_final=0.01 * self.config.temperature, n_steps=n_steps, chromatic=self._chromatic(),
        )
        ground_state = self._bits_to_spins(best_bits)
        return ground_state, self.energy(ground_state)


class IsingChain(IsingModel):
    """1-D nearest-neighbour chain (ising.py:265-304).  The reference fills a dense n x n matrix with the n - 1 bonds;
    here the matrix exists only if somebody reads or edits `.J` - sampling, annealing and energies use the two
    off-diagonals directly (chromatic CSR kernel: the chain is two-colourable), so n_spins can be 10^5."""

    def __init__(self, n_spins: int, J: float = 1.0, config: Optional[IsingConfig] = None, **kw):
        self._chain_J = float(J)
        self._lazy = True
        super().__init__(n_spins, config, **kw)

    @property
    def J(self) -> np.ndarray:
        if self._J is None:
            n = self.n_spins
            Jm = np.zeros((n, n))
            i = np.arange(n - 1)
            Jm[i, i + 1] = self._chain_J
            Jm[i + 1, i] = self._chain_J
            self._J = Jm
        return self._J

    @J.setter
    def J(self, value):
        self._J = np.asarray(value, dtype=np.float64)

    def _bit_csr(self):
        """(CSR of J_bit = 4 J, h_bit) of the untouched chain (ising.py:127-148 without the dense matrix)"""
        n, Jc = self.n_spins, self._chain_J
        deg = np.full(n, 2.0)
        if n > 0:
            deg[0] -= 1
            deg[-1] -= 1
        if n == 1:
            deg[:] = 0
        rowsum = Jc * deg
        rowptr = np.concatenate([[0], np.cumsum(deg.astype(np.int64))]).astype(np.int32)
        col = np.empty(int(rowptr[-1]), dtype=np.int32)
        k = 0
        for i in range(n):   # ascending columns: i - 1 then i + 1
            if i > 0:
                col[k] = i - 1
                k += 1
            if i < n - 1:
                col[k] = i + 1
                k += 1
        val = np.full(col.size, 4.0 * Jc)
        bias = (-2 * self.h + 2 * rowsum) if self.compat_reference_bias else (2 * self.h - 2 * rowsum)
        return (rowptr, col, val, n), bias

    def sample(self, n_samples: int = 1000, initial_state: Optional[np.ndarray] = None) -> np.ndarray:
        if self._J is not None:   # somebody edited the couplings: general path
            return super().sample(n_samples, initial_state)
        self._sync_sampler_config()
        csr, bias = self._bit_csr()
        initial_bits = self._spins_to_bits(initial_state) if initial_state is not None else None
        bits = self.sampler.sample_boltzmann(csr, bias=bias, n_samples=n_samples, initial_state=initial_bits, chromatic=True)
        return self._bits_to_spins(bits)

    def find_ground_state(self, n_steps: int = 1000) -> Tuple[np.ndarray, float]:
        if self._J is not None:
            return super().find_ground_state(n_steps)
        self._sync_sampler_config()
        csr, bias = self._bit_csr()
        best_bits, _ = self.sampler.simulated_annealing(csr, bias=bias, T_initial=10.0 * self.config.temperature,
                                                        T_final=0.01 * self.config.temperature, n_steps=n_steps,
                                                        chromatic=True)
        gs = self._bits_to_spins(best_bits)
        return gs, self.energy(gs)

    def energy(self, state: np.ndarray) -> float:
        if self._J is not None:
            return super().energy(state)
        s = np.asarray(state, dtype=np.float64)
        return float(-self._chain_J * np.dot(s[:-1], s[1:]) - self.h.dot(s))

    def _energies(self, samples: np.ndarray) -> np.ndarray:
        if self._J is not None:
            return super()._energies(samples)
        s = np.asarray(samples, dtype=np.float64)
        return -self._chain_J * np.sum(s[:, :-1] * s[:, 1:], axis=1) - s @ self.h


def _grid_coupling_matrix(rows: int, cols: int, J: float, periodic: bool) -> np.ndarray:
    """dense J exactly as ising.py:343-361 wires it (assignment semantics for coinciding wrap bonds)"""
    n = rows * cols
    Jm = np.zeros((n, n))
    idx = np.arange(n).reshape(rows, cols)
    Jm[idx[:, :-1].ravel(), idx[:, 1:].ravel()] = J
    Jm[idx[:, 1:].ravel(), idx[:, :-1].ravel()] = J
    Jm[idx[:-1, :].ravel(), idx[1:, :].ravel()] = J
    Jm[idx[1:, :].ravel(), idx[:-1, :].ravel()] = J
    if periodic:
        Jm[idx[:, -1], idx[:, 0]] = J
        Jm[idx[:, 0], idx[:, -1]] = J
        Jm[idx[-1, :], idx[0, :]] = J
        Jm[idx[0, :], idx[-1, :]] = J
    return Jm


class IsingGrid(IsingModel):
    """2-D square lattice (ising.py:307-421) on the bit-packed checkerboard kernel.

    `periodic=False` default as in the reference (ising.py:325).  Periodic lattices need even rows and
    cols (bipartite colouring).
    """

    def __init__(self, size: Tuple[int, int], J: float = 1.0, config: Optional[IsingConfig] = None,
                 periodic: bool = False, *, compat_reference_bias: Optional[bool] = None,
                 seed: Optional[int] = None):
        self.rows, self.cols = size
        self.periodic = periodic
        self.coupling = float(J)
        n_spins = self.rows * self.cols
        self.n_spins = n_spins
        self.config = config or IsingConfig()
        self.compat_reference_bias = _resolve_compat_bias(compat_reference_bias)
        self._J = None  # dense matrix built on demand only
        self.h = np.ones(n_spins) * self.config.external_field
        self._seed = int(seed) if seed is not None else int(np.random.randint(0, 2**31 - 1))
        self.sampler = GibbsSampler(
            GibbsConfig(temperature=self.config.temperature, n_burnin=self.config.n_burnin,
                        n_sweeps=self.config.n_sweeps),
            seed=self._seed,
        )
        self._calls = 0
        if periodic and (self.rows == 1 or self.cols == 1):
            raise ValueError("a periodic dimension of size 1 would couple a spin to itself")

    @property
    def J(self) -> np.ndarray:
        if self._J is None:
            self._J = _grid_coupling_matrix(self.rows, self.cols, self.coupling, self.periodic)
        return self._J

    @J.setter
    def J(self, value):
        self._J = np.asarray(value, dtype=np.float64)

    def _uniform_field(self) -> Optional[float]:
        h0 = float(self.h[0]) if self.n_spins else 0.0
        return h0 if np.all(self.h == h0) else None

    def _lattice_ok(self) -> bool:
        """the stencil kernel covers uniform J (untouched wiring), a uniform field and bipartite lattices: a periodic
        dimension of odd length closes odd cycles, the two-colour update does not apply and the grid takes the
        dense-J path like any other IsingModel (the reference wires any size, ising.py:343-361)"""
        if self.periodic and (self.rows % 2 or self.cols % 2):
            return False
        return self._J is None and self._uniform_field() is not None

    def sample(self, n_samples: int = 1000, initial_state: Optional[np.ndarray] = None, *, n_replicas: int = 1,
               as_tensor: bool = False):
        """ising.py:150-181 semantics: burn-in n_burnin sweeps, then one sample every n_sweeps sweeps.

        Runs on the lattice kernel when the couplings are the unmodified nearest-neighbour wiring and the
        field is uniform; otherwise falls through to the dense-J path of IsingModel.sample.
        n_replicas > 1 advances that many independent lattices together and returns
        (n_replicas, n_samples, n_spins)."""
        if not self._lattice_ok():
            if n_replicas != 1:
                raise ValueError("n_replicas > 1 needs the uniform nearest-neighbour lattice")
            return super().sample(n_samples, initial_state)
        import torch

        cfg = self.config
        self._calls += 1
        eng = Ising2DEngine(
            self.rows, self.cols, n_replicas=n_replicas, coupling=self.coupling, field=self._uniform_field(),
            temperature=cfg.temperature, periodic=self.periodic, seed=self._seed + 7919 * self._calls,
            bias_mode="reference" if self.compat_reference_bias else "physical",
        )
        if initial_state is not None:
            s = np.asarray(initial_state).reshape(-1, self.rows, self.cols)
            eng.set_spins(np.broadcast_to(s, (n_replicas, self.rows, self.cols)))
        else:
            eng.init_random()
        eng.sweep(cfg.n_burnin)
        out = torch.empty((n_samples, n_replicas, self.rows, self.cols), dtype=torch.int8, device=eng.device)
        for k in range(n_samples):
            eng.sweep(cfg.n_sweeps)
            out[k] = eng.spins_tensor(pm1=True)
        self.sampler.sample_count += n_samples
        if as_tensor:
            return out
        res = out.cpu().numpy().astype(int).reshape(n_samples, n_replicas, self.n_spins)
        if n_replicas == 1:
            return res[:, 0, :]
        return np.ascontiguousarray(res.transpose(1, 0, 2))

    def energy(self, state: np.ndarray) -> float:
        """ising.py:98-117 without the dense matrix when the wiring is untouched"""
        if self._J is not None:
            return super().energy(state)
        s = np.asarray(state).reshape(self.rows, self.cols).astype(np.float64)
        bonds = (s[:, :-1] * s[:, 1:]).sum() + (s[:-1, :] * s[1:, :]).sum()
        if self.periodic and self.cols > 2:
            bonds += (s[:, -1] * s[:, 0]).sum()
        if self.periodic and self.rows > 2:
            bonds += (s[-1, :] * s[0, :]).sum()
        return float(-self.coupling * bonds - self.h.dot(s.ravel()))

    def _energies(self, samples: np.ndarray) -> np.ndarray:
        if self._J is not None:
            return super()._energies(samples)
        s = np.asarray(samples, dtype=np.float64).reshape(-1, self.rows, self.cols)
        bonds = (s[:, :, :-1] * s[:, :, 1:]).sum((1, 2)) + (s[:, :-1, :] * s[:, 1:, :]).sum((1, 2))
        if self.periodic and self.cols > 2:
            bonds += (s[:, :, -1] * s[:, :, 0]).sum(1)
        if self.periodic and self.rows > 2:
            bonds += (s[:, -1, :] * s[:, 0, :]).sum(1)
        return -self.coupling * bonds - s.reshape(len(s), -1) @ self.h

    def _flat_to_grid(self, flat_state: np.ndarray) -> np.ndarray:
        return np.asarray(flat_state).reshape(self.rows, self.cols)

    def _grid_to_flat(self, grid_state: np.ndarray) -> np.ndarray:
        return np.asarray(grid_state).flatten()

    def compute_domains(self, state: np.ndarray) -> int:
        """ising.py:403-421"""
        state = np.asarray(state)
        if state.ndim == 1:
            state = self._flat_to_grid(state)
        horizontal_boundaries = np.sum(state[:, :-1] != state[:, 1:])
        vertical_boundaries = np.sum(state[:-1, :] != state[1:, :])
        return (horizontal_boundaries + vertical_boundaries) // 2 + 1


class IsingModel2D:
    """README.md:116-131: `IsingModel2D(size=50, coupling=1.0, temperature=2.5)` with a persistent lattice.

        ising = IsingModel2D(size=50, coupling=1.0, temperature=2.5)
        for _ in range(1000): ising.gibbs_update()
        m, e = ising.magnetization(), ising.energy()
        ms = [ising.equilibrate(T).magnetization() for T in temps]

    The lattice lives bit-packed in HBM; gibbs_update() is one checkerboard sweep.  The README is silent on
    boundary conditions: `periodic=True` by default (the textbook model with T_c = 2.269), `periodic=False`
    gives IsingGrid's default geometry.
    """

    def __init__(self, size=50, coupling: float = 1.0, temperature: float = 2.5, *, field: float = 0.0,
                 periodic: bool = True, seed: Optional[int] = None, n_burnin: int = 100, n_replicas: int = 1,
                 initial_state: Optional[np.ndarray] = None):
        if temperature <= 0:
            raise ValueError("Temperature must be positive")
        rows, cols = (size, size) if np.isscalar(size) else size
        self.rows, self.cols = int(rows), int(cols)
        self.size = size
        self.coupling = float(coupling)
        self.temperature = float(temperature)
        self.n_burnin = int(n_burnin)
        seed = int(seed) if seed is not None else int(np.random.randint(0, 2**31 - 1))
        self.engine = Ising2DEngine(self.rows, self.cols, n_replicas=n_replicas, coupling=coupling, field=field,
                                    temperature=temperature, periodic=periodic, seed=seed)
        if initial_state is not None:
            self.engine.set_spins(initial_state)
        else:
            self.engine.init_random()
        self._obs_at = None   # sweep index the cached observables belong to
        self._obs = None

    def gibbs_update(self, n_sweeps: int = 1):
        """one full heat-bath sweep (black then white sublattice)"""
        self.engine.sweep(n_sweeps)
        return self

    def _observables(self):
        """(# up spins, # anti-aligned bonds) of the current lattice: ONE kernel launch and one device-to-host copy
        serve magnetization() and energy() until the lattice is updated again"""
        if self._obs_at != self.engine.sweep_index:
            self._obs = self.engine.observables_tensor().cpu().numpy().astype(np.float64)
            self._obs_at = self.engine.sweep_index
        return self._obs

    def equilibrate(self, temperature: Optional[float] = None, n_sweeps: Optional[int] = None):
        """set the temperature and run n_burnin (default 100, IsingConfig.n_burnin) sweeps; chainable"""
        if temperature is not None:
            if temperature <= 0:
                raise ValueError("Temperature must be positive")
            self.temperature = float(temperature)
            self.engine.set_temperature(self.temperature)
        self.engine.sweep(self.n_burnin if n_sweeps is None else int(n_sweeps))
        return self

    def magnetization(self):
        """signed magnetisation per spin of the current lattice (array if n_replicas > 1)"""
        eng = self.engine
        m = (2.0 * self._observables()[:, 0] - eng.n_sites) / eng.n_sites
        return float(m[0]) if m.size == 1 else m

    def energy(self):
        """total energy of the current lattice"""
        eng, obs = self.engine, self._observables()
        e = -eng.coupling * (eng.n_bonds - 2.0 * obs[:, 1]) - eng.field * (2.0 * obs[:, 0] - eng.n_sites)
        return float(e[0]) if e.size == 1 else e

    @property
    def spins(self) -> np.ndarray:
        """current configuration in {-1,+1}, shape (rows, cols) (or (n_replicas, rows, cols))"""
        s = self.engine.get_spins(pm1=True)
        return s[0] if s.shape[0] == 1 else s

    state = spins


def demonstrate_phase_transition(sizes: List[int] = [8, 16, 32], temperatures: Optional[np.ndarray] = None,
                                 n_samples: int = 500, verbose: bool = True, seed: Optional[int] = None) -> dict:
    """ising.py:424-476: |M|, chi and C versus T for several lattice sizes (open boundaries, burn-in 200,
    10 sweeps between samples, 500 samples).  All temperatures of one size advance together as replicas of
    one batched lattice engine."""
    import torch

    if temperatures is None:
        temperatures = np.linspace(0.5, 4.0, 15)
    temperatures = np.asarray(temperatures, dtype=np.float64)
    seed = int(seed) if seed is not None else int(np.random.randint(0, 2**31 - 1))
    results = {}
    for size in sizes:
        nT = len(temperatures)
        eng = Ising2DEngine(size, size, n_replicas=nT, coupling=1.0, temperature=temperatures, periodic=False,
                            seed=seed + size)
        eng.init_random()
        eng.sweep(200)
        ms = torch.empty((n_samples, nT), dtype=torch.float64, device=eng.device)
        es = torch.empty((n_samples, nT), dtype=torch.float64, device=eng.device)
        N = size * size
        for k in range(n_samples):
            eng.sweep(10)
            obs = eng.observables_tensor().to(torch.float64)
            ms[k] = (2.0 * obs[:, 0] - N) / N
            es[k] = -(eng.n_bonds - 2.0 * obs[:, 1])
        m = ms.cpu().numpy()
        e = es.cpu().numpy()
        mags = np.abs(m.mean(0))
        chi = (np.mean(m**2, 0) - np.mean(m, 0) ** 2) * N / temperatures
        C = (np.mean(e**2, 0) - np.mean(e, 0) ** 2) / (temperatures**2 * N)
        if verbose:
            print(f"\nSimulating {size}x{size} Ising grid...")
            for T, a, b, c in zip(temperatures, mags, chi, C):
                print(f"  T={T:.2f}: |M|={a:.3f}, chi={b:.3f}, C={c:.3f}")
        results[size] = {
            "temperatures": temperatures,
            "magnetizations": mags,
            "susceptibilities": chi,
            "specific_heats": C,
        }
    return results

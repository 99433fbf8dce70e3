"""timing probes (not the contract bench): Langevin 1e6 chains, dense-J C3 shape, lattice tempering C5 shape"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tsu_emulator_b200 import (GibbsConfig, GibbsSampler, Ising2DEngine, QuadraticEnergy, ThermalSamplingUnit, TSUConfig)
from tsu_emulator_b200.distributed import LatticeTempering

def timed(fn, reps=1):
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps, out

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "langevin"):
    for dtype in ("float32", "float64"):
        tsu = ThermalSamplingUnit(TSUConfig(temperature=1.0, dt=0.01, friction=1.0, n_burnin=100, n_steps=500), seed=1, dtype=dtype)
        n = 1_000_000
        tsu.sample_from_energy(QuadraticEnergy(), np.zeros(10), 1000, as_tensor=True)
        ms, x = timed(lambda: tsu.sample_from_energy(QuadraticEnergy(), np.zeros(10), n, as_tensor=True))
        print(f"[langevin {dtype}] 1e6 chains x dim 10 x 600 steps: {ms:.1f} ms  {n*600/ms*1e3:.3e} chain-steps/s  "
              f"{n*6000/ms*1e3:.3e} coord-updates/s  var={x.var().item():.4f} (EM theory 0.5051)")
if what in ("all", "dense"):
    N, chains, sweeps = 4096, int(os.environ.get("CHAINS", 2048)), 10
    rng = np.random.default_rng(7)
    J = rng.normal(size=(N, N)) / np.sqrt(N); J = (J + J.T) / np.sqrt(2); np.fill_diagonal(J, 0)
    J = torch.from_numpy(J).to(torch.bfloat16).to(torch.float64).numpy()   # bf16-representable couplings
    for prec in ("float32",):
        smp = GibbsSampler(GibbsConfig(temperature=1.0, n_sweeps=sweeps), seed=3, precision=prec)
        smp.sample_chains(J, n_chains=8, n_sweeps=1, as_tensor=True)
        ms, (st, e) = timed(lambda: smp.sample_chains(J, n_chains=chains, n_sweeps=sweeps, as_tensor=True, return_energy=True))
        upd = chains * N * sweeps
        print(f"[dense {prec}] N={N} chains={chains} sweeps={sweeps}: {ms:.1f} ms  {upd/ms*1e3:.3e} updates/s  "
              f"tensor-roofline frac {upd/ms*1e3*2*N/1399.3e12:.4f}  E/N={e.mean().item()/N:.4f}")
if what in ("all", "pt"):
    temps = np.linspace(0.1, 5.0, 50)
    K = int(os.environ.get("LADDERS", 8))
    fac = lambda n, r0, T: Ising2DEngine(1024, 1024, n_replicas=n, temperature=T, periodic=True, seed=5, replica0=r0).init_random()
    pt = LatticeTempering(temps, n_ladders=K, engine_factory=fac, n_sweeps=10, swap_interval=10, seed=9)
    for _ in range(10): pt.step()
    ms, _ = timed(lambda: [pt.step() for _ in range(20)])
    upd = 20 * 10 * K * 50 * 1024 * 1024
    m, e = pt.observables_by_slot()
    print(f"[pt] 50 temps x {K} ladders x 1024^2, 20 iterations of 10 sweeps (+2 swap passes): {ms:.1f} ms  {upd/ms*1e3:.3e} updates/s")
    print("     |m| by slot (first ladder):", np.round(np.abs(m[0, ::7]), 3), " swap accept rate:", (pt.stats[1] / pt.stats[0]).item())

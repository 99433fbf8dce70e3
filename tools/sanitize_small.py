"""small invocations of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py [lattice|dense|all]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
what = sys.argv[1] if len(sys.argv) > 1 else "all"
from tsu_emulator_b200 import (GibbsConfig, GibbsSampler, Ising2DEngine, IsingChain, IsingConfig, QuadraticEnergy,
                               ThermalSamplingUnit, TSUConfig)
if what in ("lattice", "all"):
    for rows, cols, periodic, n_rep in [(70, 1024, True, 3), (33, 1000, False, 2), (50, 50, False, 2), (12, 70, True, 1)]:
        os.environ["TSU_LATTICE_RESIDENT"] = "0"
        from tsu_emulator_b200 import _lib
        _lib.load().tsu_ising2d_reload_tuning()
        e = Ising2DEngine(rows, cols, n_replicas=n_rep, temperature=2.269, periodic=periodic, seed=3).init_random()
        e.specialise()
        e.sweep(3)
        e.half_sweep(0, rows=(0, 1)); e.half_sweep(0, rows=(1, rows))
        print(rows, cols, periodic, e.magnetization()[:2], e.energy()[:1])
    os.environ["TSU_LATTICE_RESIDENT"] = "1"
    _lib.load().tsu_ising2d_reload_tuning()
    e = Ising2DEngine(50, 50, n_replicas=2, temperature=2.5, periodic=True, seed=3).init_random().sweep(5)
    print("resident", e.magnetization())
if what in ("dense", "all"):
    rng = np.random.default_rng(0)
    N = 40
    J = rng.normal(size=(N, N)); J = (J + J.T) / 2
    s = GibbsSampler(GibbsConfig(temperature=1.0, n_burnin=2, n_sweeps=2), seed=1)
    print("dense", s.sample_boltzmann(J, None, n_samples=2, n_chains=3).shape)
    print("anneal", s.simulated_annealing(J, None, n_steps=5)[1])
    Js = np.zeros((N, N)); i = np.arange(N - 1); Js[i, i + 1] = Js[i + 1, i] = 1.0
    print("sparse", s.sample_boltzmann(Js, None, n_samples=2, n_chains=3, chromatic=True).shape)
    print("chain", IsingChain(500, config=IsingConfig(temperature=1.5, n_burnin=3, n_sweeps=2), seed=2).sample(2).shape)
    Jt = rng.integers(-1, 2, (256, 256)).astype(float); Jt = np.triu(Jt, 1); Jt = Jt + Jt.T
    t = GibbsSampler(GibbsConfig(temperature=1.5, n_burnin=1, n_sweeps=1), seed=1, precision="bf16")
    print("tc", t.sample_boltzmann(Jt, None, n_samples=1, n_chains=70).shape)
    tsu = ThermalSamplingUnit(TSUConfig(n_burnin=3, n_steps=5), seed=1)
    print("langevin", tsu.sample_from_energy(QuadraticEnergy(), np.zeros(3), 40).shape)
    print("traced", tsu.sample_from_energy(lambda x: np.sum(x ** 4) + np.sum(np.cos(x)), np.zeros(3), 40).shape)
print("done")

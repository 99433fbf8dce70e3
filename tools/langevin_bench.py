"""C5 Langevin timing: python tools/langevin_bench.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsu_emulator_b200 import QuadraticEnergy, ThermalSamplingUnit, TSUConfig
for dtype in ("float32", "float64", "float32", "float64"):
    tsu = ThermalSamplingUnit(TSUConfig(temperature=1.0, dt=0.01, friction=1.0, n_burnin=100, n_steps=500), seed=1, dtype=dtype)
    x = tsu.sample_boltzmann(QuadraticEnergy(), n_samples=10**6, dim=10, as_tensor=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); x = tsu.sample_boltzmann(QuadraticEnergy(), n_samples=10**6, dim=10, as_tensor=True); b.record()
    torch.cuda.synchronize()
    print(f"{dtype}: {a.elapsed_time(b):.2f} ms  var={float(x.var()):.5f}")

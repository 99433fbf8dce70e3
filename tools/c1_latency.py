"""C1 (IsingModel2D size 50): per-call and batched latency of gibbs_update: python tools/c1_latency.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsu_emulator_b200 import IsingModel2D
for per in (True, False):
    m = IsingModel2D(size=50, coupling=1.0, temperature=2.5, periodic=per, seed=0)
    m.gibbs_update(10); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(1000):
        m.gibbs_update()
    mag, en = m.magnetization(), m.energy(); dt = time.perf_counter() - t0
    t0 = time.perf_counter(); m.gibbs_update(1000); mag, en = m.magnetization(), m.energy(); dt2 = time.perf_counter() - t0
    print(f"C1 periodic={per}: 1000 x gibbs_update() {dt*1e3:.2f} ms; gibbs_update(1000) {dt2*1e3:.2f} ms  (M={mag:.4f}, E={en:.1f})")

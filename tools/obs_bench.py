"""bandwidth of the observables kernel (reads the packed state once): python tools/obs_bench.py [replicas] [L]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsu_emulator_b200.lattice import Ising2DEngine
n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
eng = Ising2DEngine(L, L, n_replicas=n_rep, temperature=2.269, periodic=True, seed=1).init_random()
eng.sweep(1); eng.observables_tensor(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    eng.observables_tensor()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
gb = eng.state.numel() * 4 / 1e9
print(f"observables of {n_rep} x {L}^2 ({gb:.2f} GB): {ms:.3f} ms = {gb/ms*1e3:.0f} GB/s = {gb/ms*1e3/6553.6:.3f} of measured HBM peak")

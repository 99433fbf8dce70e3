"""throughput of the generic half-sweep kernel (open boundaries = IsingGrid default): python tools/quick_open.py [replicas] [L] [sweeps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsu_emulator_b200.lattice import Ising2DEngine
n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
n_sw = int(sys.argv[3]) if len(sys.argv) > 3 else 5
for periodic, cols in ((False, L), (True, L + 2), (False, L + 1), (True, L)):
    eng = Ising2DEngine(L, cols, n_replicas=n_rep, temperature=2.269, periodic=periodic, seed=1).init_random()
    eng.sweep(2); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.sweep(n_sw); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b); upd = n_rep * L * cols * n_sw
    print(f"periodic={periodic} {L}x{cols} x{n_rep}: {ms:.2f} ms  {upd/ms*1e3:.3e} updates/s  E/N={float(eng.energy()[0]/eng.n_sites):.4f}")

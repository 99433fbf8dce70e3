"""quick timing probe of the lattice kernel (not the contract bench): python tools/quick_lattice.py [replicas] [L] [sweeps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsu_emulator_b200.lattice import Ising2DEngine

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
n_sw = int(sys.argv[3]) if len(sys.argv) > 3 else 5
eng = Ising2DEngine(L, L, n_replicas=n_rep, temperature=2.269, periodic=True, seed=1234)
eng.init_random()
eng.sweep(2)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); eng.sweep(n_sw); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
upd = n_rep * L * L * n_sw
print(f"replicas={n_rep} L={L} sweeps={n_sw}: {ms:.2f} ms  {upd/ms*1e3:.3e} updates/s  "
      f"alg GB/s={upd*0.25/ms*1e3/1e9:.1f}  frac_hbm={upd*0.25/ms*1e3/6553.6e9:.3f}")
print("M", eng.magnetization()[:2], "E/N", (eng.energy()/eng.n_sites)[:2])

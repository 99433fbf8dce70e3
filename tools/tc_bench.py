"""C3-shaped timing of the tensor-core dense path: python tools/tc_bench.py [N] [chains] [sweeps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tsu_emulator_b200 import GibbsConfig, GibbsSampler
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
C = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
SW = int(sys.argv[3]) if len(sys.argv) > 3 else 10
rng = np.random.default_rng(7)
J = rng.normal(size=(N, N)) / np.sqrt(N); J = (J + J.T) / np.sqrt(2); np.fill_diagonal(J, 0)
smp = GibbsSampler(GibbsConfig(temperature=1.0, n_sweeps=SW), seed=3, precision="bf16")
smp.sample_chains(J, n_chains=128, n_sweeps=1, as_tensor=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time; t0 = time.time()
st, e = smp.sample_chains(J, n_chains=C, n_sweeps=SW, as_tensor=True, return_energy=True)
torch.cuda.synchronize(); wall = time.time() - t0
# device time of the sweep kernel alone
from tsu_emulator_b200 import _lib
Jd = torch.from_numpy(J).cuda().to(torch.bfloat16)
a.record()
_lib.call("tsu_dense_gibbs_tc_run", _lib.ptr(Jd), None, _lib.ptr(st), C, N, 1.0, None, SW, 3, 100, 0, None, _lib.current_stream())
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
upd = C * N * SW
print(f"[dense tcgen05] N={N} chains={C} sweeps={SW}: {ms:.2f} ms  {upd/ms*1e3:.3e} updates/s  "
      f"{upd*2*N/ms*1e3/1e12:.1f} TFLOP/s = {upd*2*N/ms*1e3/1399.3e12:.4f} of sustained bf16 peak; E/N={e.mean().item()/N:.4f} (host wall incl. upload {wall*1e3:.0f} ms)")
H = torch.empty((C, N), device="cuda")
a.record()
_lib.call("tsu_dense_tc_debug_fields", _lib.ptr(Jd), _lib.ptr(st), C, N, _lib.ptr(H), _lib.current_stream())
b.record(); torch.cuda.synchronize()
print(f"[gemm only, 1 pass over all blocks] {a.elapsed_time(b):.2f} ms")

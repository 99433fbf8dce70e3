"""C3-shaped timing of the tensor-core dense sweep: python tools/tc_bench.py [N=4096] [sweeps=10]  (2048 chains and one
128-chain tile per SM); TSU_TC_M=64/128 forces the tile height"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tsu_emulator_b200 import _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
SW = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rng = np.random.default_rng(7)
J = rng.normal(size=(N, N)) / np.sqrt(N); J = (J + J.T) / np.sqrt(2); np.fill_diagonal(J, 0)
Jd = torch.from_numpy(J).cuda().to(torch.bfloat16).contiguous()
n_sm = torch.cuda.get_device_properties(0).multi_processor_count
for C in (2048, 128 * n_sm):
    st = (torch.rand(C, N, device="cuda") < 0.5).to(torch.uint8)
    _lib.call("tsu_dense_gibbs_tc_run", _lib.ptr(Jd), None, _lib.ptr(st), C, N, 1.0, None, 1, 3, 0, 0, None, _lib.current_stream())
    best = 1e30
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.call("tsu_dense_gibbs_tc_run", _lib.ptr(Jd), None, _lib.ptr(st), C, N, 1.0, None, SW, 3, 1, 0, None, _lib.current_stream())
        b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    print(f"N={N} chains={C} sweeps={SW}: {best:.3f} ms  {C*N*SW/best*1e3:.3e} updates/s  {2.0*N*N*C*SW/best*1e3/1e12:.1f} TFLOP/s  mean bit {st.float().mean().item():.4f}")

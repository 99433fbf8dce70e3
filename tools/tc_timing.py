"""Per-role stall accounting of the tcgen05 dense kernel (needs a -DTSU_TC_TIMING build)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tsu_emulator_b200 import _lib
lib = _lib.load()
N, C = 4096, 2048
J = (torch.randn(N, N, device="cuda") / N**0.5).to(torch.bfloat16)
st = (torch.rand(C, N, device="cuda") < 0.5).to(torch.uint8)
f = lib.tsu_dense_tc_debug_timing; f.restype = ctypes.c_int; f.argtypes = [ctypes.c_void_p, ctypes.c_int]
_lib.call("tsu_dense_gibbs_tc_run", _lib.ptr(J), None, _lib.ptr(st), C, N, 1.0, None, 1, 3, 0, 0, None, _lib.current_stream())
torch.cuda.synchronize(); f(None, 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); _lib.call("tsu_dense_gibbs_tc_run", _lib.ptr(J), None, _lib.ptr(st), C, N, 1.0, None, 2, 3, 1, 0, None, _lib.current_stream()); b.record()
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 32)(); f(out, 0)
ms = a.elapsed_time(b); total_clk = ms * 1e-3 * 1.965e9
names = ["issuer0 wait full", "issuer1 wait full", "issuers wait acc_free", "P0 wait empty", "P0 wait state_ready", "P0 load+expand+arrive",
         "-", "P1 wait empty", "P1 wait state_ready", "P1 load+expand+arrive", "-", "EPI wait acc_full", "EPI ld+update", "-", "-", "-",
         "P0w0 load_b issue", "P0w0 lds+alu+sttm issue", "P0w0 wait::st", "P0w0 cp.async.wait", "P0w0 fence+syncwarp+arrive", "issuer0 membar+fence", "issuer0 mma+commit+syncwarp"]
nchunks = 2 * 128 * 32
print(f"2 sweeps: {ms:.2f} ms = {total_clk:.3e} clk; chunks per CTA = {nchunks}; {total_clk/nchunks:.0f} clk/chunk")
for i, n in enumerate(names):
    print(f"  {n:22s} {out[i]:14d} clk  {100*out[i]/total_clk:6.1f}% of kernel   {out[i]/nchunks:8.1f} clk/chunk")

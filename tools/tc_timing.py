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
names = {0: "issuer wait full (per chunk)", 2: "issuer wait acc_free", 3: "P0 wait empty (per own chunk = 1/2)", 7: "P1 wait empty",
         4: "P0w0 wait panel_done", 16: "P0w0 J tile cp.async issue", 17: "P0w0 lds+alu+sttm issue", 18: "P0w0 wait::st",
         19: "P0w0 cp.async.wait_group 0", 20: "P0w0 fence+syncwarp+arrive", 21: "issuer membar+fence", 22: "issuer mma+commit+syncwarp",
         23: "corr issuer wait delta_ready", 24: "corr issuer fence+mma+commit", 11: "EPI wait acc_full (per panel)",
         13: "EPI wait corr_done (3 per panel)", 12: "EPI ld+update+delta (4 per panel)",
         25: "EPI (unused)", 26: "EPI thr wait+load(+prepass M128)", 27: "EPI waits", 28: "EPI ldtm", 29: "EPI scale+chain", 30: "EPI delta sttm+arrive"}
nchunks = 2 * 32 * 32
print(f"2 sweeps: {ms:.2f} ms = {total_clk:.3e} clk; chunks per CTA = {nchunks}; {total_clk/nchunks:.0f} clk/chunk")
for i, n in sorted(names.items()):
    print(f"  {n:42s} {out[i]:14d} clk  {100*out[i]/total_clk:6.1f}% of kernel   {out[i]/nchunks:8.1f} clk/chunk  {out[i]/(nchunks/32):9.0f} clk/panel")

"""timing experiments on the GEMM stage of the tcgen05 dense kernel: TSU_TC_DEBUG=flags python tools/tc_dbg.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsu_emulator_b200 import _lib
N, C = 4096, 2048
J = (torch.randn(N, N, device="cuda") / N**0.5).to(torch.bfloat16)
st = (torch.rand(C, N, device="cuda") < 0.5).to(torch.uint8)
for it in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); _lib.call("tsu_dense_gibbs_tc_run", _lib.ptr(J), None, _lib.ptr(st), C, N, 1.0, None, 4, 3, 0, 0, None, _lib.current_stream()); b.record()
    torch.cuda.synchronize()
ms = a.elapsed_time(b) / 4
print(f"TSU_TC_DEBUG={os.environ.get('TSU_TC_DEBUG','0')}: sweep {ms:.3f} ms = {ms*1e-3*1.965e9/4096:.0f} clk per N=32 chunk-equivalent = {ms*1e-3*1.965e9/1024:.0f} clk per panel-chunk")

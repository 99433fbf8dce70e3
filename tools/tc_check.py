"""validate the tcgen05 local-field GEMM against torch: python tools/tc_check.py [N] [chains]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tsu_emulator_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256  # multiple of 128
C = int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.manual_seed(0)
J = (torch.randn(N, N, device="cuda") / N**0.5).to(torch.bfloat16)
S = (torch.rand(C, N, device="cuda") < 0.5).to(torch.uint8)
H = torch.full((C, N), float("nan"), device="cuda", dtype=torch.float32)
_lib.call("tsu_dense_tc_debug_fields", _lib.ptr(J), _lib.ptr(S), C, N, _lib.ptr(H), _lib.current_stream())
torch.cuda.synchronize()
ref = S.double() @ J.double().T
err = (H.double() - ref).abs().max().item()
print(f"N={N} chains={C}: max abs err {err:.3e}  (|ref| max {ref.abs().max().item():.3f})  nan={torch.isnan(H).sum().item()}")
if err > 1e-3:
    bad = ((H.double() - ref).abs() > 1e-3).nonzero()
    print("first mismatches (chain, site):", bad[:8].tolist())
    print("H[0,:8]  ", H[0, :8].tolist()); print("ref[0,:8]", ref[0, :8].tolist())
    sys.exit(1)
print("tcgen05 fields OK")

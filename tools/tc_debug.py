import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tsu_emulator_b200 import _lib
from oracle import dense_oracle as D
N, C, seed, T = 128, 128, 99, 1.0
rng = np.random.default_rng(4)
J = rng.integers(-2, 3, (N, N)).astype(np.float64); J = np.triu(J, 1); J = J + J.T
init = rng.integers(0, 2, (C, N))
Jd = torch.from_numpy(J).cuda().to(torch.bfloat16).contiguous()
st = torch.from_numpy(init.astype(np.uint8)).cuda()
F = torch.zeros((C, N), device="cuda")
_lib.call("tsu_dense_gibbs_tc_run", _lib.ptr(Jd), None, _lib.ptr(st), C, N, T, None, 1, seed, 0, 0, _lib.ptr(F), _lib.current_stream())
torch.cuda.synchronize()
out = st.cpu().numpy().astype(int); F = F.cpu().numpy()
for c in (0, 1, 77):
    U = D.philox_uniforms_tc(seed, c, 0, N)
    # oracle with field trace
    s = init[c].copy(); fields = np.zeros(N); blockfields = np.zeros(N)
    for i in range(N):
        if i % 32 == 0:
            blockfields[i:i+32] = J[i:i+32] @ s
        h = J[i] @ s; fields[i] = h
        p = D.sigmoid_ref(h / T); s[i] = 1 if U[i] < p else 0
    diff = np.flatnonzero(out[c] != s)
    print(f"chain {c}: first diffs {diff[:10]}  n={diff.size}")
    fd = np.flatnonzero(np.abs(F[c] - fields) > 1e-3)
    print(f"   block-start fields differ at {fd[:10]} n={fd.size}; F[:8]={F[c,:8]} want {fields[:8]} blockstart {blockfields[:8]}")
    print("   u[:4]", U[:4], " p[:4]", [D.sigmoid_ref(fields[i]) for i in range(4)], " out", out[c,:8], " want", s[:8], "init", init[c,:8])

"""Offline build of the run-time specialised lattice kernel (what tsu_ising2d_jit_prepare hands to NVRTC), with nvcc:
registers / spills and the SASS opcode mix without a GPU.
    python tools/jit_offline.py [T=2.269] [W=4] [MINB=4] [extra -D...]
writes /tmp/tsu_jit_W<W>_B<MINB>.{cu,cubin,sass}"""
import os, subprocess, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsu_emulator_b200.lattice import build_lut

T = float(sys.argv[1]) if len(sys.argv) > 1 else 2.269
W = int(sys.argv[2]) if len(sys.argv) > 2 else 4
MINB = int(sys.argv[3]) if len(sys.argv) > 3 else 4
extra = sys.argv[4:]
lut = [int(x) for x in build_lut(1.0, 0.0, T)]
tab = [sum(((lut[20 + u] >> (31 - k)) & 1) << u for u in range(5)) for k in range(8)]
fz = sum(1 << u for u in range(5) if (lut[20 + u] & 0xFFFFFF) == 0 and not ((lut[25] >> (20 + u)) & 1))
src = "".join(f"#define TSU_FT{k} {tab[k]}\n" for k in range(8))
src += f"#define TSU_FZ {fz}\n#define TSU_ALWAYS {(lut[25] >> 20) & 31}u\n#define TSU_JIT_MINB {MINB}\n#define TSU_JIT_W {W}\n"
src += ('#include "ising2d_fast.cuh"\n'
        'extern "C" __global__ void __launch_bounds__(128, TSU_JIT_MINB) tsu_jit_half_sweep(tsu_fast::SweepParams P) {\n'
        '  tsu_fast::half_sweep_fast_body<TSU_JIT_W>(P);\n}\n')
base = f"/tmp/tsu_jit_W{W}_B{MINB}"
open(base + ".cu", "w").write(src)
csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tsu_emulator_b200", "csrc")
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-I" + csrc, "-cubin",
       "-Xptxas", "-v", "-o", base + ".cubin", base + ".cu"] + extra
out = subprocess.run(cmd, capture_output=True, text=True)
print("\n".join(l for l in out.stderr.splitlines() if "registers" in l or "spill" in l or "error" in l))
sass = subprocess.run(["cuobjdump", "-sass", base + ".cubin"], capture_output=True, text=True).stdout
open(base + ".sass", "w").write(sass)
ops = collections.Counter()
n = 0
for line in sass.splitlines():
    parts = line.split()
    if len(parts) > 2 and parts[0].startswith("/*") and len(parts[0]) == 8:
        op = parts[1] if not parts[1].startswith("@") else parts[2]
        ops[op.rstrip(";").split(".")[0]] += 1
        n += 1
print("static instructions:", n)
print(", ".join(f"{k} {v}" for k, v in ops.most_common(24)))

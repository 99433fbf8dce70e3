"""one GPU: half-sweep time against rows at 131072 columns - the intercept is the fixed cost per launch
(python tools/slab_fit.py [strip])"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tsu_emulator_b200 import _lib
from tsu_emulator_b200.lattice import Ising2DEngine

COLS, SWEEPS = 131072, 20
if len(sys.argv) > 1:
    os.environ["TSU_LATTICE_STRIP"] = sys.argv[1]
    _lib.load().tsu_ising2d_reload_tuning()
xs, ys = [], []
for rows in (512, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
    eng = Ising2DEngine(rows, COLS, n_replicas=1, temperature=2.269, periodic=True, seed=7)
    eng.specialise()
    eng.init_random()
    eng.sweep(3)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.sweep(SWEEPS); b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / SWEEPS / 2 * 1e3
    xs.append(rows); ys.append(us)
    print(f"rows={rows}: {us:.2f} us per half-sweep  {rows * COLS / 2 / us * 1e6:.3e} updates/s", flush=True)
    del eng
slope, icept = np.polyfit(xs[3:], ys[3:], 1)
print(f"fit (rows >= 4096): {icept:.2f} us + {slope * 1e3:.4f} us per 1000 rows  -> asymptotic {COLS / 2 / slope * 1e6:.3e} updates/s")

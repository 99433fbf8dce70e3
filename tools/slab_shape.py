"""one GPU, one slab of the C4 lattice: sweep time against rows per slab and strip length
(python tools/slab_shape.py) - tells kernel-shape losses of the row-slab split from exchange losses"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsu_emulator_b200 import _lib
from tsu_emulator_b200.lattice import Ising2DEngine

COLS, SWEEPS = 131072, 10
for rows in (131072, 16384):
    for strip in (0, 16, 32, 64, 89, 128):
        if strip:
            os.environ["TSU_LATTICE_STRIP"] = str(strip)
        else:
            os.environ.pop("TSU_LATTICE_STRIP", None)
        _lib.load().tsu_ising2d_reload_tuning()
        eng = Ising2DEngine(rows, COLS, n_replicas=1, temperature=2.269, periodic=True, seed=7)
        eng.specialise()
        eng.init_random()
        eng.sweep(3)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.sweep(SWEEPS); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / SWEEPS
        print(f"rows={rows} strip={strip or 'auto'}: {ms:.4f} ms/sweep  {rows * COLS / ms * 1e3:.3e} updates/s", flush=True)
        del eng

"""one GPU, one slab of the C4 lattice without neighbours: time per sweep of the slab driver against the number of
interior row ranges (TSU_LATTICE_SPLIT) - python tools/slab_split.py [rows]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsu_emulator_b200 import _lib
from tsu_emulator_b200.lattice import Ising2DEngine

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
COLS, SWEEPS = 131072, 10
eng = Ising2DEngine(rows, COLS, n_replicas=1, temperature=2.269, periodic=True, seed=7)
eng.wrap_rows = False
eng.specialise()
eng.init_random()
wpr = eng.state.shape[-1]
halo = torch.zeros(4 * wpr + 16, dtype=torch.int32, device="cuda")
flags = halo[4 * wpr:]
side = torch.cuda.Stream()


def run(n):
    _lib.call("tsu_ising2d_slab_sweeps_p2p", int(eng._jit), _lib.ptr(eng.state), 1, rows, COLS, 1, _lib.ptr(eng.lut), None,
              eng.seed, eng.sweep_index, n, 0, 0, _lib.ptr(halo), _lib.ptr(flags), None, None, None, None, 0, 0,
              int(torch.cuda.current_stream().cuda_stream), int(side.cuda_stream))
    eng.sweep_index += n


for split, strip in [(1, 0), (4, 0), (8, 0), (-1, 0), (8, 32)]:
    os.environ["TSU_LATTICE_SPLIT"] = str(split)
    if strip:
        os.environ["TSU_LATTICE_STRIP"] = str(strip)
    else:
        os.environ.pop("TSU_LATTICE_STRIP", None)
    _lib.load().tsu_ising2d_reload_tuning()
    run(3)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(SWEEPS); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / SWEEPS
    print(f"rows={rows} split={split} strip={strip or 'auto'}: {ms:.4f} ms/sweep  {rows * COLS / ms * 1e3:.3e} updates/s", flush=True)

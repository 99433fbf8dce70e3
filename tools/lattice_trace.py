"""Where the time of one half-sweep launch goes: start / end / SM of every CTA of the last launches.
Needs a library built with -DTSU_LATTICE_TRACE (TSU_B200_LIB=<path>); traces the prebuilt kernel (TSU_B200_NO_JIT=1).
python tools/lattice_trace.py [rows] [cols]"""
import sys, os, ctypes
os.environ.setdefault("TSU_B200_NO_JIT", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tsu_emulator_b200 import _lib
from tsu_emulator_b200.lattice import Ising2DEngine

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
lib = _lib.load()
f = lib.tsu_debug_lattice_trace
f.restype = ctypes.c_int
f.argtypes = [ctypes.c_void_p]
eng = Ising2DEngine(rows, cols, n_replicas=1, temperature=2.269, periodic=True, seed=7)
eng.init_random()
eng.sweep(4)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); eng.sweep(10); b.record(); torch.cuda.synchronize()
print(f"{rows} x {cols}: {a.elapsed_time(b) / 20 * 1e3:.2f} us per half-sweep (traced build, prebuilt kernel)")
buf = np.zeros((4, 8192, 3), dtype=np.uint64)
assert f(buf.ctypes.data) == 0
n_cta = int((buf[0, :, 1] != 0).sum())
tr = buf[:, :n_cta].astype(np.int64)
order = np.argsort(tr[:, :, 0].min(axis=1))  # launches in time order
t00 = tr[order[0], :, 0].min()
prev_end = None
for k in order:
    st, en, sm = tr[k, :, 0] - t00, tr[k, :, 1] - t00, tr[k, :, 2]
    life = en - st
    line = f"launch slot {k}: {n_cta} CTAs on {len(np.unique(sm))} SMs; first start {st.min() / 1e3:8.2f} us, last start {st.max() / 1e3:8.2f}, first end {en.min() / 1e3:8.2f}, last end {en.max() / 1e3:8.2f}"
    if prev_end is not None:
        line += f"; gap after previous launch {(st.min() - prev_end) / 1e3:.2f} us"
    prev_end = en.max()
    print(line)
    first_wave = np.sort(st)[: min(n_cta, 592)]
    print(f"   first-wave starts spread {(first_wave.max() - first_wave.min()) / 1e3:.2f} us; CTA life: first 592 median {np.median(life[np.argsort(st)[:592]]) / 1e3:.2f} us, "
          f"middle median {np.median(life[np.argsort(st)[n_cta // 3: 2 * n_cta // 3]]) / 1e3:.2f}, last 592 median {np.median(life[np.argsort(st)[-592:]]) / 1e3:.2f}")
    # per-SM finishing times and work
    last_by_sm = np.array([en[sm == s].max() for s in np.unique(sm)])
    cnt_by_sm = np.array([(sm == s).sum() for s in np.unique(sm)])
    print(f"   per SM: CTAs {cnt_by_sm.min()}..{cnt_by_sm.max()}, last end {last_by_sm.min() / 1e3:.2f}..{last_by_sm.max() / 1e3:.2f} us "
          f"(mean {last_by_sm.mean() / 1e3:.2f}); kernel span {(en.max() - st.min()) / 1e3:.2f} us")
    # resident CTAs over time (10 samples over the span)
    ts = np.linspace(st.min(), en.max(), 21)
    res = [int(((st <= t) & (en > t)).sum()) for t in ts]
    print("   resident CTAs at 5% steps:", res)

"""Secondary configurations of BASELINE.json (C1, C3, C4 on one GPU, C5) on one B200: one JSON line each.
The contract benchmark (C2) is bench.py.  python tools/bench_all.py [--cpu]  (--cpu also times the literal CPU port of C1)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tsu_emulator_b200 import (GibbsConfig, GibbsSampler, Ising2DEngine, IsingModel2D, QuadraticEnergy,
                               ThermalSamplingUnit, TSUConfig, _lib)
from tsu_emulator_b200.distributed import LatticeTempering, SlabShardedIsing2D

def ev():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

out = []
# ---- C1: IsingModel2D size=50, 1000 gibbs_update sweeps + magnetization/energy -------------------------------
for periodic in (True, False):
    m = IsingModel2D(size=50, coupling=1.0, temperature=2.5, periodic=periodic, seed=0)
    m.gibbs_update(10); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(1000):
        m.gibbs_update()
    mag, en = m.magnetization(), m.energy()
    dt = time.perf_counter() - t0
    out.append({"config": "C1", "workload": f"IsingModel2D(50, 1.0, 2.5) periodic={periodic}: 1000 gibbs_update + magnetization + energy",
                "wall_s": dt, "spin_updates_per_s": 2.5e6 / dt, "magnetization": mag, "energy": en,
                "note": "launch-latency bound: 1000 launches (one resident-kernel launch per call) of a 2500-spin lattice"})
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m.gibbs_update(1000)
    mag, en = m.magnetization(), m.energy()
    dt = time.perf_counter() - t0
    out.append({"config": "C1-batched", "workload": f"IsingModel2D(50, 1.0, 2.5) periodic={periodic}: gibbs_update(1000) + magnetization + energy",
                "wall_s": dt, "spin_updates_per_s": 2.5e6 / dt, "note": "one launch: the lattice stays with one thread block for all 1000 sweeps"})
if "--cpu" in sys.argv:
    from oracle import ising2d_oracle as O
    Jb, hb = O.dense_bit_model(50, 50, 1.0, 0.0, False)
    rng = np.random.default_rng(0); st = rng.integers(0, 2, 2500); order = np.arange(2500)
    t0 = time.perf_counter()
    for _ in range(100):
        st = O.gibbs_sweep_port(st, Jb, hb, 2.5, order, rng.random(2500))
    dt = (time.perf_counter() - t0) * 10
    out.append({"config": "C1-cpu", "workload": "literal port of tsu/gibbs.py:128-162, IsingGrid((50,50)) dense J, 1 core, 100 sweeps x10",
                "wall_s_1000_sweeps": dt, "spin_updates_per_s": 2.5e6 / dt})
# ---- C3: dense SK N=4096, 2048 chains, 10 sweeps on the tensor cores ---------------------------------------
N, SW = 4096, 10
rng = np.random.default_rng(7)
J = rng.normal(size=(N, N)) / np.sqrt(N); J = (J + J.T) / np.sqrt(2); np.fill_diagonal(J, 0)
Jd = torch.from_numpy(J).cuda().to(torch.bfloat16).contiguous()
n_sm = torch.cuda.get_device_properties(0).multi_processor_count
for C, tag in ((2048, "C3"), (128 * n_sm, "C3-full-wave (same N, one 128-chain tile per SM)")):
    st = (torch.rand(C, N, device="cuda") < 0.5).to(torch.uint8)
    _lib.call("tsu_dense_gibbs_tc_run", _lib.ptr(Jd), None, _lib.ptr(st), C, N, 1.0, None, 1, 3, 0, 0, None, _lib.current_stream())
    a, b = ev(); a.record()
    _lib.call("tsu_dense_gibbs_tc_run", _lib.ptr(Jd), None, _lib.ptr(st), C, N, 1.0, None, SW, 3, 1, 0, None, _lib.current_stream())
    b.record(); torch.cuda.synchronize(); ms = a.elapsed_time(b)
    upd = C * N * SW
    m = 64 if (C + 127) // 128 < n_sm else 128      # chains per CTA, as chosen in csrc/dense_tc.cu
    out.append({"config": tag, "workload": f"GibbsSampler dense SK J (bf16) N={N}, {C} chains, {SW} sequential sweeps, tcgen05",
                "ms": ms, "spin_updates_per_s": upd / ms * 1e3, "tflops": upd * 2 * N / ms * 1e3 / 1e12,
                "tensor_roofline_frac_of_1399.3": upd * 2 * N / ms * 1e3 / 1399.3e12, "chains_per_cta": m, "ctas": (C + m - 1) // m})
    del st
# ---- C4 (one GPU): 131072 x 131072 single lattice --------------------------------------------------------
rows = cols = 131072
fac = lambda lr, r0: Ising2DEngine(lr, cols, temperature=2.269, periodic=True, seed=1, row0=r0, global_rows=rows).init_random()
drv = SlabShardedIsing2D(rows, cols, fac, periodic=True); drv.sweep(3); torch.cuda.synchronize()
a, b = ev(); a.record(); drv.sweep(20); b.record(); torch.cuda.synchronize(); ms = a.elapsed_time(b)
out.append({"config": "C4@1gpu", "workload": "2D Ising 131072x131072, T=2.269, slab driver on 1 GPU, 20 sweeps",
            "ms_per_sweep": ms / 20, "spin_updates_per_s": rows * cols * 20 / ms * 1e3,
            "hbm_frac": rows * cols * 20 / ms * 1e3 * 0.25 / 6553.6e9})
del drv; torch.cuda.empty_cache()
# ---- C5a: 50 temperatures x K ladders x 1024^2 with replica exchange ----------------------------------------
temps = np.linspace(0.1, 5.0, 50); K = 16
fac = lambda n, r0, T: Ising2DEngine(1024, 1024, n_replicas=n, temperature=T, periodic=True, seed=5, replica0=r0).init_random()
pt = LatticeTempering(temps, n_ladders=K, engine_factory=fac, n_sweeps=10, swap_interval=10, seed=9)
for _ in range(10): pt.step()
torch.cuda.synchronize(); a, b = ev(); a.record()
for _ in range(20): pt.step()
b.record(); torch.cuda.synchronize(); ms = a.elapsed_time(b)
out.append({"config": "C5-lattice", "workload": f"50 temperatures x {K} ladders x 1024^2, 20 iterations x 10 sweeps, swap pass every 10 iterations",
            "ms": ms, "spin_updates_per_s": 20 * 10 * K * 50 * 1024 * 1024 / ms * 1e3,
            "swap_accept_rate": float(pt.stats[1] / pt.stats[0])})
del pt; torch.cuda.empty_cache()
# ---- C5b: Langevin 1e6 chains, dim 10 Gaussian ----------------------------------------------------------------
for dtype in ("float32", "float64"):
    tsu = ThermalSamplingUnit(TSUConfig(temperature=1.0, dt=0.01, friction=1.0, n_burnin=100, n_steps=500), seed=1, dtype=dtype)
    tsu.sample_from_energy(QuadraticEnergy(), np.zeros(10), 1000, as_tensor=True); torch.cuda.synchronize()
    ms = 1e30
    for _ in range(2):  # best of two (the first full-size call also pays for lazy kernel loading)
        a, b = ev(); a.record(); x = tsu.sample_from_energy(QuadraticEnergy(), np.zeros(10), 1_000_000, as_tensor=True); b.record()
        torch.cuda.synchronize(); ms = min(ms, a.elapsed_time(b))
    out.append({"config": "C5-langevin", "workload": f"sample_boltzmann E=sum x^2, dim 10, 1e6 chains, 100+500 steps, {dtype}",
                "ms": ms, "chain_steps_per_s": 1e6 * 600 / ms * 1e3, "coordinate_updates_per_s": 1e6 * 6000 / ms * 1e3,
                "variance": float(x.var()), "euler_maruyama_theory": 0.5 / (1 - 0.01)})
for o in out:
    print(json.dumps(o), flush=True)

"""torchrun --nproc-per-node N tools/multigpu_check.py : row-slab sharding on N GPUs
   (1) bit-identical to the 1-GPU run of the same lattice, (2) C4-shaped timing (131072^2, strong scaling)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from tsu_emulator_b200 import Ising2DEngine
from tsu_emulator_b200.distributed import SlabShardedIsing2D

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

# (1) parity: 4096 x 4096, 3 sweeps, sharded vs whole (whole computed on every rank)
rows = cols = 4096
fac = lambda lr, r0: Ising2DEngine(lr, cols, temperature=2.269, periodic=True, seed=7, row0=r0, global_rows=rows).init_random()
drv = SlabShardedIsing2D(rows, cols, fac, periodic=True).sweep(3)
whole = Ising2DEngine(rows, cols, temperature=2.269, periodic=True, seed=7).init_random().sweep(3)
lr = rows // world
same = torch.equal(drv.engine.state[0], whole.state[0][:, rank * lr:(rank + 1) * lr, :])
obs = drv.observables()
obs_whole = whole.observables_tensor()
ok = torch.tensor([int(same and torch.equal(obs, obs_whole))], device="cuda")
if world > 1:
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"[parity] {world} slab(s) == single lattice (state and observables): {bool(ok.item())}", flush=True)

# (2) timing: one 131072 x 131072 lattice, strong scaling
rows = cols = int(os.environ.get("C4_SIZE", 131072))
fac = lambda lr, r0: Ising2DEngine(lr, cols, temperature=2.269, periodic=True, seed=1, row0=r0, global_rows=rows).init_random()
drv = SlabShardedIsing2D(rows, cols, fac, periodic=True)
drv.sweep(3)
n_sw = 20
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); drv.sweep(n_sw); b.record(); torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = t.item()
    upd = rows * cols * n_sw
    print(f"[C4] {rows}x{cols} on {world} GPU(s): {ms/n_sw:.3f} ms/sweep  {upd/ms*1e3:.3e} updates/s "
          f"({upd/ms*1e3/world:.3e} per GPU)", flush=True)
m = drv.observables()
if rank == 0:
    print("[C4] E/N after 23 sweeps:", float(-(2 * rows * cols - 2 * m[0, 1].item()) / (rows * cols)), flush=True)

# (3) replica exchange: K ladders x R temperatures sharded over the ranks (energies all-gathered, every rank runs the same
#     deterministic swap pass) must give the observables of the unsharded run
import torch.distributed  # noqa: E402
from tsu_emulator_b200.distributed import LatticeTempering  # noqa: E402

temps = np.linspace(2.0, 3.0, 8)
def run_pt(group_on):
    fac = lambda n, r0, T: Ising2DEngine(16, 32, n_replicas=n, temperature=T, periodic=True, seed=5, replica0=r0).init_random()
    if group_on:
        pt = LatticeTempering(temps, n_ladders=4, engine_factory=fac, n_sweeps=3, swap_interval=2, seed=123)
    else:  # whole problem on this rank: hide the process group from the driver
        real = (dist.is_initialized,)
        dist.is_initialized = lambda: False
        try:
            pt = LatticeTempering(temps, n_ladders=4, engine_factory=fac, n_sweeps=3, swap_interval=2, seed=123)
        finally:
            dist.is_initialized = real[0]
    for _ in range(40):
        pt.step()
    return pt
pt = run_pt(world > 1)
m_sh, e_sh = pt.observables_by_slot()
sr_sh = pt.slot_replica.cpu().numpy()
ref = run_pt(False)
real_init = dist.is_initialized
dist.is_initialized = lambda: False
try:
    m_1, e_1 = ref.observables_by_slot()
finally:
    dist.is_initialized = real_init
ok = torch.tensor([int(np.array_equal(m_sh, m_1) and np.array_equal(e_sh, e_1) and np.array_equal(sr_sh, ref.slot_replica.cpu().numpy()))], device="cuda")
if world > 1:
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"[tempering] 4 ladders x 8 temperatures sharded over {world} rank(s) == unsharded run: {bool(ok.item())} "
          f"(swap acceptance {float(pt.stats[1]) / max(1.0, float(pt.stats[0])):.3f})", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()

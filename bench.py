#!/usr/bin/env python
"""
bench.py - spin-updates/s of the bit-packed 2-D checkerboard Gibbs path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 2-D Ising 8192 x 8192, T = 2.269, periodic, 4096 independent
replicas per GPU, heat-bath checkerboard sweeps.  One "step" = SWEEPS_PER_STEP full sweeps of all
replicas.  With N > 1 the replica index range is sharded (4096 replicas per rank, no data-path
collective): weak scaling.

value  : updates/s with the lattices resident in HBM (CUDA events, max over ranks).
e2e    : the same through the host API with HOST buffers: every step uploads the packed initial
         lattices from pinned host memory (chunked, overlapped with the sweeps of the previous
         chunk), runs the sweeps and reads magnetisation/energy back.
roofline: HBM bound; algorithmic bytes = 0.25 B per spin update (read 1 neighbour-colour bit,
         write 1 bit), per half-sweep launch, against the measured copy bandwidth.
cpu_baseline: the oracle's literal port of the reference's per-spin NumPy loop
         (oracle/ising2d_oracle.py:gibbs_sweep_port, tsu/gibbs.py:128-162) on the host cores.
"""

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L = 8192
N_REPLICAS = 4096
TEMPERATURE = 2.269
SWEEPS_PER_STEP = 10
SEED = 1234
BYTES_PER_UPDATE = 0.25
WORKLOAD = "ising2d_8192x8192_T2.269_periodic_4096replicas_checkerboard_gibbs"
METRIC = "spin_updates_per_s"
UNIT = "spin-updates/s"


# ----------------------------------------------------------------------------- CPU baseline
def _cpu_worker(args):
    """one process: literal port of the reference loop on its own small periodic lattice"""
    rank, size, n_sweeps = args
    import numpy as np

    from oracle import ising2d_oracle as O

    rng = np.random.default_rng(1000 + rank)
    Jb, hb = O.dense_bit_model(size, size, 1.0, 0.0, True)
    order = O.checkerboard_order(size, size)
    state = rng.integers(0, 2, size * size)
    t0 = time.perf_counter()
    for _ in range(n_sweeps):
        state = O.gibbs_sweep_port(state, Jb, hb, TEMPERATURE, order, rng.random(size * size))
    return time.perf_counter() - t0, int(state.sum())


def cpu_baseline(size=64, n_sweeps=300, cores=None):
    """aggregate updates/s of `cores` independent replicas of a size x size lattice (bounded sample)"""
    import multiprocessing as mp

    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(r, size, n_sweeps) for r in range(cores)])
    wall = time.perf_counter() - t0
    inner = max(r[0] for r in res)
    updates = cores * size * size * n_sweeps
    return {
        "value": updates / inner,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"{cores} independent {size}x{size} periodic lattices x {n_sweeps} sweeps at T={TEMPERATURE}, "
                  f"literal per-spin NumPy loop of tsu/gibbs.py:128-162 (dense {size*size}x{size*size} J), "
                  f"one process per core; same update rule as the workload at reduced size",
        "seconds": wall,
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    vals = []
    last = None
    for i in range(warmup + steps):
        last = cpu_baseline(size=64, n_sweeps=30)
        if i >= warmup:
            vals.append(last["value"])
    v = sum(vals) / len(vals)
    upd_per_step = last["cores"] * 64 * 64 * 30
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": v,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": warmup,
        "ms_per_step": upd_per_step / v * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference CPU path timed on a bounded sample of the workload"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.device_index = device_index
        self.proc = None
        self.path = f"/tmp/tsu_bench_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.device_index)],
                stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for n, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        return out


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic():
    """dram bytes per half-sweep launch from the committed ncu capture of this workload, or None"""
    p = os.path.join(ROOT, "profiles", "lattice_traffic.json")
    try:
        d = json.load(open(p))
        if d.get("workload") == WORKLOAD:
            return float(d["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def profiled_limiter():
    """what the committed ncu capture says actually limits the kernel (the HBM fraction alone would mislead)"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "lattice_traffic.json"))).get("limiter")
    except Exception:
        return None


# ----------------------------------------------------------------------------- B200 arm
def run_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from tsu_emulator_b200.lattice import Ising2DEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # Host side of the end-to-end leg: run this rank (and first-touch its pinned staging buffer) on the CPU cores that
    # are local to its GPU, so that the N ranks of one box do not pull their uploads across the socket interconnect.
    cpu_mask0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa_note = None
    if world > 1 and cpu_mask0 is not None:
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            n_words = (os.cpu_count() + 63) // 64
            words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
            cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1} & cpu_mask0
            if cpus:
                os.sched_setaffinity(0, cpus)
                numa_note = f"rank pinned to the {len(cpus)} CPU cores local to its GPU"
        except Exception as exc:  # no NVML / no permission: keep the inherited mask
            numa_note = f"cpu affinity unchanged ({type(exc).__name__})"
    json_fd = None
    if world > 1:
        # NCCL prints its version / debug lines on fd 1 (the image sets NCCL_DEBUG=VERSION).  stdout must carry only
        # the JSON line: park the real stdout and let everything else go to stderr until the line is written.
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_rep = args.replicas
    size = args.size

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = Ising2DEngine(size, size, n_replicas=n_rep, temperature=TEMPERATURE, periodic=True, seed=SEED,
                        replica0=rank * n_rep)
    eng.init_random()
    updates_per_step = n_rep * size * size * SWEEPS_PER_STEP

    # ---- device-resident timing -------------------------------------------------------
    for _ in range(args.warmup):
        eng.sweep(SWEEPS_PER_STEP)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        eng.sweep(SWEEPS_PER_STEP)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    n_launches = args.steps * SWEEPS_PER_STEP * 2
    value = world * updates_per_step * args.steps / (ms * 1e-3)
    launch_s = ms * 1e-3 / n_launches
    alg_bytes_per_launch = BYTES_PER_UPDATE * n_rep * size * size / 2
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes_per_launch / launch_s / 1e9

    # ---- end to end through the host API: pinned host lattices -> sweeps -> observables -----
    state_bytes = eng.state.numel() * 4
    chunk = max(1, n_rep // 16)
    e2e = None
    try:
        # pinned host source: the full 34.4 GB state at N = 1; with N ranks sharing one host the buffer is capped
        # (host RAM / N) and its chunks are re-used round-robin as the source of the 16 per-step uploads - every
        # step still copies state_bytes from pinned host memory to the device
        host_chunks = 16 if world == 1 else max(1, 16 // world)
        host = torch.empty((host_chunks * chunk,) + tuple(eng.state.shape[1:]), dtype=torch.int32, pin_memory=True)
        host.copy_(eng.state[: host_chunks * chunk])  # synthetic input lattices (any valid packed configuration)
        obs_host = torch.empty((n_rep, 2), dtype=torch.int64, pin_memory=True)
        copy_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()

        def e2e_step():
            evs = []
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_stream(main)
                for k, c0 in enumerate(range(0, n_rep, chunk)):
                    n_c = min(chunk, n_rep - c0)
                    h0 = (k % host_chunks) * chunk
                    eng.state[c0:c0 + n_c].copy_(host[h0:h0 + n_c], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    evs.append(ev)
            # sweeps of chunk k overlap the upload of chunk k+1
            for k, c0 in enumerate(range(0, n_rep, chunk)):
                main.wait_event(evs[k])
                sub = eng.chunk_view(c0, min(chunk, n_rep - c0))
                sub.sweep(SWEEPS_PER_STEP)
            obs = eng.observables_tensor()
            obs_host.copy_(obs, non_blocking=True)
            main.synchronize()
            return obs_host

        for _ in range(max(1, args.warmup - 1)):
            e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        e2e = {
            "value": world * updates_per_step * args.steps / (e2e_ms * 1e-3),
            "unit": UNIT,
            "h2d_bytes_per_step": int(state_bytes) * world,
            "d2h_bytes_per_step": int(n_rep * 16) * world,
            "ms_per_step": e2e_ms / args.steps,
            "api": "Ising2DEngine: pinned host state -> H2D (16 chunks, overlapped) -> sweep(10) -> observables -> D2H",
            "pinned_host_bytes": int(host.numel() * 4),
            "host_affinity": numa_note,
        }
        del host
    except Exception as exc:  # e.g. pinned allocation refused
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(exc)}

    cpu = None
    if cpu_mask0 is not None:
        os.sched_setaffinity(0, cpu_mask0)  # the CPU baseline may use every host core again
    if rank == 0 and not args.no_cpu:
        cpu = cpu_baseline()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u32",
        "data": "synthetic",
        "config": {
            "workload": WORKLOAD if (size == L and n_rep == N_REPLICAS) else f"ising2d_{size}x{size}_T2.269_periodic_{n_rep}replicas",
            "lattice": [size, size],
            "replicas_per_gpu": n_rep,
            "sweeps_per_step": SWEEPS_PER_STEP,
            "temperature": TEMPERATURE,
            "rng": "philox4x32-10, in-register",
            "cache": "inputs_larger_than_L2 (34.4 GB of packed state per GPU)",
            "sharding": "replica index range per rank, no data-path collective",
        },
        "roofline": {
            "bound": "hbm",
            "achieved": achieved,
            "peak": peak,
            "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": profiled_traffic(),
            "peak_source": peak_src,
            "kernel": ("tsu_jit_half_sweep (NVRTC specialisation of half_sweep_fast_body for this temperature)"
                       if getattr(eng, "_jit", 0) > 0 else "half_sweep_fast_kernel"),
            "algorithmic_bytes_per_launch": alg_bytes_per_launch,
            "launch_ms": launch_s * 1e3,
            "measured_limiter": profiled_limiter(),
        },
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": n_launches,
        "clocks": clocks,
    }
    if json_fd is not None:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
        os.close(json_fd)
    else:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--replicas", type=int, default=N_REPLICAS, help="replicas per GPU (default = the named workload)")
    ap.add_argument("--size", type=int, default=L)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()

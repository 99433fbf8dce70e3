"""GPU parity: dense-J Gibbs kernels (through GibbsSampler -> C-ABI) against the reference goldens and the oracle."""
import glob
import os

import numpy as np
import pytest

from oracle import dense_oracle as D

pytestmark = pytest.mark.gpu


def load(golden_dir, pattern):
    paths = sorted(glob.glob(os.path.join(golden_dir, pattern)))
    assert paths, pattern
    return [(p, np.load(p, allow_pickle=False)) for p in paths]


def sampler(T=1.0, **kw):
    from tsu_emulator_b200 import GibbsConfig, GibbsSampler

    cfg = {k: kw.pop(k) for k in list(kw) if k in ("n_burnin", "n_sweeps", "update_order")}
    return GibbsSampler(GibbsConfig(temperature=T, **cfg), **kw)


def test_sweep_goldens_bit_exact(golden_dir):
    """float64 fields: identical decisions to the reference unless a uniform lands within ~1e-15 of p"""
    for path, g in load(golden_dir, "dense_sweep_*.npz"):
        mode = str(g["order_mode"])
        smp = sampler(float(g["T"]), update_order=mode, seed=1)
        out = smp.gibbs_sweep(g["s0"], g["J"], g["b"], n_sweeps=len(g["uniforms"]), _uniforms=g["uniforms"],
                              _orders=g["orders"] if mode == "random" else None)
        assert (out == g["out"]).all(), path


def test_sweep_goldens_float32_mismatch_budget(golden_dir):
    """float32 fields: mismatches only where |u - p| is tiny; counted (north_star item 5)"""
    total = bad = 0
    for path, g in load(golden_dir, "dense_sweep_*.npz"):
        mode = str(g["order_mode"])
        smp = sampler(float(g["T"]), update_order=mode, seed=1, precision="float32")
        out = smp.gibbs_sweep(g["s0"], g["J"], g["b"], n_sweeps=1, _uniforms=g["uniforms"][:1],
                              _orders=g["orders"][:1] if mode == "random" else None)
        want = D.gibbs_sweeps(g["s0"], g["J"], g["b"], float(g["T"]), 1, g["uniforms"][:1],
                              g["orders"][:1] if mode == "random" else None)
        total += out.size
        bad += int((out != want).sum())
    assert bad <= max(1, total // 1000), f"{bad}/{total} float32 mismatches"


def test_boltzmann_golden(golden_dir):
    for path, g in load(golden_dir, "dense_boltzmann_*.npz"):
        smp = sampler(float(g["T"]), n_burnin=int(g["burnin"]), n_sweeps=int(g["n_sweeps"]), seed=2)
        out = smp.sample_boltzmann(g["J"], g["b"], n_samples=int(g["n_samples"]), initial_state=g["s0"], _uniforms=g["uniforms"])
        assert out.dtype == int and out.shape == g["samples"].shape
        assert (out == g["samples"]).all(), path
        assert smp.sample_count == int(g["n_samples"])


def test_annealing_goldens(golden_dir):
    for path, g in load(golden_dir, "dense_anneal_*.npz"):
        smp = sampler(seed=3)
        best, e = smp.simulated_annealing(g["J"], g["b"], T_initial=5.0, T_final=0.2, n_steps=int(g["n_steps"]),
                                          cooling_schedule=str(g["schedule"]), _uniforms=g["uniforms"], _initial_state=g["s0"])
        assert (best == g["best_state"]).all(), path
        assert isinstance(e, float) and e == pytest.approx(float(g["best_energy"]), abs=1e-9)
        assert smp.config.temperature == pytest.approx(float(g["final_temperature"]), rel=1e-12)


def test_tempering_golden(golden_dir):
    for path, g in load(golden_dir, "dense_tempering_*.npz"):
        smp = sampler(1.0, n_burnin=int(g["burnin"]), n_sweeps=int(g["n_sweeps"]), seed=4)
        inj = {"inits": g["inits"], "burn_uniforms": g["burn_uniforms"], "sweep_uniforms": g["sweep_uniforms"],
               "swap_uniforms": g["swap_uniforms"]}
        samples, info = smp.parallel_tempering(g["J"], list(g["temps"]), g["b"], n_samples=int(g["n_samples"]),
                                               swap_interval=int(g["swap_interval"]), _inject=inj)
        assert (samples == g["samples"]).all(), path
        assert info["swap_attempts"] == int(g["swap_attempts"]) and info["swap_accepts"] == int(g["swap_accepts"])
        assert np.allclose(np.array(info["energies"]), g["energies"], atol=1e-9)
        assert (np.array(info["final_states"]) == g["final_states"]).all()
        assert set(info) >= {"swap_acceptance_rate", "energies", "final_states"}


def test_philox_mode_matches_oracle_stream():
    rng = np.random.default_rng(0)
    N, n_chains, n_sweeps, seed = 48, 5, 4, 777
    J = rng.normal(size=(N, N)); J = (J + J.T) / 2
    b = rng.normal(size=N)
    smp = sampler(1.2, seed=seed)
    init = rng.integers(0, 2, (n_chains, N))
    out = smp.sample_chains(J, b, n_chains=n_chains, n_sweeps=n_sweeps, initial_state=init)
    for c in range(n_chains):
        U = np.stack([D.philox_uniforms(seed, c, s, np.arange(N)) for s in range(n_sweeps)])
        want = D.gibbs_sweeps(init[c], J, b, 1.2, n_sweeps, U)
        assert (out[c] == want).all()


def test_random_init_matches_oracle():
    import torch
    from tsu_emulator_b200 import _lib
    st = torch.empty((3, 70), dtype=torch.uint8, device="cuda")
    _lib.call("tsu_dense_init_random", _lib.ptr(st), 3, 70, 99, 4, _lib.current_stream())
    got = st.cpu().numpy()
    for c in range(3):
        assert (got[c] == D.philox_init_state(99, 4 + c, 70)).all()


def test_energy_kernel_matches_reference_formula():
    import torch
    from tsu_emulator_b200 import _lib
    rng = np.random.default_rng(1)
    N, C = 37, 6
    J = rng.normal(size=(N, N))  # asymmetric on purpose
    b = rng.normal(size=N)
    s = rng.integers(0, 2, (C, N))
    Jt = torch.from_numpy(np.ascontiguousarray(J.T)).cuda()
    bt = torch.from_numpy(b).cuda()
    st = torch.from_numpy(s.astype(np.uint8)).cuda()
    e = torch.empty(C, dtype=torch.float64, device="cuda")
    _lib.call("tsu_dense_energy", _lib.ptr(Jt), 1, _lib.ptr(bt), _lib.ptr(st), C, N, _lib.ptr(e), _lib.current_stream())
    for c in range(C):
        assert e[c].item() == pytest.approx(D.compute_energy(s[c], J, b), abs=1e-10)


def test_exact_boltzmann_distribution_small_system():
    """statistical: 4 spins, exact enumeration vs 20000 parallel chains"""
    rng = np.random.default_rng(5)
    N, T = 4, 1.0
    J = rng.normal(size=(N, N)); J = (J + J.T) / 2; np.fill_diagonal(J, 0)
    b = rng.normal(size=N) * 0.3
    states = np.array([[(k >> i) & 1 for i in range(N)] for k in range(2**N)])
    E = np.array([D.compute_energy(s, J, b) for s in states])
    p = np.exp(-E / T); p /= p.sum()
    smp = sampler(T, seed=11)
    out = smp.sample_chains(J, b, n_chains=20000, n_sweeps=30)
    idx = (out * (1 << np.arange(N))).sum(1)
    freq = np.bincount(idx, minlength=2**N) / len(idx)
    assert np.abs(freq - p).max() < 0.015


def test_tensor_core_fields_match_float64():
    """tcgen05 GEMM stage: fields of every site for a batch of chains (gibbs.py:79-100 for all (chain, site))"""
    import torch
    from tsu_emulator_b200 import _lib
    torch.manual_seed(0)
    for N, C in [(128, 128), (256, 130), (1024, 64)]:
        J = (torch.randn(N, N, device="cuda") / N**0.5).to(torch.bfloat16)
        S = (torch.rand(C, N, device="cuda") < 0.5).to(torch.uint8)
        H = torch.full((C, N), float("nan"), device="cuda")
        _lib.call("tsu_dense_tc_debug_fields", _lib.ptr(J), _lib.ptr(S), C, N, _lib.ptr(H), _lib.current_stream())
        ref = S.double() @ J.double().T
        assert (H.double() - ref).abs().max().item() < 1e-4


def tc_run_sweep_by_sweep(J, b, T_chain, init, n_sweeps, seed, sweep0=0, chain0=0):
    """tsu_dense_gibbs_tc_run one sweep per launch: the state of every chain after every sweep, [n_sweeps, C, N]"""
    import torch
    from tsu_emulator_b200 import _lib
    C, N = init.shape
    Jd = torch.from_numpy(np.asarray(J, dtype=np.float32)).cuda().to(torch.bfloat16).contiguous()
    bd = None if b is None else torch.from_numpy(np.asarray(b, dtype=np.float32)).cuda()
    Td = torch.from_numpy(np.asarray(T_chain, dtype=np.float64)).cuda()
    st = torch.from_numpy(np.asarray(init, dtype=np.uint8)).cuda()
    seq = []
    for s_ in range(n_sweeps):
        _lib.call("tsu_dense_gibbs_tc_run", _lib.ptr(Jd), _lib.ptr(bd), _lib.ptr(st), C, N, 1.0, _lib.ptr(Td), 1, seed,
                  sweep0 + s_, chain0, None, _lib.current_stream())
        seq.append(st.cpu().numpy().copy())
    return np.stack(seq)


def tc_disagreements(seq, init, J, b, T_chain, seed, sweep0=0, chain0=0):
    """every (chain, sweep, site) where the tensor-core kernel's bit differs from the float64 rule of the reference
    evaluated on the kernel's own trajectory, with |u - sigmoid(h/T)| of that site (north_star item 5: counted
    and reported).  Chains whose sweep equals the plain oracle sweep are skipped (no disagreement by construction)."""
    n_sweeps, C, N = seq.shape
    found = []
    for s_ in range(n_sweeps):
        start = init if s_ == 0 else seq[s_ - 1]
        for c in range(C):
            U = D.philox_uniforms_tc(seed, chain0 + c, sweep0 + s_, N)
            want = D.gibbs_sweeps(start[c].astype(int), J, b, T_chain[c], 1, U[None])
            if (want != seq[s_, c]).any():
                for site, eps in D.tc_sweep_disagreements(start[c], seq[s_, c], J, b, T_chain[c], U):
                    found.append((c, s_, site, eps))
    return found


# Tolerances of the tensor-core path, |u - sigmoid(h/T)| at a site where it may disagree with the float64 rule:
#   * couplings exactly representable in bf16 with exact fp32 sums (small integers): only the acceptance threshold
#     T * logit(u) is inexact (fp32, lg2.approx: relative error ~1e-6 of |logit| <= 17) -> 5e-6;
#   * Gaussian couplings rounded to bf16 (the oracle gets the same rounded J): plus the fp32 accumulation error of
#     an N-term field, ~1e-6 * sqrt(N) * |J| -> 5e-5 up to N = 4096.
TC_EPS_EXACT_J = 5e-6
TC_EPS_GAUSSIAN_J = 5e-5


def test_tensor_core_sweeps_integer_couplings_exact():
    """integer couplings: bf16 and the fp32 accumulation are exact, so the blocked tensor-core sweep must equal
    the site-by-site oracle (same uniforms) bit for bit unless a uniform falls within float32 rounding of p"""
    from tsu_emulator_b200 import GibbsConfig, GibbsSampler
    rng = np.random.default_rng(3)
    N, C, n_sweeps, seed, T = 128, 40, 3, 4242, 1.7
    J = rng.integers(-2, 3, (N, N)).astype(np.float64)
    J = np.triu(J, 1); J = J + J.T
    np.fill_diagonal(J, rng.integers(-1, 2, N))          # self-couplings are part of the field (gibbs.py:97)
    b = rng.integers(-1, 2, N).astype(np.float64) * 0.5
    init = rng.integers(0, 2, (C, N))
    smp = GibbsSampler(GibbsConfig(temperature=T), seed=seed, precision="bf16")
    out = smp.sample_chains(J, b, n_chains=C, n_sweeps=n_sweeps, initial_state=init)
    seq = tc_run_sweep_by_sweep(J, b, np.full(C, T), init, n_sweeps, seed)
    assert (seq[-1] == out).all()                        # one launch of 3 sweeps == 3 launches of one sweep
    found = tc_disagreements(seq, init, J, b, np.full(C, T), seed)
    print(f"tensor-core path, integer J: {len(found)} disagreeing sites of {C * N * n_sweeps}, "
          f"max |u - p| = {max([f[3] for f in found], default=0.0):.2e}")
    assert all(eps < TC_EPS_EXACT_J for *_, eps in found), found
    assert len(found) <= 2


def test_tensor_core_sweeps_gaussian_couplings_mismatch_budget():
    """SK couplings rounded to bf16 (the oracle gets the same rounded J), three sweeps: the kernel disagrees with the
    float64 rule only at sites whose uniform lands within the stated fp32 field error of the acceptance
    probability; every such site is counted and its |u - p| checked (north_star item 5)"""
    import torch
    from tsu_emulator_b200 import GibbsConfig, GibbsSampler
    rng = np.random.default_rng(4)
    N, C, seed, T, n_sweeps = 256, 64, 99, 1.0, 3
    J = rng.normal(size=(N, N)) / np.sqrt(N); J = (J + J.T) / np.sqrt(2); np.fill_diagonal(J, 0)
    J = torch.from_numpy(J).to(torch.bfloat16).to(torch.float64).numpy()
    init = rng.integers(0, 2, (C, N))
    smp = GibbsSampler(GibbsConfig(temperature=T), seed=seed, precision="bf16")
    out, e = smp.sample_chains(J, None, n_chains=C, n_sweeps=n_sweeps, initial_state=init, return_energy=True)
    for c in range(C):
        assert e[c] == pytest.approx(D.compute_energy(out[c].astype(float), J), abs=1e-3)
    seq = tc_run_sweep_by_sweep(J, None, np.full(C, T), init, n_sweeps, seed)
    assert (seq[-1] == out).all()
    found = tc_disagreements(seq, init, J, None, np.full(C, T), seed)
    print(f"tensor-core path, Gaussian J N={N}: {len(found)} disagreeing sites of {C * N * n_sweeps}, "
          f"max |u - p| = {max([f[3] for f in found], default=0.0):.2e}")
    assert all(eps < TC_EPS_GAUSSIAN_J for *_, eps in found), found
    assert len(found) <= 8


def test_tensor_core_full_size_n4096():
    """BASELINE config 3's matrix size: N = 4096 = 32 panels (the J ring wraps ten times per panel, every
    panel_done phase is used), Gaussian SK couplings, 6 chains x 2 sweeps against the float64 rule"""
    import torch
    rng = np.random.default_rng(7)
    N, C, seed, T, n_sweeps = 4096, 6, 31337, 1.0, 2
    J = rng.normal(size=(N, N)) / np.sqrt(N); J = (J + J.T) / np.sqrt(2); np.fill_diagonal(J, 0)
    J = torch.from_numpy(J).to(torch.bfloat16).to(torch.float64).numpy()
    init = rng.integers(0, 2, (C, N))
    seq = tc_run_sweep_by_sweep(J, None, np.full(C, T), init, n_sweeps, seed, sweep0=3, chain0=40)
    found = tc_disagreements(seq, init, J, None, np.full(C, T), seed, sweep0=3, chain0=40)
    print(f"tensor-core path, Gaussian J N={N}: {len(found)} disagreeing sites of {C * N * n_sweeps}, "
          f"max |u - p| = {max([f[3] for f in found], default=0.0):.2e}")
    assert all(eps < TC_EPS_GAUSSIAN_J for *_, eps in found), found
    assert len(found) <= 6
    # and the two sweeps in ONE launch give the same bits as two launches
    from tsu_emulator_b200 import _lib
    Jd = torch.from_numpy(J.astype(np.float32)).cuda().to(torch.bfloat16).contiguous()
    st = torch.from_numpy(init.astype(np.uint8)).cuda()
    _lib.call("tsu_dense_gibbs_tc_run", _lib.ptr(Jd), None, _lib.ptr(st), C, N, T, None, n_sweeps, seed, 3, 40, None,
              _lib.current_stream())
    assert (st.cpu().numpy() == seq[-1]).all()


def test_sample_boltzmann_on_tensor_cores_schedule_and_shapes():
    """GibbsSampler(precision='bf16').sample_boltzmann / .sample: burn-in + n_samples x n_sweeps schedule of
    gibbs.py:198-211 on the tcgen05 kernel (sample k = state after burnin + (k+1) n_sweeps sweeps), N padded to the
    panel size, the reference's return contract (int array (n_samples, n_bits) / (n_chains, n_samples, n_bits))"""
    from tsu_emulator_b200 import GibbsConfig, GibbsSampler, HardwareEmulator
    rng = np.random.default_rng(12)
    N, C, seed, T = 100, 70, 5, 1.3           # 100 bits: padded to one 128-site panel
    J = rng.integers(-1, 2, (N, N)).astype(np.float64)
    J = np.triu(J, 1); J = J + J.T
    b = rng.integers(-1, 2, N).astype(np.float64)
    init = rng.integers(0, 2, (C, N))
    cfg = GibbsConfig(temperature=T, n_burnin=2, n_sweeps=3)
    smp = GibbsSampler(cfg, seed=seed, precision="bf16")
    out = smp.sample_boltzmann(J, b, n_samples=2, initial_state=init, n_chains=C)
    assert out.shape == (C, 2, N) and out.dtype == int and set(np.unique(out)) <= {0, 1}
    assert smp.sample_count == 2
    # the same schedule, sweep by sweep, on the padded problem (what the host mirror launches)
    Jp = np.zeros((128, 128)); Jp[:N, :N] = J
    bp = np.zeros(128); bp[:N] = b
    ip = np.zeros((C, 128), dtype=np.int64); ip[:, :N] = init
    seq = tc_run_sweep_by_sweep(Jp, bp, np.full(C, T), ip, 8, seed)
    assert (out[:, 0, :] == seq[4][:, :N]).all() and (out[:, 1, :] == seq[7][:, :N]).all()
    found = tc_disagreements(seq, ip, Jp, bp, np.full(C, T), seed)
    assert all(eps < TC_EPS_EXACT_J for *_, eps in found)
    one = GibbsSampler(cfg, seed=seed, precision="bf16").sample(J, n_samples=3)
    assert one.shape == (3, N) and one.dtype == int
    with pytest.raises(ValueError):
        GibbsSampler(GibbsConfig(update_order="random"), precision="bf16").sample_boltzmann(J, n_samples=1)
    with pytest.raises(ValueError):
        GibbsSampler(cfg, precision="bf16").sample_boltzmann(np.zeros((4224, 4224)), n_samples=1)
    # HardwareEmulator.sample_parallel (gibbs.py:450-487): +-1 couplings and >= 64 chains take the tensor cores
    hw = HardwareEmulator(n_bits=N, parallel_chains=80)
    samples, timing = hw.sample_parallel(J, 200, temperature=2.0)
    assert samples.shape == (200, N) and set(np.unique(samples)) <= {0, 1} and timing["batches_needed"] == 3


@pytest.mark.parametrize("tile_m", ["64", "128"])
def test_tensor_core_multi_panel_multi_sweep_exact(tile_m, monkeypatch):
    """three 128-site panels, two sweeps, bias and per-chain temperatures through the C-ABI, for both tile heights
    (chains per CTA = 64 / 128): exercises the late K-chunk ordering, the in-panel correction MMAs and the ring
    wrap-around.  Integer couplings keep bf16 / fp32 exact, so the result must equal the site-by-site oracle."""
    import torch
    from tsu_emulator_b200 import _lib
    monkeypatch.setenv("TSU_TC_M", tile_m)
    rng = np.random.default_rng(11)
    N, C, n_sweeps, seed = 384, 70, 2, 777
    J = rng.integers(-2, 3, (N, N)).astype(np.float64)
    J = np.triu(J, 1); J = J + J.T
    np.fill_diagonal(J, rng.integers(-1, 2, N))
    b = rng.integers(-2, 3, N).astype(np.float64) * 0.25
    T = rng.uniform(0.8, 3.0, C)
    init = rng.integers(0, 2, (C, N)).astype(np.uint8)
    Jd = torch.from_numpy(J).cuda().to(torch.bfloat16).contiguous()
    bd = torch.from_numpy(b.astype(np.float32)).cuda()
    Td = torch.from_numpy(T).cuda()
    st = torch.from_numpy(init).cuda()
    sweep0, chain0 = 5, 1000
    _lib.call("tsu_dense_gibbs_tc_run", _lib.ptr(Jd), _lib.ptr(bd), _lib.ptr(st), C, N, 1.0, _lib.ptr(Td), n_sweeps, seed,
              sweep0, chain0, None, _lib.current_stream())
    out = st.cpu().numpy()
    bad = 0
    for c in range(C):
        U = np.stack([D.philox_uniforms_tc(seed, chain0 + c, sweep0 + s, N) for s in range(n_sweeps)])
        want = D.gibbs_sweeps(init[c].astype(int), J, b, T[c], n_sweeps, U)
        bad += int((out[c] != want).any())
    assert bad <= 1, f"{bad} of {C} chains differ (tile height {tile_m})"


def test_uniforms_on_the_acceptance_probability_and_clamp_edges():
    """the float64 kernel decides by logit(u) < h / T and falls back to the reference's u < sigmoid(h / T) when the two
    sides are close: uniforms placed exactly on p, one ulp below and above it, plus fields on both sides of the +-20
    clamp (gibbs.py:65-70), must give the reference's bits"""
    rng = np.random.default_rng(11)
    N, n_sweeps, T = 24, 6, 0.5
    # integer couplings, biases in quarters, T = 1/2: every field and h / T is exact in any summation order, so the
    # kernel's incrementally maintained fields equal the oracle's dot products bit for bit
    J = rng.integers(-3, 4, size=(N, N)).astype(float)
    J = np.triu(J, 1)
    J = J + J.T
    J[0, 1] = J[1, 0] = 30.0     # fields far beyond the clamp once both are up
    J[2, 3] = J[3, 2] = -30.0
    b = rng.integers(-8, 9, size=N) / 4.0
    b[4] = 20.0 * T              # h / T exactly 20: not clamped (strict >)
    b[5] = -20.0 * T
    b[6] = 20.0 * T + 0.25       # just beyond
    b[7] = -20.0 * T - 0.25
    for k in (4, 5, 6, 7):
        J[k, :] = J[:, k] = 0.0
    s0 = rng.integers(0, 2, N)
    # walk the oracle and put every uniform on / next to the acceptance probability of its visit
    state = s0.copy()
    U = np.zeros((n_sweeps, N))
    for s in range(n_sweeps):
        for i in range(N):
            p = D.sigmoid_ref(float(np.dot(J[i, :], state) + b[i]) / T)
            choice = (s + i) % 4
            u = [p, np.nextafter(p, 0.0), np.nextafter(p, 1.0), rng.random()][choice]
            u = min(max(u, 0.0), np.nextafter(1.0, 0.0))
            U[s, i] = u
            state[i] = 1 if u < p else 0
    want = D.gibbs_sweeps(s0, J, b, T, n_sweeps, U)
    assert (want == state).all()
    got = sampler(T, seed=1).gibbs_sweep(s0, J, b, n_sweeps=n_sweeps, _uniforms=U)
    assert (got == want).all()

"""Chromatic sparse-coupling Gibbs sampler (csrc/sparse_gibbs.cu).  Goldens: the UNMODIFIED reference's gibbs_sweep with
update_order="random" and numpy.random.permutation fixed to the greedy colour-class order (oracle/make_golden_dense.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import dense_oracle as D


def goldens(golden_dir):
    paths = sorted(glob.glob(os.path.join(golden_dir, "sparse_sweep_*.npz")))
    assert len(paths) >= 2
    return [(p, np.load(p)) for p in paths]


def test_oracle_and_host_colouring_reproduce_the_reference_goldens(golden_dir):
    from tsu_emulator_b200.sparse import colour_classes, greedy_colouring, to_csr

    for path, g in goldens(golden_dir):
        order, colour = D.greedy_colour_order(g["J"])
        assert (order == g["order"]).all() and (colour == g["colour"]).all()
        J = g["J"]
        for i, j in zip(*np.nonzero(J)):
            assert i == j or colour[i] != colour[j]          # a proper colouring: coupled sites never share a class
        orders = np.tile(g["order"], (len(g["uniforms"]), 1))
        out = D.gibbs_sweeps(g["s0"], J, g["b"], float(g["T"]), len(g["uniforms"]), g["uniforms"], orders)
        assert (out == g["out"]).all(), path
        rowptr, col, val, N = to_csr(J)
        assert N == J.shape[0] and val.size == np.count_nonzero(J)
        for i in range(N):
            assert (col[rowptr[i]:rowptr[i + 1]] == np.flatnonzero(J[i])).all()
        c = greedy_colouring(rowptr, col, N)
        assert (c == colour).all()
        cptr, sites = colour_classes(c)
        assert (sites == order).all() and cptr[-1] == N


def test_csr_from_scipy_and_tuple():
    import scipy.sparse as sp
    from tsu_emulator_b200.sparse import to_csr

    rng = np.random.default_rng(0)
    J = np.where(rng.random((30, 30)) < 0.1, rng.normal(size=(30, 30)), 0.0)
    a = to_csr(J)
    b = to_csr(sp.coo_matrix(J))
    c = to_csr(a)
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)
    for x, y in zip(a[:3], c[:3]):
        assert np.array_equal(x, y)
    with pytest.raises(ValueError):
        to_csr(np.zeros((3, 4)))


@pytest.mark.gpu
def test_kernel_reproduces_the_reference_goldens(golden_dir):
    """injected uniforms in visiting order: bit-exact with the unmodified reference (float64 fields, sums of at most a
    handful of terms in ascending column order)"""
    from tsu_emulator_b200 import GibbsConfig, GibbsSampler

    for path, g in goldens(golden_dir):
        smp = GibbsSampler(GibbsConfig(temperature=float(g["T"])), seed=1)
        out = smp.gibbs_sweep(g["s0"], g["J"], g["b"], n_sweeps=len(g["uniforms"]), chromatic=True, _uniforms=g["uniforms"])
        assert (out == g["out"]).all(), path


@pytest.mark.gpu
def test_philox_mode_equals_dense_kernel_in_colour_order_and_oracle():
    """same Philox uniform per (site, chain, sweep) as the dense sampler: the chromatic kernel, the dense kernel driven
    with the colour order as its permutation, and the oracle give the same bits, for several chains at once"""
    from tsu_emulator_b200 import GibbsConfig, GibbsSampler
    from tsu_emulator_b200.sparse import colour_classes, greedy_colouring, to_csr

    rng = np.random.default_rng(5)
    n, C, n_sweeps, seed, T = 200, 5, 3, 77, 1.1
    J = np.zeros((n, n))
    for _ in range(500):
        i, j = rng.integers(0, n, 2)
        if i != j:
            J[i, j] = J[j, i] = float(rng.integers(-3, 4)) * 0.5
    b = rng.integers(-2, 3, n) * 0.25
    init = rng.integers(0, 2, (C, n))
    rowptr, col, _, _ = to_csr(J)
    order = colour_classes(greedy_colouring(rowptr, col, n))[1]
    a = GibbsSampler(GibbsConfig(temperature=T, n_sweeps=1, n_burnin=0), seed=seed)
    got = a.sample_boltzmann(J, b, n_samples=n_sweeps, burnin=0, initial_state=init, n_chains=C, chromatic=True)
    assert got.shape == (C, n_sweeps, n)
    d = GibbsSampler(GibbsConfig(temperature=T, n_sweeps=1, n_burnin=0, update_order="random"), seed=seed)
    dense = d.sample_boltzmann(J, b, n_samples=n_sweeps, burnin=0, initial_state=init, n_chains=C,
                               _orders=np.tile(order, (n_sweeps, 1)))
    assert (got == dense).all()
    for c in range(C):
        U = np.stack([D.philox_uniforms(seed, c, s, order) for s in range(n_sweeps)])
        want = D.gibbs_sweeps(init[c], J, b, T, n_sweeps, U, np.tile(order, (n_sweeps, 1)))
        assert (got[c, -1] == want).all()


@pytest.mark.gpu
def test_ising_chain_runs_without_a_dense_matrix():
    """IsingChain (ising.py:265-304): 100000 spins sample in milliseconds and never allocate the 80 GB matrix the
    reference would; nearest-neighbour correlation of the open chain is tanh(J/T) exactly"""
    from tsu_emulator_b200 import IsingChain, IsingConfig

    n, Jc, T = 100_000, 1.0, 1.5
    ch = IsingChain(n, J=Jc, config=IsingConfig(temperature=T, n_burnin=200, n_sweeps=20), seed=3)
    s = ch.sample(4)
    assert s.shape == (4, n) and set(np.unique(s)) <= {-1, 1} and ch._J is None
    corr = float(np.mean(s[:, :-1] * s[:, 1:]))
    assert abs(corr - np.tanh(Jc / T)) < 0.01, corr
    assert abs(ch.energy(s[0]) / n + Jc * np.tanh(Jc / T)) < 0.02
    gs, e = IsingChain(300, J=1.0, config=IsingConfig(temperature=1.0), seed=4).find_ground_state(n_steps=400)
    assert e <= -0.97 * 299 and gs.shape == (300,)   # at most a few domain walls survive the anneal
    # a chain whose couplings were edited falls back to the general (matrix) path and still samples
    c2 = IsingChain(12, J=1.0, config=IsingConfig(temperature=2.0, n_burnin=10, n_sweeps=2), seed=5)
    c2.set_coupling(0, 11, 1.0)
    assert c2.sample(3).shape == (3, 12) and c2.J[0, 11] == 1.0


@pytest.mark.gpu
def test_sparse_annealing_and_exact_distribution():
    """chromatic sweeps sample the exact Boltzmann distribution of a small sparse model; simulated annealing on a sparse
    MAX-CUT-like instance returns (state, energy) with energy == compute_energy(state) as the reference test asks
    (tests/test_gibbs.py:170-182)"""
    from tsu_emulator_b200 import GibbsConfig, GibbsSampler

    rng = np.random.default_rng(2)
    n, T = 6, 1.0
    J = np.zeros((n, n))
    for i, j, w in [(0, 1, 1.0), (1, 2, -0.7), (2, 3, 0.5), (3, 4, 1.2), (4, 5, -1.0), (0, 5, 0.8)]:
        J[i, j] = J[j, i] = w
    b = rng.normal(size=n) * 0.3
    states = np.array([[(k >> i) & 1 for i in range(n)] for k in range(2**n)])
    E = np.array([D.compute_energy(s_, J, b) for s_ in states])
    p = np.exp(-E / T); p /= p.sum()
    smp = GibbsSampler(GibbsConfig(temperature=T, n_burnin=30, n_sweeps=1), seed=11)
    out = smp.sample_boltzmann(J, b, n_samples=1, n_chains=30000, chromatic=True)[:, 0, :]
    freq = np.bincount((out * (1 << np.arange(n))).sum(1), minlength=2**n) / len(out)
    assert np.abs(freq - p).max() < 0.012
    m = 80
    A = np.zeros((m, m))
    for _ in range(240):
        i, j = rng.integers(0, m, 2)
        if i != j:
            A[i, j] = A[j, i] = -1.0
    st, e = GibbsSampler(GibbsConfig(), seed=3).simulated_annealing(A, None, T_initial=5.0, T_final=0.05, n_steps=300,
                                                                   chromatic=True, n_chains=16)
    assert e == pytest.approx(D.compute_energy(st, A, None), abs=1e-9)

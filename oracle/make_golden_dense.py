"""golden fixtures for the dense-J path from the UNMODIFIED reference (run via python -m oracle.make_golden)."""
import inspect
import os

import numpy as np

from .ref_loader import injected_numpy_random, load_reference

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _problem(n, seed, integer=False):
    rng = np.random.default_rng(seed)
    if integer:
        J = rng.integers(-2, 3, (n, n)).astype(np.float64)
    else:
        J = rng.normal(size=(n, n))
    J = (J + J.T) / 2
    b = rng.normal(size=n) * 0.5
    return rng, J, b


def sweep_case(name, n, T, n_sweeps, order_mode, seed, self_coupling=False, integer=False):
    gibbs, _, _ = load_reference()
    rng, J, b = _problem(n, seed, integer)
    if not self_coupling:
        np.fill_diagonal(J, 0.0)
    s0 = rng.integers(0, 2, n)
    U = rng.random((n_sweeps, n))
    orders = None
    smp = gibbs.GibbsSampler(gibbs.GibbsConfig(temperature=T, update_order=order_mode))
    if order_mode == "random":
        orders = np.stack([rng.permutation(n) for _ in range(n_sweeps)])
        it = iter(orders)
        from unittest import mock
        with injected_numpy_random(uniforms=U.ravel()), mock.patch("numpy.random.permutation", lambda k: next(it)):
            out = smp.gibbs_sweep(s0.copy(), J, b, n_sweeps=n_sweeps)
    else:
        with injected_numpy_random(uniforms=U.ravel()):
            out = smp.gibbs_sweep(s0.copy(), J, b, n_sweeps=n_sweeps)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"dense_sweep_{name}.npz"), J=J, b=b, T=T, s0=s0, uniforms=U,
                        orders=orders if orders is not None else np.zeros(0), order_mode=order_mode, out=out,
                        energy=smp.compute_energy(out, J, b))


def boltzmann_case(name, n, T, burnin, n_samples, n_sweeps, seed):
    gibbs, _, _ = load_reference()
    rng, J, b = _problem(n, seed)
    s0 = rng.integers(0, 2, n)
    total = burnin + n_samples * n_sweeps
    U = rng.random((total, n))
    smp = gibbs.GibbsSampler(gibbs.GibbsConfig(temperature=T, n_burnin=burnin, n_sweeps=n_sweeps))
    with injected_numpy_random(uniforms=U.ravel(), randint=s0):
        out = smp.sample_boltzmann(J, b, n_samples=n_samples)  # random init comes from the patched randint
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"dense_boltzmann_{name}.npz"), J=J, b=b, T=T, s0=s0, uniforms=U,
                        burnin=burnin, n_samples=n_samples, n_sweeps=n_sweeps, samples=out)


def annealing_case(name, n, n_steps, schedule, seed):
    gibbs, _, _ = load_reference()
    rng, J, b = _problem(n, seed)
    s0 = rng.integers(0, 2, n)
    U = rng.random((n_steps, n))
    smp = gibbs.GibbsSampler(gibbs.GibbsConfig())
    with injected_numpy_random(uniforms=U.ravel(), randint=s0):
        best, e = smp.simulated_annealing(J, b, T_initial=5.0, T_final=0.2, n_steps=n_steps, cooling_schedule=schedule)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"dense_anneal_{name}.npz"), J=J, b=b, s0=s0, uniforms=U,
                        n_steps=n_steps, schedule=schedule, best_state=best, best_energy=e,
                        final_temperature=smp.config.temperature)


def tempering_case(name, n, temps, burnin, n_sweeps, n_samples, swap_interval, seed):
    """the reference draws sweep and swap uniforms from ONE stream; record which call site consumed each draw"""
    gibbs, _, _ = load_reference()
    rng, J, b = _problem(n, seed)
    R = len(temps)
    inits = [rng.integers(0, 2, n) for _ in range(R)]
    stream = rng.random(200000)
    log = []  # (kind, value)
    pos = [0]

    def fake_rand(*a):
        v = stream[pos[0]]
        pos[0] += 1
        caller = inspect.stack()[1].function
        log.append(("swap" if caller == "parallel_tempering" else "sweep", v))
        return v

    init_it = iter(inits)
    from unittest import mock
    smp = gibbs.GibbsSampler(gibbs.GibbsConfig(temperature=1.0, n_burnin=burnin, n_sweeps=n_sweeps))
    with mock.patch("numpy.random.rand", fake_rand), mock.patch("numpy.random.randint", lambda *a, **k: next(init_it).copy()):
        samples, info = smp.parallel_tempering(J, list(temps), b, n_samples=n_samples, swap_interval=swap_interval)
    sweep_draws = np.array([v for k, v in log if k == "sweep"])
    swap_draws = [v for k, v in log if k == "swap"]
    nb = R * burnin * n
    burn_u = sweep_draws[:nb].reshape(R, burnin, n)
    sweep_u = sweep_draws[nb:].reshape(n_samples, R, n_sweeps, n)
    # swap draws are consumed only when delta < 0: replay the reference's decisions to place them per (iteration, pair)
    from . import dense_oracle as D
    swap_u = np.full((n_samples, R - 1), 2.0)  # 2.0 = "never accept" filler for unconsumed slots
    states = [s.copy() for s in inits]
    for i in range(R):
        states[i] = D.gibbs_sweeps(states[i], J, b, temps[i], burnin, burn_u[i])
    k = 0
    for it in range(n_samples):
        for i in range(R):
            states[i] = D.gibbs_sweeps(states[i], J, b, temps[i], n_sweeps, sweep_u[it][i])
        if (it + 1) % swap_interval == 0:
            for i in range(R - 1):
                Ei, Ej = D.compute_energy(states[i], J, b), D.compute_energy(states[i + 1], J, b)
                delta = (1.0 / temps[i] - 1.0 / temps[i + 1]) * (Ej - Ei)
                if delta >= 0:
                    states[i], states[i + 1] = states[i + 1], states[i]
                else:
                    u = swap_draws[k]
                    k += 1
                    swap_u[it, i] = u
                    if u < np.exp(delta):
                        states[i], states[i + 1] = states[i + 1], states[i]
    assert k == len(swap_draws)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"dense_tempering_{name}.npz"), J=J, b=b, temps=np.array(temps),
                        burnin=burnin, n_sweeps=n_sweeps, n_samples=n_samples, swap_interval=swap_interval,
                        inits=np.array(inits), burn_uniforms=burn_u, sweep_uniforms=sweep_u, swap_uniforms=swap_u,
                        samples=samples, swap_attempts=info["swap_attempts"], swap_accepts=info["swap_accepts"],
                        energies=np.array(info["energies"]), final_states=np.array(info["final_states"]))


def chromatic_case(name, J, b, T, n_sweeps, seed):
    """sparse couplings: the reference's sweep with update_order="random" and the permutation fixed to the greedy
    colour-class order - what csrc/sparse_gibbs.cu computes with the classes updated concurrently"""
    from unittest import mock

    from . import dense_oracle as D

    gibbs, _, _ = load_reference()
    rng = np.random.default_rng(seed)
    n = J.shape[0]
    order, colour = D.greedy_colour_order(J)
    s0 = rng.integers(0, 2, n)
    U = rng.random((n_sweeps, n))   # U[t, k] is consumed by the k-th visited site of sweep t
    smp = gibbs.GibbsSampler(gibbs.GibbsConfig(temperature=T, update_order="random"))
    with injected_numpy_random(uniforms=U.ravel()), mock.patch("numpy.random.permutation", lambda k: order.copy()):
        out = smp.gibbs_sweep(s0.copy(), J, b, n_sweeps=n_sweeps)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"sparse_sweep_{name}.npz"), J=J, b=b, T=T, s0=s0, uniforms=U,
                        order=order, colour=colour, out=out, energy=smp.compute_energy(out, J, b))


def chromatic_cases():
    _, _, ising = load_reference()
    # the bit model of the reference's IsingChain (ising.py:265-304, 127-138) with the physical bias
    ch = ising.IsingChain(41, J=0.8, config=ising.IsingConfig(temperature=1.2, external_field=0.3))
    chromatic_case("chain41", 4 * ch.J, 2 * ch.h - 2 * ch.J.sum(1), 1.2, 4, 31)
    # irregular sparse graph with a few self-couplings and integer weights
    rng = np.random.default_rng(32)
    n = 60
    J = np.zeros((n, n))
    for _ in range(130):
        i, j = rng.integers(0, n, 2)
        if i != j:
            J[i, j] = J[j, i] = float(rng.integers(-2, 3))
    for i in rng.choice(n, 5, replace=False):
        J[i, i] = float(rng.integers(1, 3))
    chromatic_case("graph60", J, rng.normal(size=n) * 0.5, 0.9, 3, 33)


def main():
    sweep_case("n12_seq", 12, 1.3, 5, "sequential", 21)
    sweep_case("n33_rand", 33, 0.8, 4, "random", 22)
    sweep_case("n8_self_int", 8, 1.0, 6, "sequential", 23, self_coupling=True, integer=True)
    sweep_case("n70_cold", 70, 0.05, 3, "sequential", 24)
    boltzmann_case("n10", 10, 1.1, 4, 6, 3, 25)
    annealing_case("n14_exp", 14, 25, "exponential", 26)
    annealing_case("n9_lin", 9, 12, "linear", 27)
    tempering_case("n10_r4", 10, [0.5, 1.0, 2.0, 4.0], 2, 2, 8, 2, 28)
    chromatic_cases()
    print("dense goldens written")


if __name__ == "__main__":
    main()

"""
ORACLE (test infrastructure, not product code): CPU restatement of the reference's Langevin sampler
with the normal draws injected.

Follows /root/reference/tsu/core.py: _langevin_step (64-80), _numerical_gradient (82-98),
sample_from_energy (100-162).  Pinned by tests/golden/langevin_*.npz produced by
oracle/make_golden_langevin.py from the unmodified reference with numpy.random.randn patched.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""

import numpy as np


def numerical_gradient(energy_fn, x, eps=1e-5):
    x = np.atleast_1d(x)
    grad = np.zeros_like(x)
    for i in range(len(x)):
        xp = x.copy()
        xp[i] += eps
        xm = x.copy()
        xm[i] -= eps
        grad[i] = (float(energy_fn(xp)) - float(energy_fn(xm))) / (2 * eps)
    return grad


def sample_from_energy(energy_fn, x_init, n_samples, normals, temperature=1.0, dt=0.01, friction=1.0, n_burnin=100,
                       n_steps=500, return_trajectory=False):
    """core.py:100-162.  normals[c] has 1 + n_burnin + n_steps rows of dim draws for chain c; row 0 is the
    start jitter (unused for c == 0, core.py:142-143)."""
    x_init = np.atleast_1d(np.asarray(x_init, dtype=np.float64))
    noise_scale = np.sqrt(2 * temperature * dt / friction)
    samples, traj = [], []
    x = x_init.copy()
    for c in range(n_samples):
        if c > 0:
            x = x_init + 0.1 * normals[c][0]
        for s in range(n_burnin + n_steps):
            grad = numerical_gradient(energy_fn, x)
            drift = -grad * dt / friction
            x = x + drift + noise_scale * normals[c][1 + s]
            if return_trajectory and s >= n_burnin:
                traj.append(x.copy())
        samples.append(x.copy())
    samples = np.array(samples)
    return (samples, traj) if return_trajectory else samples


def quadratic_energy(x):
    """README.md:60-61"""
    return (np.asarray(x) ** 2).sum()


def gaussian_energy(mu, sigma):
    """core.py:227-230 generalised to vectors"""
    return lambda x: float(0.5 * np.sum(((np.atleast_1d(x) - mu) / sigma) ** 2))


def mixture_energy(centers, weights):
    """tsu/api.py:143-149"""
    centers = [np.array(c, dtype=np.float64) for c in centers]
    weights = np.array(weights, dtype=np.float64) / np.sum(weights)

    def f(x):
        x = np.atleast_1d(x)
        prob = 0
        for i, c in enumerate(centers):
            prob += weights[i] * np.exp(-0.5 * np.sum((x - c) ** 2))
        return -np.log(prob + 1e-10)

    return f


def langevin_chain_steps_per_s_port(n_chains=4, dim=10, n_burnin=100, n_steps=500, seed=0):
    """timing helper for the CPU baseline: README quadratic energy through the literal port"""
    import time

    rng = np.random.default_rng(seed)
    normals = rng.normal(size=(n_chains, 1 + n_burnin + n_steps, dim))
    t0 = time.perf_counter()
    sample_from_energy(quadratic_energy, np.zeros(dim), n_chains, normals, n_burnin=n_burnin, n_steps=n_steps)
    dt = time.perf_counter() - t0
    return n_chains * (n_burnin + n_steps) / dt

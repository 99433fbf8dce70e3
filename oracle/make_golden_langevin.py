"""golden fixtures for the Langevin path from the UNMODIFIED reference (run via python -m oracle.make_golden)."""
import os

import numpy as np

from . import langevin_oracle as LO
from .ref_loader import injected_numpy_random, load_reference

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def case(name, energy_fn, meta, dim, x_init, n_samples, n_burnin, n_steps, T, dt, friction, seed):
    _, core, _ = load_reference()
    rng = np.random.default_rng(seed)
    normals = rng.normal(size=(n_samples, 1 + n_burnin + n_steps, dim))
    # the reference draws: (for c > 0: one jitter vector) then n_burnin + n_steps step vectors, chain after chain
    seq = []
    for c in range(n_samples):
        if c > 0:
            seq.append(normals[c][0])
        seq.extend(normals[c][1:])
    tsu = core.ThermalSamplingUnit(core.TSUConfig(temperature=T, dt=dt, friction=friction, n_burnin=n_burnin, n_steps=n_steps))
    with injected_numpy_random(normals=seq):
        samples, traj = tsu.sample_from_energy(energy_fn, np.array(x_init, dtype=np.float64), n_samples, return_trajectory=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"langevin_{name}.npz"), dim=dim, x_init=np.array(x_init, dtype=np.float64),
                        n_samples=n_samples, n_burnin=n_burnin, n_steps=n_steps, T=T, dt=dt, friction=friction,
                        normals=normals, samples=samples, trajectory=np.array(traj), **meta)


def main():
    case("quadratic_d3", LO.quadratic_energy, {"kind": "quadratic"}, 3, [0.5, -1.0, 2.0], 4, 5, 7, 1.0, 0.01, 1.0, 31)
    case("quadratic_d10", LO.quadratic_energy, {"kind": "quadratic"}, 10, np.zeros(10), 3, 10, 20, 1.0, 0.01, 1.0, 32)
    mu, sigma = np.array([5.0]), np.array([2.0])
    case("gaussian_d1", LO.gaussian_energy(mu, sigma), {"kind": "gaussian", "mu": mu, "sigma": sigma}, 1, [5.0], 5, 6, 9, 0.7, 0.05, 2.0, 33)
    centers, weights = [[-2.0, 0.0], [2.0, 1.0], [0.0, -3.0]], [0.5, 0.3, 0.2]
    case("mixture_d2", LO.mixture_energy(centers, weights), {"kind": "mixture", "centers": np.array(centers), "weights": np.array(weights)},
         2, [0.1, 0.2], 4, 8, 12, 1.0, 0.02, 1.0, 34)
    print("langevin goldens written")


if __name__ == "__main__":
    main()

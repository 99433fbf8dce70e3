"""CPU: dense-J and Langevin oracles against the golden vectors produced by the unmodified reference."""
import glob
import os

import numpy as np
import pytest

from oracle import dense_oracle as D
from oracle import langevin_oracle as LO


def load(golden_dir, pattern):
    paths = sorted(glob.glob(os.path.join(golden_dir, pattern)))
    assert paths, pattern
    return [(p, np.load(p, allow_pickle=False)) for p in paths]


def test_dense_sweep_goldens(golden_dir):
    for path, g in load(golden_dir, "dense_sweep_*.npz"):
        orders = g["orders"] if str(g["order_mode"]) == "random" else None
        out = D.gibbs_sweeps(g["s0"], g["J"], g["b"], float(g["T"]), len(g["uniforms"]), g["uniforms"], orders)
        assert (out == g["out"]).all(), path
        assert D.compute_energy(out, g["J"], g["b"]) == pytest.approx(float(g["energy"]), abs=1e-12)


def test_dense_boltzmann_golden(golden_dir):
    for path, g in load(golden_dir, "dense_boltzmann_*.npz"):
        out = D.sample_boltzmann(g["J"], g["b"], float(g["T"]), int(g["burnin"]), int(g["n_samples"]), int(g["n_sweeps"]),
                                 g["s0"], g["uniforms"])
        assert (out == g["samples"]).all(), path


def test_dense_annealing_goldens(golden_dir):
    for path, g in load(golden_dir, "dense_anneal_*.npz"):
        best, e = D.simulated_annealing(g["J"], g["b"], 5.0, 0.2, int(g["n_steps"]), str(g["schedule"]), g["s0"], g["uniforms"])
        assert (best == g["best_state"]).all(), path
        assert e == pytest.approx(float(g["best_energy"]), abs=1e-12)


def test_dense_tempering_golden(golden_dir):
    for path, g in load(golden_dir, "dense_tempering_*.npz"):
        samples, info = D.parallel_tempering(g["J"], g["b"], list(g["temps"]), int(g["burnin"]), int(g["n_sweeps"]),
                                             int(g["n_samples"]), int(g["swap_interval"]), g["inits"], g["burn_uniforms"],
                                             g["sweep_uniforms"], g["swap_uniforms"])
        assert (samples == g["samples"]).all(), path
        assert info["swap_attempts"] == int(g["swap_attempts"]) and info["swap_accepts"] == int(g["swap_accepts"])
        assert np.allclose(np.array(info["energies"]), g["energies"], atol=1e-12)
        assert (np.array(info["final_states"]) == g["final_states"]).all()


def energy_from_golden(g):
    kind = str(g["kind"])
    if kind == "quadratic":
        return LO.quadratic_energy
    if kind == "gaussian":
        return LO.gaussian_energy(g["mu"], g["sigma"])
    return LO.mixture_energy(g["centers"], g["weights"])


def test_langevin_goldens(golden_dir):
    for path, g in load(golden_dir, "langevin_*.npz"):
        samples, traj = LO.sample_from_energy(energy_from_golden(g), g["x_init"], int(g["n_samples"]), g["normals"],
                                              float(g["T"]), float(g["dt"]), float(g["friction"]), int(g["n_burnin"]),
                                              int(g["n_steps"]), return_trajectory=True)
        assert np.array_equal(samples, g["samples"]), path
        assert np.array_equal(np.array(traj), g["trajectory"]), path


def test_known_answers_from_reference_tests():
    # tsu/tests/test_gibbs.py:47-61,131-145 known answers
    J = np.array([[0, 1, 2], [1, 0, 1], [2, 1, 0]], dtype=float)
    s = np.array([1, 0, 1])
    assert float(np.dot(J[0], s)) == 2.0
    assert D.compute_energy(s, J) == -2.0
    assert D.compute_energy(s, J, np.array([1.0, 1.0, 1.0])) == -4.0
    assert D.sigmoid_ref(100) == 1.0 and D.sigmoid_ref(-100) == 0.0 and abs(D.sigmoid_ref(0) - 0.5) < 1e-6

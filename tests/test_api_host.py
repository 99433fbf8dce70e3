"""CPU: host-side mirror of the reference API - validation messages, scalar helpers, known answers
(modelled on the reference's tests/test_gibbs.py, test_ising.py, test_core.py; no GPU needed)."""
import numpy as np
import pytest

from tsu_emulator_b200 import (ConfigurationError, GibbsConfig, GibbsSampler, HardwareEmulator, IsingConfig,
                               SamplingError, TSUConfig, TSUError)
from tsu_emulator_b200.core import GaussianEnergy, MixtureEnergy, QuadraticEnergy, recognise_quadratic


def test_gibbs_config_validation_messages():
    with pytest.raises(ValueError, match="Temperature must be positive"):
        GibbsConfig(temperature=-1.0)
    with pytest.raises(ValueError, match="Burn-in steps must be non-negative"):
        GibbsConfig(n_burnin=-1)
    with pytest.raises(ValueError, match="Number of sweeps must be positive"):
        GibbsConfig(n_sweeps=0)
    with pytest.raises(ValueError, match="Update order must be"):
        GibbsConfig(update_order="bogus")
    c = GibbsConfig()
    assert (c.temperature, c.n_burnin, c.n_sweeps, c.update_order) == (1.0, 100, 10, "sequential")


def test_tsu_config_validation():
    for kw in ({"temperature": -1.0}, {"dt": -0.01}, {"dt": 1.0}, {"n_steps": -10}, {"friction": 0.0}, {"n_burnin": -1}):
        with pytest.raises(ConfigurationError):
            TSUConfig(**kw)
    assert issubclass(ConfigurationError, TSUError) and issubclass(SamplingError, TSUError)
    assert TSUConfig(temperature=1.0, dt=0.01, n_steps=100).temperature == 1.0


def test_ising_config_validation():
    with pytest.raises(ValueError, match="Temperature must be positive"):
        IsingConfig(temperature=0.0)


def test_sigmoid_and_local_field_known_answers():
    s = GibbsSampler(seed=0)
    assert abs(s._sigmoid(0) - 0.5) < 1e-6
    assert s._sigmoid(10) > 0.99 and s._sigmoid(-10) < 0.01
    assert s._sigmoid(100) == 1.0 and s._sigmoid(-100) == 0.0   # the +-20 clamp
    state = np.array([1, 0, 1])
    J = np.array([[0, 1, 2], [1, 0, 1], [2, 1, 0]], dtype=float)
    assert s._compute_local_field(0, state, J) == 2.0
    assert s._compute_local_field(0, state, J, np.array([0.5, -0.5, 1.0])) == 2.5
    assert s.compute_energy(state, J) == -2.0
    assert s.compute_energy(state, J, np.array([1.0, 1.0, 1.0])) == -4.0


def test_hardware_emulator_timing_arithmetic():
    hw = HardwareEmulator(n_bits=100, clock_speed_ghz=1.0, parallel_chains=1000)
    t = hw.estimate_hardware_time(n_samples=10000, n_sweeps_per_sample=10)
    assert t["time_per_sweep_ns"] == 100.0 and t["batches_needed"] == 10
    assert t["total_time_ns"] == 10 * 10 * 100.0 and t["speedup_vs_classical"] is None


def test_quadratic_recognition():
    q = recognise_quadratic(lambda x: (x**2).sum(), 4, np.zeros(4))
    d = q.as_diagonal()
    def canon(p, dim):  # (a * w, mu): the parametrisation-independent content
        return np.concatenate([p[0] * p[1 + dim:], p[1:1 + dim]])
    assert d is not None and np.allclose(canon(d.params(4), 4), canon(QuadraticEnergy().params(4), 4))
    g = recognise_quadratic(lambda x: 0.5 * ((float(np.atleast_1d(x)[0]) - 5.0) / 2.0) ** 2, 1, np.array([5.0]))
    assert np.allclose(canon(g.as_diagonal().params(1), 1), canon(GaussianEnergy(5.0, 2.0).params(1), 1))
    assert recognise_quadratic(lambda x: float(np.sum(np.abs(x) ** 3)), 2, np.zeros(2)) is None
    A = np.array([[2.0, 0.5], [0.5, 1.0]])
    c = recognise_quadratic(lambda x: 0.5 * x @ A @ x - x[0], 2, np.array([0.3, -0.2]))
    assert c.as_diagonal() is None and np.allclose(c.A, A) and np.allclose(c.b, [1.0, 0.0])


def test_builtin_energy_values_match_reference_formulas():
    from oracle import langevin_oracle as LO
    x = np.array([0.3, -1.2])
    assert QuadraticEnergy()(x) == pytest.approx(LO.quadratic_energy(x))
    m = MixtureEnergy([[-2.0, 0.0], [2.0, 1.0]], [0.6, 0.4])
    assert m(x) == pytest.approx(LO.mixture_energy([[-2.0, 0.0], [2.0, 1.0]], [0.6, 0.4])(x))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tsu_emulator_b200 import IsingModel2D, ThermalSamplingUnit
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        IsingModel2D(size=8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GibbsSampler(seed=0).sample_boltzmann(np.zeros((3, 3)), n_samples=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ThermalSamplingUnit(seed=0).sample_gaussian(0.0, 1.0, 4)


def test_compat_shim_exposes_reference_module_paths():
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tsu_emulator_b200", "compat"))
    try:
        import importlib
        for mod, names in {"tsu.gibbs": ["GibbsSampler", "GibbsConfig", "HardwareEmulator"],
                           "tsu.core": ["ThermalSamplingUnit", "TSUConfig", "ConfigurationError"],
                           "tsu.models.ising": ["IsingModel", "IsingGrid", "IsingChain", "IsingModel2D", "IsingConfig"]}.items():
            m = importlib.import_module(mod)
            for n in names:
                assert hasattr(m, n)
    finally:
        sys.path.pop(0)
        for k in [k for k in sys.modules if k == "tsu" or k.startswith("tsu.")]:
            del sys.modules[k]


def test_p_bit_and_categorical_validation_precedes_device_work():
    """tsu/core.py:177-180: argument errors are ConfigurationError, raised before anything touches the GPU"""
    from tsu_emulator_b200 import ConfigurationError, ProbabilisticNeuron, ThermalSamplingUnit, validate_distribution

    tsu = ThermalSamplingUnit(seed=1)
    for bad in (-0.1, 1.5):
        with pytest.raises(ConfigurationError, match="Probability must be in"):
            tsu.p_bit(prob=bad)
    with pytest.raises(ConfigurationError, match="n_samples must be positive"):
        tsu.p_bit(prob=0.5, n_samples=0)
    with pytest.raises(ConfigurationError):
        tsu.sample_categorical(np.array([]), 3)
    with pytest.raises(ConfigurationError):
        tsu.sample_categorical(np.array([0.5, -0.1]), 3)
    assert ProbabilisticNeuron(tsu).tsu is tsu
    rng = np.random.default_rng(0)
    r = validate_distribution(rng.normal(2.0, 3.0, 4000), "gaussian", {"mu": 2.0, "sigma": 3.0})
    assert r["passes_ks_test"] and abs(r["mean"] - 2.0) < 0.2 and r["expected_std"] == 3.0
    r = validate_distribution(rng.random(4000) < 0.3, "bernoulli", {"p": 0.3})
    assert r["passes_test"] and abs(r["empirical_prob"] - 0.3) < 0.03

"""GPU parity: fused Langevin kernel (through ThermalSamplingUnit -> C-ABI) against the reference goldens."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def load(golden_dir):
    paths = sorted(glob.glob(os.path.join(golden_dir, "langevin_*.npz")))
    assert len(paths) >= 4
    return [(p, np.load(p, allow_pickle=False)) for p in paths]


def energy_obj(g):
    from tsu_emulator_b200 import GaussianEnergy, MixtureEnergy, QuadraticEnergy

    kind = str(g["kind"])
    if kind == "quadratic":
        return QuadraticEnergy()
    if kind == "gaussian":
        return GaussianEnergy(g["mu"], g["sigma"])
    return MixtureEnergy(g["centers"], g["weights"])


def run(g, dtype):
    from tsu_emulator_b200 import ThermalSamplingUnit, TSUConfig

    cfg = TSUConfig(temperature=float(g["T"]), dt=float(g["dt"]), friction=float(g["friction"]),
                    n_burnin=int(g["n_burnin"]), n_steps=int(g["n_steps"]))
    tsu = ThermalSamplingUnit(cfg, dtype=dtype, seed=1)
    return tsu.sample_from_energy(energy_obj(g), g["x_init"], int(g["n_samples"]), return_trajectory=True,
                                  _normals=g["normals"])


# tolerance: the reference differentiates numerically (central difference, eps=1e-5, error ~1e-10 relative) and
# runs in float64; the kernel uses the analytic gradient.  float64 kernel: 1e-8; float32 kernel: 2e-4.
def test_goldens_float64(golden_dir):
    for path, g in load(golden_dir):
        samples, traj = run(g, "float64")
        assert samples.shape == g["samples"].shape
        assert np.allclose(samples, g["samples"], rtol=0, atol=1e-8), path
        assert np.allclose(np.array(traj), g["trajectory"], rtol=0, atol=1e-8), path


def test_goldens_float32(golden_dir):
    for path, g in load(golden_dir):
        samples, traj = run(g, "float32")
        assert np.allclose(samples, g["samples"], rtol=0, atol=2e-4), path
        assert np.allclose(np.array(traj), g["trajectory"], rtol=0, atol=2e-4), path


def test_readme_callable_is_recognised_and_matches_builtin(golden_dir):
    from tsu_emulator_b200 import ThermalSamplingUnit, TSUConfig

    g = dict(np.load(os.path.join(golden_dir, "langevin_quadratic_d3.npz")))
    cfg = TSUConfig(n_burnin=int(g["n_burnin"]), n_steps=int(g["n_steps"]))
    tsu = ThermalSamplingUnit(cfg, dtype="float64", seed=1)
    out = tsu.sample_from_energy(lambda x: (x**2).sum(), g["x_init"], int(g["n_samples"]), _normals=g["normals"])
    assert np.allclose(out, g["samples"], atol=1e-8)
    # coupled quadratic form goes through the QUADRATIC_FORM path
    A = np.array([[2.0, 0.5, 0.0], [0.5, 1.0, 0.2], [0.0, 0.2, 3.0]])
    out2 = tsu.sample_from_energy(lambda x: 0.5 * x @ A @ x, g["x_init"], int(g["n_samples"]), _normals=g["normals"])
    from oracle import langevin_oracle as LO
    want = LO.sample_from_energy(lambda x: 0.5 * x @ A @ x, g["x_init"], int(g["n_samples"]), g["normals"],
                                 n_burnin=int(g["n_burnin"]), n_steps=int(g["n_steps"]))
    assert np.allclose(out2, want, atol=1e-7)


def test_unrecognised_callable_raises():
    from tsu_emulator_b200 import SamplingError, ThermalSamplingUnit

    tsu = ThermalSamplingUnit()
    with pytest.raises(SamplingError):
        tsu.sample_from_energy(lambda x: float(np.sum(np.abs(x) ** 3)), np.zeros(2), 4)
    with pytest.raises(SamplingError):
        tsu.sample_from_energy(lambda x: 1.0, np.zeros(2), 0)


def test_statistics_gaussian_philox():
    """tsu/tests/test_core.py:49-74: mean and std of sample_gaussian, KS against N(0,1)"""
    from scipy import stats
    from tsu_emulator_b200 import ThermalSamplingUnit, TSUConfig

    tsu = ThermalSamplingUnit(TSUConfig(n_steps=300), seed=5)
    s = tsu.sample_gaussian(mu=5.0, sigma=1.0, n_samples=20000)
    assert abs(s.mean() - 5.0) < 0.05
    s = tsu.sample_gaussian(mu=0.0, sigma=2.0, n_samples=20000)
    # 400 steps of dt=0.01 is one relaxation time for sigma=2: var = sigma^2 (1 - exp(-2 t / sigma^2)) -> std 1.86
    # (the reference's own tolerance is 0.3, tests/test_core.py:58-66)
    assert abs(s.std() - 2.0) < 0.3
    assert abs(s.std() - np.sqrt(4.0 * (1 - np.exp(-2 * 4.0 / 4.0)))) < 0.05
    s = tsu.sample_gaussian(mu=0.0, sigma=1.0, n_samples=1000)
    assert stats.kstest(s, "norm")[1] > 0.01


def test_readme_sample_boltzmann_shape_and_variance():
    """README.md:46-64; Euler-Maruyama stationary variance of E = sum x^2 is T / (2 (1 - dt)) = 0.505"""
    from tsu_emulator_b200 import ThermalSamplingUnit, TSUConfig

    tsu = ThermalSamplingUnit(TSUConfig(temperature=1.0, dt=0.01, friction=1.0, n_burnin=100, n_steps=500), seed=9)
    out = tsu.sample_boltzmann(lambda x: (x**2).sum(), n_samples=100000, dim=10)
    assert out.shape == (100000, 10) and out.dtype == np.float64
    assert abs(out.var() - 0.505) < 0.01
    assert abs(out.mean()) < 0.01
    assert tsu.sample_count == 100000


def test_p_bit_categorical_and_neuron():
    """tsu/core.py:164-206,241-294 contracts (shape, binary, probability within the reference test's 0.05, errors);
    exact Bernoulli / inverse-CDF draws on Philox words"""
    from tsu_emulator_b200 import ConfigurationError, ProbabilisticNeuron, ThermalSamplingUnit, validate_distribution

    tsu = ThermalSamplingUnit(seed=5)
    s = tsu.p_bit(prob=0.5, n_samples=100)
    assert len(s) == 100 and set(s).issubset({0, 1})
    for p in (0.2, 0.5, 0.8):                      # reference tests/test_core.py:86-97
        assert abs(tsu.p_bit(prob=p, n_samples=20000).mean() - p) < 0.015
    assert tsu.p_bit(0.0, 1000).sum() == 0 and tsu.p_bit(1.0, 1000).sum() == 1000
    a, b = tsu.p_bit(0.5, 64), tsu.p_bit(0.5, 64)
    assert (a != b).any()                          # successive calls use fresh random words
    with pytest.raises(ConfigurationError):
        tsu.p_bit(prob=-0.1)
    with pytest.raises(ConfigurationError):
        tsu.p_bit(prob=0.5, n_samples=0)
    probs = np.array([1.0, 2.0, 3.0, 4.0])
    c = tsu.sample_categorical(probs, n_samples=40000)
    assert c.min() >= 0 and c.max() <= 3
    assert np.abs(np.bincount(c, minlength=4) / 40000 - probs / 10).max() < 0.01
    n = ProbabilisticNeuron(tsu)
    assert n.activate(np.array([10.0]), np.array([5.0])) == 1
    assert abs(n.forward_stochastic(np.array([1.0, -1.0]), np.array([0.3, 0.3]), n_samples=4000) - 0.5) < 0.04
    r = validate_distribution(tsu.p_bit(0.3, 5000), "bernoulli", {"p": 0.3})
    assert r["passes_test"] and r["n_samples"] == 5000
    g = ThermalSamplingUnit(seed=6).sample_gaussian(mu=0, sigma=1, n_samples=2000)
    assert validate_distribution(g, "gaussian", {"mu": 0, "sigma": 1}, alpha=0.001)["passes_ks_test"]


def _oracle_vs_kernel(energy, energy_fn, x_init, n_samples, cfg_kw, atol, dtype="float64", seed=3):
    """injected normals through the fused kernel and through the oracle's restatement of core.py:100-162 (which
    differentiates numerically, eps = 1e-5, like the reference)"""
    from oracle import langevin_oracle as LO
    from tsu_emulator_b200 import ThermalSamplingUnit, TSUConfig

    cfg = TSUConfig(**cfg_kw)
    x_init = np.atleast_1d(np.asarray(x_init, dtype=np.float64))
    rng = np.random.default_rng(seed)
    normals = rng.normal(size=(n_samples, 1 + cfg.n_burnin + cfg.n_steps, x_init.size))
    got, traj = ThermalSamplingUnit(cfg, dtype=dtype, seed=1).sample_from_energy(
        energy, x_init, n_samples, return_trajectory=True, _normals=normals)
    want, wtraj = LO.sample_from_energy(energy_fn, x_init, n_samples, normals, temperature=cfg.temperature, dt=cfg.dt,
                                        friction=cfg.friction, n_burnin=cfg.n_burnin, n_steps=cfg.n_steps,
                                        return_trajectory=True)
    assert np.allclose(got, want, rtol=0, atol=atol), np.abs(got - want).max()
    assert np.allclose(np.array(traj), np.array(wtraj), rtol=0, atol=atol)


def test_double_well_energy_matches_reference_loop():
    """DoubleWellEnergy E = sum a (x^2 - b)^2 (bimodal landscape, wells at +-sqrt(b)): analytic gradient in
    the kernel against the numerically differentiated reference loop, float64 1e-7 / float32 5e-4 (tolerances: the
    central difference of a quartic has a relative error ~eps^2 x''' ~ 1e-9 per step, accumulated over 60 steps)"""
    from tsu_emulator_b200 import DoubleWellEnergy

    a, b = 0.7, 1.3
    fn = lambda x: float(np.sum(a * (np.asarray(x) ** 2 - b) ** 2))
    e = DoubleWellEnergy(a, b)
    assert e(np.array([0.3, -1.2])) == pytest.approx(fn(np.array([0.3, -1.2])), abs=1e-12)
    kw = dict(temperature=0.8, dt=0.02, friction=1.5, n_burnin=20, n_steps=40)
    _oracle_vs_kernel(e, fn, [0.2, -0.4, 1.1], 6, kw, 1e-7)
    _oracle_vs_kernel(e, fn, [0.2, -0.4, 1.1], 6, kw, 5e-4, dtype="float32")
    # by name (a = b = 1), and statistically: the two wells at +-1 are both populated
    from tsu_emulator_b200 import ThermalSamplingUnit, TSUConfig
    s = ThermalSamplingUnit(TSUConfig(temperature=0.5, n_burnin=200, n_steps=800), seed=2).sample_from_energy(
        "double_well", np.zeros(1), 4000)
    assert 0.35 < (s > 0).mean() < 0.65 and abs(np.abs(s).mean() - 1.0) < 0.15


def test_mean_reduced_gaussian_energy_matches_reference_loop():
    """GaussianEnergy(reduce='mean') = the energy of the reference's GaussianSampler (api.py:124-126: the MEAN over the
    coordinates, so each coordinate feels 1/dim of the force)"""
    from tsu_emulator_b200 import GaussianEnergy

    mu, sigma = np.array([1.0, -2.0, 0.5]), np.array([0.5, 2.0, 1.0])
    fn = lambda x: float(np.mean(0.5 * ((np.asarray(x) - mu) / sigma) ** 2))
    e = GaussianEnergy(mu, sigma, reduce="mean")
    kw = dict(temperature=1.0, dt=0.05, friction=1.0, n_burnin=10, n_steps=50)
    _oracle_vs_kernel(e, fn, [0.0, 0.0, 0.0], 5, kw, 1e-8)
    _oracle_vs_kernel(e, fn, [0.0, 0.0, 0.0], 5, kw, 2e-4, dtype="float32")


def test_api_samplers_on_the_b200_backend():
    """tsu/api.py:38-218 facade: Backend.B200 (and EMULATOR as its alias), GaussianSampler / MultimodalSampler /
    BayesianSampler.sample, SamplingResult metadata, the functional helpers; unknown backends raise like the reference"""
    from tsu_emulator_b200 import SamplingError, TSUConfig
    from tsu_emulator_b200.api import (Backend, BayesianSampler, GaussianSampler, MultimodalSampler, SamplingResult,
                                       sample_gaussian, sample_multimodal)

    cfg = TSUConfig(n_burnin=200, n_steps=800)
    g = GaussianSampler(mu=3.0, sigma=0.5, config=cfg, seed=1)
    assert g.backend == Backend.B200
    s = g.sample(4000)
    assert s.shape == (4000, 1) and abs(s.mean() - 3.0) < 0.05 and abs(s.std() - 0.5) < 0.05
    r = GaussianSampler(mu=0.0, sigma=1.0, backend=Backend.EMULATOR, config=cfg, seed=2).sample(100, return_metadata=True)
    assert isinstance(r, SamplingResult) and r.samples.shape == (100, 1) and r.time_elapsed > 0 and r.backend_used == "b200"
    with pytest.raises(NotImplementedError):
        GaussianSampler(backend=Backend.CLOUD).sample(10)
    m = MultimodalSampler(centers=[[-3.0, 0.0], [3.0, 0.0]], weights=[1.0, 1.0], config=cfg, seed=3)
    x = m.sample(4000)
    assert x.shape == (4000, 2)
    # every chain starts at the sampler's one initial state (api.py:151-152) + 0.1 N(0, I) and relaxes into a mode
    near = np.minimum(np.abs(x[:, 0] + 3.0), np.abs(x[:, 0] - 3.0))
    assert near.mean() < 1.2 and np.abs(x[:, 1]).mean() < 1.2
    assert sample_gaussian(1.0, 2.0, n=50).shape == (50, 1) and sample_multimodal([[0.0], [4.0]], [1, 1], n=20).shape == (20, 1)
    # linear-Gaussian posterior (the reference's docstring example) is a quadratic form: recognised and exact
    rng = np.random.default_rng(0)
    X = rng.normal(size=(30, 2)); y = X @ np.array([1.5, -0.5]) + 0.1 * rng.normal(size=30)
    b = BayesianSampler(lambda th, X_, y_: -0.5 * np.sum((y_ - X_ @ th) ** 2), lambda th: -0.5 * np.sum(th ** 2), X, y,
                        dim=2, config=TSUConfig(dt=0.005, n_burnin=400, n_steps=1200), seed=4)
    post = b.sample(3000)
    cov = np.linalg.inv(X.T @ X + np.eye(2)); mean = cov @ X.T @ y
    assert np.abs(post.mean(0) - mean).max() < 0.05
    # a non-quadratic posterior is traced and compiled; one whose control flow depends on theta cannot run on the device
    heavy = BayesianSampler(lambda th: -np.sum(np.abs(th) ** 3), lambda th: -0.5 * np.sum(th ** 2), dim=2, config=cfg, seed=5)
    hs = heavy.sample(2000)
    assert hs.shape == (2000, 2) and abs(hs.mean()) < 0.1 and 0.2 < hs.std() < 0.8
    with pytest.raises(SamplingError):
        BayesianSampler(lambda th: -1.0 if th[0] > 0 else -2.0, lambda th: 0.0, dim=2).sample(4)


def test_traced_python_energies_match_the_reference_loop():
    """arbitrary Python energies (core.py:100-162 accepts any callable): traced, differentiated analytically, compiled by
    NVRTC into the chain loop; against the numerically differentiated reference loop on injected normals.
    Tolerance 1e-6 (float64): the reference's central difference (eps = 1e-5) of non-polynomial energies carries a
    relative error ~1e-10 per step that the stiff Rosenbrock valley amplifies over 60 steps."""
    centers = [np.array([-2.0, 0.0, 1.0]), np.array([2.0, 1.0, 0.0])]
    weights = [0.3, 0.7]

    def mixture(x):  # the loop of tsu/api.py:143-149
        x = np.atleast_1d(x)
        prob = 0
        for i, c in enumerate(centers):
            prob += weights[i] * np.exp(-0.5 * np.sum((x - c) ** 2))
        return -np.log(prob + 1e-10)

    cases = {
        "double well + tilt": lambda x: np.sum(x ** 4) - 2.0 * np.sum(x ** 2) + 0.3 * x[0] * x[1],
        "mixture loop": mixture,
        "rosenbrock-like": lambda x: 2.0 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2 + np.cos(x[2]) * np.tanh(x[1])
                                     + np.sqrt(1 + x[2] ** 2),
        "abs + matmul": (lambda A: (lambda x: 0.5 * x @ A @ x + 0.2 * np.sum(np.abs(x)) ** 1.5))(
            np.array([[2.0, 0.5, 0.0], [0.5, 1.0, 0.2], [0.0, 0.2, 3.0]])),
    }
    kw = dict(temperature=0.7, dt=0.01, friction=1.0, n_burnin=20, n_steps=40)
    for name, fn in cases.items():
        _oracle_vs_kernel(fn, lambda x, f=fn: float(f(np.asarray(x, dtype=np.float64))), [0.3, -0.2, 0.5], 5, kw, 1e-6)
    _oracle_vs_kernel(cases["double well + tilt"], lambda x: float(cases["double well + tilt"](np.asarray(x))), [0.3, -0.2, 0.5],
                      5, kw, 1e-3, dtype="float32")
    # the traced mixture and the built-in MixtureEnergy are the same function: same chains
    from tsu_emulator_b200 import MixtureEnergy, ThermalSamplingUnit, TSUConfig
    a = ThermalSamplingUnit(TSUConfig(**kw), seed=9).sample_from_energy(mixture, np.zeros(3), 64)
    b = ThermalSamplingUnit(TSUConfig(**kw), seed=9).sample_from_energy(MixtureEnergy(np.stack(centers), weights), np.zeros(3), 64)
    assert np.allclose(a, b, atol=1e-9)
    # statistics of a traced double well: both wells at +-1 populated
    s_ = ThermalSamplingUnit(TSUConfig(temperature=0.5, n_burnin=200, n_steps=800), seed=2).sample_from_energy(
        lambda x: np.sum((x ** 2 - 1.0) ** 2), np.zeros(1), 4000)
    assert 0.35 < (s_ > 0).mean() < 0.65 and abs(np.abs(s_).mean() - 1.0) < 0.15

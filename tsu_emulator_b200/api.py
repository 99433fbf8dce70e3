"""
Sampler facade of the reference's tsu/api.py (Backend, SamplingResult, Sampler, GaussianSampler,
MultimodalSampler, BayesianSampler, sample_gaussian, sample_multimodal: api.py:38-212) with a B200 backend.

The reference's samplers hand their `energy_function` to ThermalSamplingUnit.sample_from_energy, which
differentiates it numerically in Python (api.py:92-100, core.py:82-98).  Here every sampler names the analytic
energy it stands for and the chains run on the fused Langevin kernel (csrc/langevin.cu):
    GaussianSampler     E = mean(0.5 ((x - mu) / sigma)^2)         api.py:124-126  -> GaussianEnergy(reduce="mean")
    MultimodalSampler   E = -log(sum_k w_k exp(-|x - c_k|^2 / 2) + 1e-10)   api.py:143-149  -> MixtureEnergy
    BayesianSampler     E = -(log_likelihood + log_prior)          api.py:187-191  -> recognised if it is a quadratic
                                                                   form in theta (linear-Gaussian models), else SamplingError
Backend.B200 is the default; Backend.EMULATOR (the reference's only working value) is accepted and means the same
engine, so code written against the reference runs unchanged.  CLOUD / HARDWARE / HYBRID raise NotImplementedError
exactly as in the reference (api.py:100).
"""

import time
from abc import ABC, abstractmethod
from dataclasses import dataclass
from enum import Enum
from typing import Callable, Dict, List, Optional, Union

import numpy as np

from .core import GaussianEnergy, MixtureEnergy, ThermalSamplingUnit, TSUConfig


class Backend(Enum):
    """tsu/api.py:38-44 plus the backend this package adds"""

    EMULATOR = "emulator"
    CLOUD = "cloud"
    HARDWARE = "hardware"
    HYBRID = "hybrid"
    B200 = "b200"  # fused CUDA kernels on an NVIDIA B200 (sm_100a)


@dataclass
class SamplingResult:
    """tsu/api.py:47-56"""

    samples: np.ndarray
    energy: Optional[np.ndarray] = None
    acceptance_rate: Optional[float] = None
    time_elapsed: Optional[float] = None
    backend_used: str = "b200"
    hardware_projection: Optional[Dict] = None


class Sampler(ABC):
    """tsu/api.py:59-110: same constructor, `sample(n, return_metadata)` and abstract hooks"""

    def __init__(self, backend: Backend = Backend.B200, config: Optional[TSUConfig] = None, *, seed: Optional[int] = None,
                 dtype: str = "float64"):
        self.backend = backend
        self.config = config or TSUConfig()
        self._tsu = ThermalSamplingUnit(self.config, seed=seed, dtype=dtype)

    @abstractmethod
    def energy_function(self, x: np.ndarray) -> float:
        """Energy function defining the distribution"""

    @abstractmethod
    def _get_initial_state(self) -> np.ndarray:
        """Get initial state for sampling"""

    def _device_energy(self):
        """what the fused kernel evaluates; default: the Python energy function, which must be recognisable"""
        return self.energy_function

    def sample(self, n: int = 1000, return_metadata: bool = False) -> Union[np.ndarray, SamplingResult]:
        start = time.time()
        if self.backend in (Backend.B200, Backend.EMULATOR):
            x_init = self._get_initial_state()
            samples = self._tsu.sample_from_energy(self._device_energy(), x_init, n_samples=n)
        else:
            raise NotImplementedError(f"Backend {self.backend} not yet implemented")
        elapsed = time.time() - start
        if return_metadata:
            return SamplingResult(samples=samples, time_elapsed=elapsed, backend_used=Backend.B200.value)
        return samples


class GaussianSampler(Sampler):
    """tsu/api.py:113-129"""

    def __init__(self, mu: float = 0.0, sigma: float = 1.0, **kwargs):
        super().__init__(**kwargs)
        self.mu = mu
        self.sigma = sigma

    def energy_function(self, x: np.ndarray) -> float:
        result = 0.5 * ((x - self.mu) / self.sigma) ** 2
        return float(np.mean(result)) if isinstance(result, np.ndarray) else float(result)

    def _device_energy(self):
        return GaussianEnergy(self.mu, self.sigma, reduce="mean")

    def _get_initial_state(self) -> np.ndarray:
        return np.array([self.mu])


class MultimodalSampler(Sampler):
    """tsu/api.py:132-152"""

    def __init__(self, centers: Union[List[np.ndarray], List[List]], weights: List[float], **kwargs):
        super().__init__(**kwargs)
        self.centers = [np.array(c) for c in centers]
        self.weights = np.array(weights) / np.sum(weights)
        self.dim = len(self.centers[0])

    def energy_function(self, x: np.ndarray) -> float:
        x = np.atleast_1d(x)
        prob = 0
        for i, center in enumerate(self.centers):
            dist_sq = np.sum((x - center) ** 2)
            prob += self.weights[i] * np.exp(-0.5 * dist_sq)
        return -np.log(prob + 1e-10)

    def _device_energy(self):
        return MixtureEnergy(np.stack(self.centers), self.weights)

    def _get_initial_state(self) -> np.ndarray:
        return np.random.randn(self.dim) * 0.5


class BayesianSampler(Sampler):
    """tsu/api.py:155-197.  The posterior energy is an arbitrary Python callable; it runs here when it is a quadratic
    form in theta (Gaussian likelihood of a linear model + Gaussian prior, the docstring example of the reference)."""

    def __init__(self, log_likelihood: Callable, log_prior: Callable, *args, dim: Optional[int] = None, **kwargs):
        super().__init__(**kwargs)
        self.log_likelihood = log_likelihood
        self.log_prior = log_prior
        self.likelihood_args = args
        self.dim = dim

    def energy_function(self, theta: np.ndarray) -> float:
        theta = np.atleast_1d(theta)
        log_lik = self.log_likelihood(theta, *self.likelihood_args)
        log_pri = self.log_prior(theta)
        return -(log_lik + log_pri)

    def _get_initial_state(self) -> np.ndarray:
        if self.dim is None:
            raise ValueError("Must specify dimension for Bayesian sampler")
        return np.random.randn(self.dim) * 0.1


def sample_gaussian(mu: float = 0, sigma: float = 1, n: int = 1000, backend: Backend = Backend.B200) -> np.ndarray:
    """tsu/api.py:203-209"""
    return GaussianSampler(mu, sigma, backend=backend).sample(n, return_metadata=False)


def sample_multimodal(centers: List, weights: List, n: int = 1000, backend: Backend = Backend.B200) -> np.ndarray:
    """tsu/api.py:212-218"""
    return MultimodalSampler(centers, weights, backend=backend).sample(n, return_metadata=False)

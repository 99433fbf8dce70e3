"""
Import shim: put `<repo>/tsu_emulator_b200/compat` on PYTHONPATH and code written against the
reference (`from tsu.gibbs import GibbsSampler`, `from tsu.models.ising import IsingModel2D`,
`from tsu.core import ThermalSamplingUnit`) runs on the B200 engine unchanged.  Only the hot-path
names of tsu/__init__.py:11-37 exist here (no ml / visualization / api / benchmarks).
"""
from tsu_emulator_b200 import (  # noqa: F401
    TSU,
    ConfigurationError,
    GibbsConfig,
    GibbsSampler,
    HardwareEmulator,
    IsingChain,
    IsingGrid,
    IsingModel,
    SamplingError,
    ThermalSamplingUnit,
    TSUConfig,
    TSUError,
    demonstrate_phase_transition,
)

__version__ = "0.1.0+b200"

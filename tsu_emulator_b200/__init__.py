"""
tsu_emulator_b200 - B200-native sampling engine behind tsu-emulator's Python API.

Hot path only: heat-bath Gibbs spin updates (bit-packed 2-D lattices and dense couplings) and batched
Langevin steps, as hand-written sm_100a CUDA behind a C-ABI library (include/tsu_b200.h).  The names
below mirror what `tsu/__init__.py:11-37` re-exports for that path.  There is no CPU fallback.
"""

__version__ = "0.1.0"

from .core import (  # noqa: F401
    TSU,
    ConfigurationError,
    DoubleWellEnergy,
    GaussianEnergy,
    MixtureEnergy,
    ProbabilisticNeuron,
    QuadraticEnergy,
    QuadraticFormEnergy,
    SamplingError,
    ThermalSamplingUnit,
    TSUConfig,
    TSUError,
    validate_distribution,
)
from .gibbs import GibbsConfig, GibbsSampler, HardwareEmulator  # noqa: F401
from .lattice import Ising2DEngine, build_lut  # noqa: F401
from .models import (  # noqa: F401
    IsingChain,
    IsingConfig,
    IsingGrid,
    IsingModel,
    IsingModel2D,
    demonstrate_phase_transition,
)
from .api import (  # noqa: F401,E402
    Backend,
    BayesianSampler,
    GaussianSampler,
    MultimodalSampler,
    Sampler,
    SamplingResult,
)

from tsu_emulator_b200.core import *  # noqa: F401,F403
from tsu_emulator_b200.core import (ConfigurationError, ProbabilisticNeuron, SamplingError, ThermalSamplingUnit, TSUConfig, TSUError,
                                    validate_distribution)  # noqa: F401

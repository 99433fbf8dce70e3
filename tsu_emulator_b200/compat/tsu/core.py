from tsu_emulator_b200.core import *  # noqa: F401,F403
from tsu_emulator_b200.core import (ConfigurationError, SamplingError, ThermalSamplingUnit, TSUConfig, TSUError)  # noqa: F401

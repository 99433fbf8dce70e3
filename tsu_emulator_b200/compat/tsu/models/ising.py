from tsu_emulator_b200.models.ising import *  # noqa: F401,F403
from tsu_emulator_b200.models.ising import (IsingChain, IsingConfig, IsingGrid, IsingModel, IsingModel2D, demonstrate_phase_transition)  # noqa: F401

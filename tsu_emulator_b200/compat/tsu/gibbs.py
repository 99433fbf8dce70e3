from tsu_emulator_b200.gibbs import *  # noqa: F401,F403
from tsu_emulator_b200.gibbs import GibbsConfig, GibbsSampler, HardwareEmulator  # noqa: F401

from tsu_emulator_b200.api import *  # noqa: F401,F403
from tsu_emulator_b200.api import (  # noqa: F401
    Backend,
    BayesianSampler,
    GaussianSampler,
    MultimodalSampler,
    Sampler,
    SamplingResult,
    sample_gaussian,
    sample_multimodal,
)

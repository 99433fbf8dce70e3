"""energy-based models of the hot path (mirror of tsu/models/__init__.py)"""
from .ising import (  # noqa: F401
    IsingChain,
    IsingConfig,
    IsingGrid,
    IsingModel,
    IsingModel2D,
    demonstrate_phase_transition,
)

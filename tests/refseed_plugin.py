"""pytest plugin for tests/test_reference_suite.py: the reference's acceptance tests are statistical and unseeded (a
KS test at p > 0.05 fails one run in twenty by construction).  Seeding NumPy's global stream before every test - what
a caller of the reference does with np.random.seed - makes the run reproducible: the engine draws its Philox seeds
from that stream."""
import numpy as np
import pytest


@pytest.fixture(autouse=True)
def _seed_numpy_global_stream():
    np.random.seed(20240607)
    yield

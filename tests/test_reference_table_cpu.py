"""tests/reference_benchmark_table.py stays runnable: the reference's drivers on the reference's own code, quick mode."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_benchmark_table_script_runs_on_the_reference_backend():
    from oracle import make_ref

    if make_ref.ref_root() is None:
        pytest.skip("no reference tree (neither /root/reference nor oracle/_ref)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "reference_benchmark_table.py"), "--backend", "reference",
                        "--quick"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l and not l.startswith("#")]
    names = [l.split("|")[0].strip() for l in lines]
    for want in ("gaussian_1d", "boltzmann", "multimodal", "maxcut", "partition", "sampling", "optimization"):
        assert want in names, (want, names)
    assert "sampler class = tsu.gibbs.GibbsSampler" in r.stdout   # the reference's class, not the engine's

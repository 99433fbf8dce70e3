"""
Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, importable only in the build container) on seeded inputs.

    python -m oracle.make_golden

Each fixture stores the inputs and the reference's outputs; tests compare both the CPU oracle
(-m "not gpu") and the CUDA kernels (-m gpu) against them.  The reference draws from the global
numpy stream; `injected_numpy_random` (oracle/ref_loader.py) feeds it the recorded draws.
"""

import os

import numpy as np

from . import ising2d_oracle as O
from .ref_loader import injected_numpy_random, load_reference

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def lattice_case(name, rows, cols, periodic, T, J, h, bias_mode, n_sweeps, seed):
    gibbs, core, ising = load_reference()
    rng = np.random.default_rng(seed)
    bits0 = rng.integers(0, 2, (rows, cols))
    U = rng.integers(0, 2**32, (n_sweeps, rows, cols), dtype=np.uint64).astype(np.uint32)
    g = ising.IsingGrid((rows, cols), J=J, config=ising.IsingConfig(temperature=T, external_field=h), periodic=periodic)
    Jb = g._get_bit_coupling()
    hb = g._get_bit_bias() if bias_mode == "reference" else 2 * g.h - 2 * g.J.sum(1)
    order = O.checkerboard_order(rows, cols)
    seq = []
    for t in range(n_sweeps):
        seq.extend((U[t].ravel()[order].astype(np.float64) / 2**32).tolist())
    smp = gibbs.GibbsSampler(gibbs.GibbsConfig(temperature=T, update_order="random"))
    with injected_numpy_random(uniforms=seq, order=order):
        out = smp.gibbs_sweep(bits0.ravel().copy(), Jb, hb, n_sweeps=n_sweeps)  # gibbs.py:128-162
    out = out.reshape(rows, cols)
    spins = 2 * out.ravel() - 1
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, f"lattice_{name}.npz"),
        rows=rows, cols=cols, periodic=periodic, T=T, J=J, h=h, bias_mode=bias_mode,
        bits0=bits0.astype(np.uint8), uniforms=U, bits_out=out.astype(np.uint8),
        energy=float(g.energy(spins)),  # ising.py:98-117
        magnetization=float(g.magnetization(spins[None, :])),  # ising.py:183-193
    )
    return out


LATTICE_CASES = [
    # name, rows, cols, periodic, T, J, h, bias_mode, n_sweeps, seed
    ("p8x8_tc", 8, 8, True, 2.269, 1.0, 0.0, "physical", 6, 1),
    ("p16x64_cold", 16, 64, True, 0.1, 1.0, 0.0, "physical", 3, 2),
    ("o6x10_field", 6, 10, False, 2.5, 1.0, 0.3, "physical", 5, 3),
    ("o7x5_refbias", 7, 5, False, 1.0, 0.7, -0.2, "reference", 5, 4),
    ("p2x6_afm", 2, 6, True, 3.0, -1.0, 0.1, "physical", 5, 5),
    ("o1x9_chain", 1, 9, False, 1.5, 1.0, 0.0, "physical", 5, 6),
    ("o9x1_col", 9, 1, False, 1.5, 1.0, 0.2, "physical", 5, 7),
    ("o12x70_wide", 12, 70, False, 2.0, 1.0, 0.0, "physical", 3, 8),
    ("p4x130_ragged", 4, 130, True, 2.269, 1.0, 0.05, "physical", 3, 9),
]


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for case in LATTICE_CASES:
        lattice_case(*case)
        print("lattice", case[0])
    try:
        from .make_golden_dense import main as dense_main

        dense_main()
    except ImportError:
        pass
    try:
        from .make_golden_langevin import main as lang_main

        lang_main()
    except ImportError:
        pass


if __name__ == "__main__":
    main()

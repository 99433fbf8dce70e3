"""CPU: the lattice oracle against the golden vectors produced by the unmodified reference,
and (when /root/reference is present) against the reference itself on fresh inputs."""
import glob
import os

import numpy as np
import pytest

from oracle import ising2d_oracle as O
from oracle.ref_loader import injected_numpy_random, load_reference, reference_available


def golden_cases(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "lattice_*.npz")))


def test_golden_fixtures_present(golden_dir):
    assert len(golden_cases(golden_dir)) >= 9


def test_oracle_reproduces_reference_goldens(golden_dir):
    for path in golden_cases(golden_dir):
        g = np.load(path)
        out = O.checkerboard_sweeps(
            g["bits0"].astype(np.int64), g["uniforms"], float(g["J"]), float(g["h"]), float(g["T"]),
            bool(g["periodic"]), str(g["bias_mode"]),
        )
        assert (out == g["bits_out"]).all(), path
        assert O.energy(out, float(g["J"]), float(g["h"]), bool(g["periodic"])) == pytest.approx(float(g["energy"]), abs=1e-9)
        assert O.magnetization(out) == pytest.approx(float(g["magnetization"]), abs=1e-12)


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(3)
    for shape in [(1, 1), (3, 5), (7, 50), (4, 64), (5, 65), (2, 300)]:
        b = rng.integers(0, 2, shape)
        p = O.pack_spins(b)
        assert p.shape == (2, shape[0], O.words_per_row(shape[1]))
        assert (O.unpack_spins(p, *shape) == b).all()


def test_threshold_equivalence():
    # (k / 2^32 < p) <=> (k < ceil(p 2^32)) for integer k
    rng = np.random.default_rng(0)
    for p in list(rng.random(50)) + [0.0, 1.0, 0.5, 2.0**-32, 1 - 2.0**-33]:
        t = O.threshold_u32(p)
        for k in {0, 1, max(t - 1, 0), min(t, 2**32 - 1), min(t + 1, 2**32 - 1), 2**32 - 1}:
            assert ((k / 4294967296.0) < p) == (k < t)


def test_literal_port_equals_vectorised_oracle():
    rng = np.random.default_rng(11)
    R, C, T, J, h = 6, 8, 2.0, 1.0, 0.1
    bits0 = rng.integers(0, 2, (R, C))
    U = rng.integers(0, 2**32, (3, R, C), dtype=np.uint64).astype(np.uint32)
    Jb, hb = O.dense_bit_model(R, C, J, h, True)
    order = O.checkerboard_order(R, C)
    s = bits0.ravel().copy()
    for t in range(3):
        s = O.gibbs_sweep_port(s, Jb, hb, T, order, U[t].ravel()[order].astype(np.float64) / 2**32)
    assert (s.reshape(R, C) == O.checkerboard_sweeps(bits0, U, J, h, T, True)).all()


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
def test_oracle_equals_unmodified_reference_fresh_inputs():
    gibbs, core, ising = load_reference()
    rng = np.random.default_rng(2024)
    for (R, C, per, T, h, J) in [(8, 12, True, 2.269, 0.0, 1.0), (5, 9, False, 1.7, 0.25, 1.0), (10, 4, True, 0.4, 0.0, -1.0)]:
        bits0 = rng.integers(0, 2, (R, C))
        U = rng.integers(0, 2**32, (4, R, C), dtype=np.uint64).astype(np.uint32)
        g = ising.IsingGrid((R, C), J=J, config=ising.IsingConfig(temperature=T, external_field=h), periodic=per)
        Jb = g._get_bit_coupling()
        hb = 2 * g.h - 2 * g.J.sum(1)
        order = O.checkerboard_order(R, C)
        seq = []
        for t in range(4):
            seq.extend((U[t].ravel()[order].astype(np.float64) / 2**32).tolist())
        smp = gibbs.GibbsSampler(gibbs.GibbsConfig(temperature=T, update_order="random"))
        with injected_numpy_random(uniforms=seq, order=order):
            ref = smp.gibbs_sweep(bits0.ravel().copy(), Jb, hb, n_sweeps=4).reshape(R, C)
        assert (ref == O.checkerboard_sweeps(bits0, U, J, h, T, per)).all()
        # the reference's own (sign-flipped) bias is reproduced by bias_mode="reference"
        with injected_numpy_random(uniforms=seq, order=order):
            ref2 = smp.gibbs_sweep(bits0.ravel().copy(), Jb, g._get_bit_bias(), n_sweeps=4).reshape(R, C)
        assert (ref2 == O.checkerboard_sweeps(bits0, U, J, h, T, per, "reference")).all()


def test_lut_matches_oracle_probabilities():
    from tsu_emulator_b200.lattice import build_lut

    for (J, h, T, mode) in [(1.0, 0.0, 2.269, "physical"), (0.7, -0.2, 1.0, "reference"), (1.0, 0.3, 0.1, "physical"), (-1.0, 0.1, 3.0, "physical")]:
        lut = build_lut(J, h, T, mode)
        for d in range(5):
            for u in range(d + 1):
                p = O.acceptance_probability(J, h, T, np.array([u]), np.array([d]), mode)[0]
                t = O.threshold_u32(p)
                if t >= 2**32:
                    assert (int(lut[25]) >> (d * 5 + u)) & 1
                else:
                    assert not (int(lut[25]) >> (d * 5 + u)) & 1
                    assert int(lut[d * 5 + u]) == t

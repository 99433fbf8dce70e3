"""Philox4x32-10: oracle and library host code against the Random123 known-answer vectors."""
import ctypes

import numpy as np
import pytest

from oracle.philox_ref import RANDOM123_KAT, philox4x32_10


def test_oracle_matches_random123_kat():
    for ctr, key, expect in RANDOM123_KAT:
        out = philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in out) == expect


def test_library_host_philox_matches_kat():
    from tsu_emulator_b200 import _lib

    lib = _lib.load()
    for ctr, key, expect in RANDOM123_KAT:
        c = (ctypes.c_uint32 * 4)(*ctr)
        k = (ctypes.c_uint32 * 2)(*key)
        o = (ctypes.c_uint32 * 4)()
        lib.tsu_philox4x32_10_host(c, k, o)
        assert tuple(o) == expect


def test_oracle_vectorised_equals_scalar():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 2**32, (4, 50), dtype=np.uint64)
    k = rng.integers(0, 2**32, (2, 50), dtype=np.uint64)
    vec = philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1])
    for i in range(50):
        one = philox4x32_10(*[int(c[j, i]) for j in range(4)], int(k[0, i]), int(k[1, i]))
        assert [int(v[i]) for v in vec] == [int(x) for x in one]


@pytest.mark.gpu
def test_device_fill_matches_oracle():
    import torch

    from tsu_emulator_b200 import _lib

    n = 1003
    out = torch.zeros(n, dtype=torch.int32, device="cuda")
    seed = 0x1234_5678_9ABC_DEF0
    _lib.call("tsu_philox_fill_u32", _lib.ptr(out), n, seed, 7, _lib.current_stream())
    got = out.cpu().numpy().view(np.uint32)
    blk = np.arange((n + 3) // 4, dtype=np.uint64)
    o = philox4x32_10(blk, 0, 7, 0x46494C4C, seed & 0xFFFFFFFF, seed >> 32)
    want = np.stack(o, axis=1).ravel()[:n]
    assert (got == want).all()

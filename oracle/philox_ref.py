"""
ORACLE (test infrastructure, not product code): NumPy Philox4x32-10.

The reference (tsu-emulator) draws from NumPy's global MT19937 stream
(tsu/gibbs.py:126,157,201; tsu/core.py:78,143).  BASELINE.json's north_star replaces
that with a counter-based Philox layer, so there is no reference code to follow for
the generator itself.  This file restates the *published* Philox4x32-10 algorithm
(Salmon et al., "Parallel Random Numbers: As Easy as 1, 2, 3", SC'11; Random123
library, philox.h) and is pinned by the Random123 known-answer vectors in
tests/test_philox.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.

    All arguments are broadcastable integer arrays (values < 2**32).
    Returns four uint32 arrays (the 4 output words).
    """
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(
        *[np.asarray(a).astype(np.uint64) & _MASK32 for a in (c0, c1, c2, c3, k0, k1)]
    )
    c0 = c0.copy()
    c1 = c1.copy()
    c2 = c2.copy()
    c3 = c3.copy()
    k0 = k0.copy()
    k1 = k1.copy()
    for rnd in range(10):
        if rnd > 0:
            k0 = (k0 + np.uint64(PHILOX_W0)) & _MASK32
            k1 = (k1 + np.uint64(PHILOX_W1)) & _MASK32
        p0 = PHILOX_M0 * c0  # < 2**64, exact in uint64
        p1 = PHILOX_M1 * c2
        hi0 = p0 >> np.uint64(32)
        lo0 = p0 & _MASK32
        hi1 = p1 >> np.uint64(32)
        lo1 = p1 & _MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0), lo1, (hi0 ^ c3 ^ k1), lo0
    return (
        c0.astype(np.uint32),
        c1.astype(np.uint32),
        c2.astype(np.uint32),
        c3.astype(np.uint32),
    )


# Random123 kat_vectors, "philox4x32 10" rows: (counter[4], key[2], expected[4])
RANDOM123_KAT = [
    (
        (0x00000000, 0x00000000, 0x00000000, 0x00000000),
        (0x00000000, 0x00000000),
        (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8),
    ),
    (
        (0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF),
        (0xFFFFFFFF, 0xFFFFFFFF),
        (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD),
    ),
    (
        (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344),
        (0xA4093822, 0x299F31D0),
        (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1),
    ),
]

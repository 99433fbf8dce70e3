"""
ORACLE (test infrastructure, not product code): CPU restatement of the reference's
heat-bath Gibbs update specialised to the 2-D nearest-neighbour lattice.

What it follows in the reference (/root/reference, read-only):
  * acceptance rule          tsu/gibbs.py:102-126  (p = sigmoid(h_i/T); new bit = 1 iff rand() < p)
  * sigmoid with +-20 clamp  tsu/gibbs.py:61-77
  * local field              tsu/gibbs.py:79-100   (row . state + bias, bit variables)
  * sweep semantics          tsu/gibbs.py:128-162  (in-place, every update sees earlier ones)
  * lattice wiring           tsu/models/ising.py:343-361 (row-major idx, right/down bonds, optional wrap)
  * spin<->bit maps          tsu/models/ising.py:119-138 (J_bit = 4 J)
  * bit bias                 tsu/models/ising.py:140-148 (reference: -2h + 2 rowsum(J); see `bias_mode`)
  * energy / magnetisation   tsu/models/ising.py:98-117,183-193

Parity pinning: tests/test_oracle_lattice.py drives the UNMODIFIED reference
`GibbsSampler.gibbs_sweep` (update_order="random", numpy.random.permutation patched to
black-then-white order, numpy.random.rand patched to the injected uniforms) and checks
bit equality with `checkerboard_sweeps` below; the same script writes tests/golden/*.npz.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""

import math
import numpy as np

from .philox_ref import philox4x32_10

N_PLANES = 8  # top bits of every uniform come from bit-planes (format constant)
KIND_PLANE0 = 0
KIND_PLANE1 = 1
KIND_INIT = 2
KIND_LOW0 = 8  # 8..15: low-24-bit calls, one per group of 4 lanes


# --------------------------------------------------------------------------- layout
def words_per_row(cols: int) -> int:
    """uint32 words per (row, colour), padded to a multiple of 4 (16-byte vectors)."""
    ck = (cols + 1) // 2
    w = (ck + 31) // 32
    return (w + 3) // 4 * 4


def colour_count(cols: int, row: int, colour: int) -> int:
    """number of sites of `colour` in `row` (column j = 2k + ((row+colour)&1))."""
    p = (row + colour) & 1
    return (cols - p + 1) // 2


def pack_spins(bits: np.ndarray, row0: int = 0) -> np.ndarray:
    """bits[R, C] in {0,1}  ->  packed[2, R, wpr] uint32 (colour-major, frozen layout); row0 = global
    index of local row 0 (row slabs)."""
    bits = np.asarray(bits)
    R, C = bits.shape
    wpr = words_per_row(C)
    out = np.zeros((2, R, wpr), dtype=np.uint32)
    for colour in range(2):
        for i in range(R):
            p = (row0 + i + colour) & 1
            row = bits[i, p::2].astype(np.uint64)
            k = np.arange(row.size)
            np.bitwise_or.at(out[colour, i], k >> 5, (row << (k & 31).astype(np.uint64)).astype(np.uint32))
    return out


def unpack_spins(packed: np.ndarray, rows: int, cols: int, row0: int = 0) -> np.ndarray:
    """inverse of pack_spins -> bits[R, C] int64 in {0,1}."""
    packed = np.asarray(packed, dtype=np.uint32)
    bits = np.zeros((rows, cols), dtype=np.int64)
    for colour in range(2):
        for i in range(rows):
            p = (row0 + i + colour) & 1
            n = colour_count(cols, row0 + i, colour)
            k = np.arange(n)
            bits[i, p::2] = (packed[colour, i, k >> 5] >> (k & 31).astype(np.uint32)) & 1
    return bits


# --------------------------------------------------------------------------- uniforms
def _site_words(rows, cols, colour, row0=0):
    """for every site of `colour`: (i_global, k, w, b) index arrays shaped [rows, cols] masked."""
    i = np.arange(rows)[:, None] + row0
    j = np.arange(cols)[None, :]
    mask = ((i + j) & 1) == colour
    k = j >> 1
    return i, j, mask, k >> 5, k & 31


def lattice_counter(w, colour, kind, row_global):
    """Philox counter words 0 and 1 of a lattice call (words 2, 3 are sweep and replica): everything that varies
    along a thread's walk down its rows (row, word within its 4-word group, kind) sits in word 1.  Rows < 2^24."""
    w = np.asarray(w).astype(np.uint64)
    c0 = (w >> np.uint64(2)) | np.uint64(colour << 24)
    c1 = (
        np.asarray(row_global).astype(np.uint64)
        | ((w & np.uint64(3)) << np.uint64(24))
        | (np.asarray(kind).astype(np.uint64) << np.uint64(26))
    )
    return c0, c1


def lattice_uniforms_u32(seed: int, replica: int, sweep: int, colour: int, rows: int, cols: int, row0: int = 0):
    """32-bit uniform of every site of `colour` for sweep index `sweep`.

    Frozen definition (shared by the CUDA kernel):
      word w of (row i, colour c) owns compressed sites k in [32w, 32w+32), lane b = k & 31
      counter = ((w>>2) | c<<24,  i_global | (w&3)<<24 | kind<<26,  sweep,  replica),  key = (seed_lo, seed_hi)
      bit (31-p) of u, p in 0..7  = bit b of output word (p & 3) of call kind (p >> 2)
      low 24 bits of u            = output word (b & 3) of call kind 8 + (b >> 2), shifted right by 8
    Returns u32[rows, cols] (uint32; entries of the other colour are 0) and the colour mask.
    """
    i, j, mask, w, b = _site_words(rows, cols, colour, row0)
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) & 0xFFFFFFFF
    i_b, w_b, b_b = np.broadcast_arrays(i, w, b)
    u = np.zeros((rows, cols), dtype=np.uint64)
    for call in range(2):
        c0, c1 = lattice_counter(w_b, colour, KIND_PLANE0 + call, i_b)
        out = philox4x32_10(c0, c1, sweep & 0xFFFFFFFF, replica, k0, k1)
        for q in range(4):
            p = call * 4 + q
            bit = (out[q].astype(np.uint64) >> b_b.astype(np.uint64)) & np.uint64(1)
            u |= bit << np.uint64(31 - p)
    c0, c1 = lattice_counter(w_b, colour, np.uint64(KIND_LOW0) + (b_b.astype(np.uint64) >> np.uint64(2)), i_b)
    out = philox4x32_10(c0, c1, sweep & 0xFFFFFFFF, replica, k0, k1)
    sel = b_b & 3
    low = np.choose(sel, [o.astype(np.uint64) for o in out]) >> np.uint64(8)
    u |= low
    u = np.where(mask, u, 0).astype(np.uint32)
    return u, mask


def init_bits(seed: int, replica: int, rows: int, cols: int, row0: int = 0) -> np.ndarray:
    """iid Bernoulli(1/2) initial lattice: site bit = bit b of output word 0 of call kind=2, sweep 0."""
    bits = np.zeros((rows, cols), dtype=np.int64)
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) & 0xFFFFFFFF
    for colour in range(2):
        i, j, mask, w, b = _site_words(rows, cols, colour, row0)
        i_b, w_b, b_b = np.broadcast_arrays(i, w, b)
        c0, c1 = lattice_counter(w_b, colour, KIND_INIT, i_b)
        out = philox4x32_10(c0, c1, 0, replica, k0, k1)
        bit = (out[0].astype(np.uint64) >> b_b.astype(np.uint64)) & np.uint64(1)
        bits = np.where(mask, bit.astype(np.int64), bits)
    return bits


# --------------------------------------------------------------------------- update rule
def sigmoid_ref(x: float) -> float:
    """tsu/gibbs.py:61-77 verbatim semantics (strict > 20 / < -20 clamps, float64)."""
    if x > 20:
        return 1.0
    elif x < -20:
        return 0.0
    return 1.0 / (1.0 + np.exp(-x))


def neighbour_up_and_degree(bits: np.ndarray, periodic: bool):
    """number of up (bit=1) neighbours and number of neighbours of every site.

    Wiring of tsu/models/ising.py:343-361: right/down bonds, wrap iff periodic.  set_coupling
    ASSIGNS (ising.py:85-86), so a wrap bond that coincides with an existing bond (dimension 2)
    or with the site itself (dimension 1) is not double counted; callers restrict periodic
    lattices to even dimensions >= 2 and a periodic dimension of size 2 behaves as open.
    """
    R, C = bits.shape
    up = np.zeros((R, C), dtype=np.int64)
    deg = np.zeros((R, C), dtype=np.int64)
    wrap_r = periodic and R > 2
    wrap_c = periodic and C > 2
    # vertical
    up[1:, :] += bits[:-1, :]
    deg[1:, :] += 1
    up[:-1, :] += bits[1:, :]
    deg[:-1, :] += 1
    if wrap_r:
        up[0, :] += bits[-1, :]
        deg[0, :] += 1
        up[-1, :] += bits[0, :]
        deg[-1, :] += 1
    # horizontal
    up[:, 1:] += bits[:, :-1]
    deg[:, 1:] += 1
    up[:, :-1] += bits[:, 1:]
    deg[:, :-1] += 1
    if wrap_c:
        up[:, 0] += bits[:, -1]
        deg[:, 0] += 1
        up[:, -1] += bits[:, 0]
        deg[:, -1] += 1
    return up, deg


def bit_bias(J: float, h: float, deg, bias_mode: str = "physical"):
    """bias of the bit model for a site with `deg` neighbours.

    "physical":  2h - 2*rowsum(J)   (correct transformation of the spin Hamiltonian)
    "reference": -2h + 2*rowsum(J)  (what tsu/models/ising.py:140-148 returns)
    """
    rowsum = J * deg
    if bias_mode == "physical":
        return 2 * h - 2 * rowsum
    elif bias_mode == "reference":
        return -2 * h + 2 * rowsum
    raise ValueError(bias_mode)


def acceptance_probability(J, h, T, up, deg, bias_mode="physical") -> np.ndarray:
    """p = sigmoid((4J*up + bias)/T) per site, float64, via sigmoid_ref (gibbs.py:96-99,124-125)."""
    up = np.asarray(up)
    deg = np.broadcast_to(np.asarray(deg), up.shape)
    p = np.empty(up.shape, dtype=np.float64)
    cache = {}
    for idx in np.ndindex(up.shape):
        key = (int(up[idx]), int(deg[idx]))
        if key not in cache:
            field = float(4 * J * key[0]) + float(bit_bias(J, h, key[1], bias_mode))
            cache[key] = sigmoid_ref(field / T)
        p[idx] = cache[key]
    return p


def half_sweep(bits, colour, u32, J, h, T, periodic, bias_mode="physical", row0=0):
    """update every site of `colour` simultaneously (they do not interact) - in place."""
    R, C = bits.shape
    i = np.arange(R)[:, None] + row0
    j = np.arange(C)[None, :]
    mask = ((i + j) & 1) == colour
    up, deg = neighbour_up_and_degree(bits, periodic)
    p = acceptance_probability(J, h, T, up, deg, bias_mode)
    u = u32.astype(np.float64) / 4294967296.0  # exact
    new = (u < p).astype(bits.dtype)
    bits[mask] = new[mask]
    return bits


def half_sweep_slab(local, top, bot, colour, u32, J, h, T, wrap_cols, row0, bias_mode="physical"):
    """row-slab version of half_sweep: `local` are rows row0 .. row0+R-1 of a larger lattice, `top` / `bot` the
    full rows above / below the slab (None = open edge).  Updates the sites of `colour` of the local rows in place."""
    R, C = local.shape
    ext = np.vstack([np.zeros((1, C), dtype=local.dtype) if top is None else np.asarray(top).reshape(1, C), local,
                     np.zeros((1, C), dtype=local.dtype) if bot is None else np.asarray(bot).reshape(1, C)])
    up = np.zeros((R, C), dtype=np.int64)
    deg = np.zeros((R, C), dtype=np.int64)
    up += ext[0:R]
    up += ext[2:R + 2]
    deg += 2
    if top is None:
        deg[0] -= 1
    if bot is None:
        deg[-1] -= 1
    up[:, 1:] += local[:, :-1]
    deg[:, 1:] += 1
    up[:, :-1] += local[:, 1:]
    deg[:, :-1] += 1
    if wrap_cols:
        up[:, 0] += local[:, -1]
        up[:, -1] += local[:, 0]
        deg[:, 0] += 1
        deg[:, -1] += 1
    p = acceptance_probability(J, h, T, up, deg, bias_mode)
    new = ((u32.astype(np.float64) / 4294967296.0) < p).astype(local.dtype)
    i = np.arange(R)[:, None] + row0
    j = np.arange(C)[None, :]
    mask = ((i + j) & 1) == colour
    local[mask] = new[mask]
    return local


def checkerboard_sweeps(bits0, u32_per_sweep, J, h, T, periodic, bias_mode="physical"):
    """n sweeps (black half-sweep then white half-sweep) with injected per-site uniforms.

    u32_per_sweep: uint32[n_sweeps, R, C]; entry [t, i, j] is consumed by site (i, j) in sweep t.
    """
    bits = np.array(bits0, dtype=np.int64, copy=True)
    for t in range(len(u32_per_sweep)):
        for colour in (0, 1):
            half_sweep(bits, colour, u32_per_sweep[t], J, h, T, periodic, bias_mode)
    return bits


def philox_uniform_field(seed, replica, sweep, rows, cols, row0=0):
    """u32[R, C] holding, for both colours, the Philox uniform of every site for `sweep`."""
    u0, m0 = lattice_uniforms_u32(seed, replica, sweep, 0, rows, cols, row0)
    u1, _ = lattice_uniforms_u32(seed, replica, sweep, 1, rows, cols, row0)
    return np.where(m0, u0, u1).astype(np.uint32)


def checkerboard_sweeps_philox(bits0, seed, replica, sweep0, n_sweeps, J, h, T, periodic, bias_mode="physical"):
    """what the CUDA kernel must reproduce bit-for-bit in its native (Philox) mode."""
    R, C = np.asarray(bits0).shape
    bits = np.array(bits0, dtype=np.int64, copy=True)
    for t in range(n_sweeps):
        u = philox_uniform_field(seed, replica, sweep0 + t, R, C)
        for colour in (0, 1):
            half_sweep(bits, colour, u, J, h, T, periodic, bias_mode)
    return bits


# --------------------------------------------------------------------------- literal port
def dense_bit_model(rows, cols, J, h, periodic, bias_mode="physical"):
    """dense (J_bit, h_bit) exactly as IsingGrid + _get_bit_coupling/_get_bit_bias would build."""
    n = rows * cols
    Jm = np.zeros((n, n))
    for i in range(rows):
        for j in range(cols):
            idx = i * cols + j
            if j < cols - 1:
                Jm[idx, idx + 1] = J
                Jm[idx + 1, idx] = J
            elif periodic:
                r = i * cols
                Jm[idx, r] = J
                Jm[r, idx] = J
            if i < rows - 1:
                d = idx + cols
                Jm[idx, d] = J
                Jm[d, idx] = J
            elif periodic:
                Jm[idx, j] = J
                Jm[j, idx] = J
    hv = np.ones(n) * h
    if bias_mode == "physical":
        hb = 2 * hv - 2 * np.sum(Jm, axis=1)
    else:
        hb = -2 * hv + 2 * np.sum(Jm, axis=1)
    return 4 * Jm, hb


def gibbs_sweep_port(state, coupling, bias, T, order, uniforms):
    """literal restatement of tsu/gibbs.py:128-162 for one sweep: visit `order`, draw from `uniforms`.

    This is the per-spin Python loop the reference executes (np.dot + sigmoid + compare);
    it is also what bench.py times as the CPU baseline (kind "port").
    """
    state = state.copy()
    it = iter(uniforms)
    for i in order:
        hfield = np.dot(coupling[i, :], state)
        if bias is not None:
            hfield += bias[i]
        prob = sigmoid_ref(float(hfield) / T)
        state[i] = 1 if next(it) < prob else 0
    return state


def checkerboard_order(rows, cols):
    i, j = np.divmod(np.arange(rows * cols), cols)
    return np.concatenate([np.flatnonzero((i + j) % 2 == 0), np.flatnonzero((i + j) % 2 == 1)])


# --------------------------------------------------------------------------- observables
def magnetization(bits) -> float:
    """signed magnetisation per spin of one configuration (ising.py:183-193 on one sample)."""
    s = 2 * np.asarray(bits, dtype=np.int64) - 1
    return float(s.sum()) / s.size


def energy(bits, J, h, periodic) -> float:
    """E = -sum_<ij> J s_i s_j - h sum_i s_i (ising.py:98-117 with the lattice wiring)."""
    s = 2 * np.asarray(bits, dtype=np.int64) - 1
    R, C = s.shape
    bonds = (s[:, :-1] * s[:, 1:]).sum() + (s[:-1, :] * s[1:, :]).sum()
    if periodic and C > 2:
        bonds += (s[:, -1] * s[:, 0]).sum()
    if periodic and R > 2:
        bonds += (s[-1, :] * s[0, :]).sum()
    return float(-J * bonds - h * s.sum())


def threshold_u32(p: float) -> int:
    """integer threshold t with  (k / 2**32 < p)  <=>  (k < t)  for every integer 0 <= k < 2**32."""
    return int(math.ceil(p * 4294967296.0))

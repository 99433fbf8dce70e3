"""
ORACLE (test infrastructure, not product code): CPU restatement of the reference's dense-coupling
Gibbs sampler with every random draw injected.

Follows /root/reference/tsu/gibbs.py:
  sigmoid (61-77), local field incl. self term (79-100), sample_conditional (102-126),
  gibbs_sweep (128-162), sample_boltzmann (164-213), compute_energy (215-236),
  parallel_tempering (238-338), simulated_annealing (340-393).
Pinned by tests/golden/dense_*.npz, which oracle/make_golden_dense.py produced by running the
unmodified reference with numpy.random patched to the recorded draws.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""

import numpy as np

from .philox_ref import philox4x32_10

STREAM_DENSE = 0x44454E53
STREAM_DENSE_TC = 0x44454E54
STREAM_DENSE_INIT = 0x44494E49
STREAM_PT_SWAP = 0x50545357


def sigmoid_ref(x):
    if x > 20:
        return 1.0
    elif x < -20:
        return 0.0
    return 1.0 / (1.0 + np.exp(-x))


def gibbs_sweeps(state, coupling, bias, T, n_sweeps, uniforms, orders=None):
    """gibbs.py:128-162.  uniforms[s][k] is the draw of the k-th visit of sweep s; orders[s] the visiting order."""
    state = np.array(state, copy=True)
    n = len(state)
    for s in range(n_sweeps):
        idx = range(n) if orders is None else orders[s]
        for k, i in enumerate(idx):
            h = np.dot(coupling[i, :], state)
            if bias is not None:
                h += bias[i]
            prob = sigmoid_ref(float(h) / T)
            state[i] = 1 if uniforms[s][k] < prob else 0
    return state


def compute_energy(state, coupling, bias=None):
    e = -0.5 * state.dot(coupling).dot(state)
    if bias is not None:
        e -= bias.dot(state)
    return float(e)


def sample_boltzmann(coupling, bias, T, n_burnin, n_samples, n_sweeps, initial_state, uniforms, orders=None):
    """gibbs.py:164-213; uniforms has n_burnin + n_samples*n_sweeps rows"""
    state = gibbs_sweeps(initial_state, coupling, bias, T, n_burnin, uniforms[:n_burnin], None if orders is None else orders[:n_burnin])
    out = np.zeros((n_samples, len(state)), dtype=int)
    pos = n_burnin
    for k in range(n_samples):
        state = gibbs_sweeps(state, coupling, bias, T, n_sweeps, uniforms[pos:pos + n_sweeps],
                             None if orders is None else orders[pos:pos + n_sweeps])
        pos += n_sweeps
        out[k] = state
    return out


def parallel_tempering(coupling, bias, temperatures, n_burnin, n_sweeps, n_samples, swap_interval, init_states,
                       burn_uniforms, sweep_uniforms, swap_uniforms):
    """gibbs.py:238-338 with separated draw streams.

    burn_uniforms[slot][s][k], sweep_uniforms[it][slot][s][k]: draws of the sweeps of temperature slot `slot`;
    swap_uniforms[it][pair]: draw offered to pair (i, i+1) at iteration `it` (used only if delta < 0).
    """
    R = len(temperatures)
    states = [np.array(s, copy=True) for s in init_states]
    for i in range(R):
        states[i] = gibbs_sweeps(states[i], coupling, bias, temperatures[i], n_burnin, burn_uniforms[i])
    samples, attempts, accepts = [], 0, 0
    hist = [[] for _ in range(R)]
    it = 0
    while len(samples) < n_samples:
        for i in range(R):
            states[i] = gibbs_sweeps(states[i], coupling, bias, temperatures[i], n_sweeps, sweep_uniforms[it][i])
            hist[i].append(compute_energy(states[i], coupling, bias))
        it += 1
        if it % swap_interval == 0:
            for i in range(R - 1):
                Ei = compute_energy(states[i], coupling, bias)
                Ej = compute_energy(states[i + 1], coupling, bias)
                delta = (1.0 / temperatures[i] - 1.0 / temperatures[i + 1]) * (Ej - Ei)
                attempts += 1
                if delta >= 0 or swap_uniforms[it - 1][i] < np.exp(delta):
                    states[i], states[i + 1] = states[i + 1], states[i]
                    accepts += 1
        samples.append(states[0].copy())
    return np.array(samples[:n_samples]), {"swap_attempts": attempts, "swap_accepts": accepts, "energies": hist,
                                           "final_states": states}


def simulated_annealing(coupling, bias, T_initial, T_final, n_steps, schedule, init_state, uniforms):
    """gibbs.py:340-393"""
    state = np.array(init_state, copy=True)
    best_state, best_energy = state.copy(), compute_energy(state, coupling, bias)
    for step in range(n_steps):
        if schedule == "exponential":
            T = T_initial * (T_final / T_initial) ** (step / n_steps)
        else:
            T = T_initial + (T_final - T_initial) * step / n_steps
        state = gibbs_sweeps(state, coupling, bias, T, 1, uniforms[step:step + 1])
        e = compute_energy(state, coupling, bias)
        if e < best_energy:
            best_energy, best_state = e, state.copy()
    return best_state, best_energy


# ---- emulation of the kernel's Philox streams (csrc/dense_gibbs.cu) -------------------------------
def philox_uniforms(seed, chain, sweep, sites):
    """float64 uniform of (site, chain, sweep): 53 bits from words x,y of Philox(counter=(site, chain, sweep, 'DENS'))"""
    sites = np.asarray(sites, dtype=np.uint64)
    o = philox4x32_10(sites, chain, sweep & 0xFFFFFFFF, STREAM_DENSE, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    m = ((o[0].astype(np.uint64) << np.uint64(32)) | o[1].astype(np.uint64)) >> np.uint64(11)
    return m.astype(np.float64) * (1.0 / 9007199254740992.0)


def philox_init_state(seed, chain, N):
    w = np.arange((N + 31) // 32, dtype=np.uint64)
    o = philox4x32_10(w, chain, 0, STREAM_DENSE_INIT, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    i = np.arange(N)
    return ((o[0][i >> 5] >> (i & 31).astype(np.uint32)) & 1).astype(np.int64)


def philox_uniforms_tc(seed, chain, sweep, N):
    """uniforms of the tensor-core dense path (csrc/dense_tc.cu): 24 bits of word (site & 3) of
    Philox(counter = (site >> 2, chain, sweep, 'DENT')), value = (word >> 8) / 2^24 (exact in float32)"""
    sites = np.arange(N, dtype=np.uint64)
    o = philox4x32_10(sites >> np.uint64(2), chain, sweep & 0xFFFFFFFF, STREAM_DENSE_TC, seed & 0xFFFFFFFF,
                      (seed >> 32) & 0xFFFFFFFF)
    w = np.choose((sites & np.uint64(3)).astype(np.int64), [x.astype(np.uint64) for x in o])
    return (w >> np.uint64(8)).astype(np.float64) / 16777216.0


def tc_sweep_disagreements(start, after, coupling, bias, T, uniforms):
    """One sequential sweep (gibbs.py:153-160) of ONE chain replayed against a kernel's result `after`.

    The replay follows the kernel's own trajectory: at every site the float64 field of the reference
    (gibbs.py:97-99) is evaluated on the state the kernel had at that moment, the reference's decision
    `u < sigmoid(h / T)` (gibbs.py:125-126) is compared with the kernel's bit, and the kernel's bit is kept.
    Returns [(site, |u - p|)] for the sites where the two decisions differ: a reduced-precision kernel may only
    disagree where the uniform lies within its field / threshold error of the acceptance probability.
    """
    cur = np.array(start, dtype=np.float64, copy=True)
    J = np.asarray(coupling, dtype=np.float64)
    out = []
    for i in range(len(cur)):
        h = float(np.dot(J[i, :], cur)) + (0.0 if bias is None else float(bias[i]))
        p = sigmoid_ref(h / T)
        want = 1 if uniforms[i] < p else 0
        if want != int(after[i]):
            out.append((i, abs(float(uniforms[i]) - p)))
        cur[i] = after[i]
    return out


def greedy_colour_order(coupling):
    """visiting order of the chromatic sparse sampler: greedy colouring (sites in index order take the smallest colour
    no coupled, already coloured site has; coupling in either direction counts), then colour 0's sites in ascending
    index, colour 1's, ...  A valid permutation for the reference's update_order="random" (gibbs.py:155-157)."""
    J = np.asarray(coupling)
    N = J.shape[0]
    adj = (J != 0) | (J.T != 0)
    np.fill_diagonal(adj, False)
    colour = -np.ones(N, dtype=np.int64)
    for i in range(N):
        used = set(colour[np.flatnonzero(adj[i])].tolist())
        c = 0
        while c in used:
            c += 1
        colour[i] = c
    return np.argsort(colour, kind="stable"), colour

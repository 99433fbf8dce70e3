// 2-D nearest-neighbour Ising lattice: bit-packed checkerboard heat-bath Gibbs update for sm_100a.
//
// Replaces the per-spin Python loop of the reference
//   GibbsSampler.gibbs_sweep / sample_conditional   tsu/gibbs.py:102-162
// for the lattice wired by IsingGrid                 tsu/models/ising.py:320-361
// (and README's IsingModel2D.gibbs_update, README.md:116-131).
//
// Multi-spin coding: one uint32 word = 32 spins of one colour of one row.  For a word the four
// neighbour words (north, south, centre, side-shifted) are reduced to a bit-sliced up-count
// (c2 c1 c0) with 6 LOP3 + 1 funnel shift.  The acceptance test  u < t[class]  (u: 32-bit uniform
// per spin, t: integer threshold = ceil(sigmoid(h/T) * 2^32) from the host LUT) is evaluated
// bit-sliced as well: the top 8 bits of all 32 uniforms are 8 random bit-planes = 2 Philox calls;
// a borrow chain gives "less than" and "equal so far" masks.  Only lanes whose top 8 bits tie with
// the threshold (probability 2^-8) need the low 24 bits, which come from a per-lane-group Philox
// call.  The result is identical to a full 32-bit compare per spin, so it is bit-exact with the
// CPU oracle (oracle/ising2d_oracle.py) and with the reference's `rand() < prob`.
//
// HBM traffic per half-sweep: read the opposite colour once (+2 halo rows per strip), write the
// updated colour once = 2 bits per spin update; the kernel is integer-issue bound, not DRAM bound
// (see DESIGN.md for the roofline arithmetic).

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "philox.cuh"
#include "ising2d_fast.cuh"
#include "jit.cuh"

namespace {

using tsu_fast::Geom;
using tsu_fast::colour_count;
using tsu_fast::words_per_row;
using tsu_fast::Planes;
using tsu_fast::opp_row;
using tsu_fast::Coords;
using tsu_fast::lattice_call;
using tsu_fast::SweepParams;

__device__ __forceinline__ uint32_t lane_mask_lt(int n) {  // lanes [0, n), n clamped to [0, 32]
  return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u));
}

// pointers to the two colour planes of one replica
struct Hood {
  uint32_t n, s, c, side;  // neighbour bit words (missing neighbours read 0)
  uint32_t valid;          // own lanes that exist
  uint32_t missW, missE;   // own lanes without a west / east neighbour (open columns)
  int has_n, has_s;
};

__device__ __forceinline__ Hood load_hood(const Planes& P, const Geom& g, int colour, int i, int w) {
  Hood h;
  const int p = (g.row0 + i + colour) & 1;  // column offset of the own colour in this row
  const int nk_own = colour_count(g.cols, p);
  const int nk_opp = colour_count(g.cols, 1 - p);
  h.valid = lane_mask_lt(nk_own - 32 * w);
  const uint32_t* rn = opp_row(P, g, i - 1);
  const uint32_t* rs = opp_row(P, g, i + 1);
  const uint32_t* rc = P.opp + (size_t)i * g.wpr;
  h.has_n = rn != nullptr;
  h.has_s = rs != nullptr;
  h.n = rn ? rn[w] : 0u;
  h.s = rs ? rs[w] : 0u;
  h.c = rc[w];
  h.missW = 0u;
  h.missE = 0u;
  const int last = nk_own - 1;  // last own lane index in the row
  if (p) {
    // own col = 2k+1: west = opp k (centre), east = opp k+1
    uint32_t nx = (w + 1 < g.wpr) ? rc[w + 1] : 0u;
    h.side = (h.c >> 1) | (nx << 31);
    if (last >= 0 && (last >> 5) == w && 2 * last + 2 >= g.cols) {  // east neighbour would be col >= cols
      if (g.wrap_cols)
        h.side |= (rc[0] & 1u) << (last & 31);
      else
        h.missE = 1u << (last & 31);
    }
  } else {
    // own col = 2k: east = opp k (centre), west = opp k-1
    uint32_t pv = (w > 0) ? rc[w - 1] : 0u;
    h.side = (h.c << 1) | (pv >> 31);
    if (w == 0) {
      if (g.wrap_cols) {
        int lo = nk_opp - 1;
        h.side |= (rc[lo >> 5] >> (lo & 31)) & 1u;
      } else {
        h.missW = 1u;
      }
    }
    if (last >= 0 && (last >> 5) == w && 2 * last + 1 >= g.cols) h.missE = 1u << (last & 31);  // odd cols
  }
  return h;
}

// ---- bit-sliced acceptance -------------------------------------------------------------
struct LutRegs {
  uint32_t K[8][5];  // K[k][u] = all-ones iff bit (31-k) of threshold(d, u) is set
  uint32_t always;   // bit u set: class (d, u) accepts with probability 1 (threshold 2^32)
};

__device__ __forceinline__ void load_lut_regs(LutRegs& L, const uint32_t* __restrict__ lut, int d) {
#pragma unroll
  for (int u = 0; u < 5; ++u) {
    uint32_t t = __ldg(lut + d * 5 + u);
#pragma unroll
    for (int k = 0; k < 8; ++k) L.K[k][u] = 0u - ((t >> (31 - k)) & 1u);
  }
  L.always = (__ldg(lut + 25) >> (d * 5)) & 31u;
}

__device__ __forceinline__ uint32_t bitsel(uint32_t m, uint32_t a, uint32_t b) {  // m ? a : b  (bitwise)
  return (m & a) | (~m & b);
}

__device__ __forceinline__ uint32_t lane_uniform(const uint32_t r[8], const Coords& q, uint32_t w, uint32_t row_g, int j) {
  uint32_t u = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) u |= ((r[k] >> j) & 1u) << (31 - k);
  tsu_u32x4 lo = lattice_call(q, w, row_g, TSU_KIND_LOW0 + (uint32_t)(j >> 2));
  int sel = j & 3;
  uint32_t v = sel == 0 ? lo.x : (sel == 1 ? lo.y : (sel == 2 ? lo.z : lo.w));
  return u | (v >> 8);
}

__device__ __forceinline__ uint32_t lane_accept(const uint32_t* __restrict__ lut, int d, int up, uint32_t u) {
  int cls = d * 5 + up;
  if ((__ldg(lut + 25) >> cls) & 1u) return 1u;
  return u < __ldg(lut + cls) ? 1u : 0u;
}

// New value of one word of the colour being updated.
//   a,b,c,s: neighbour words; d_row: number of neighbours of a regular lane (2 + has_n + has_s);
//   special: lanes with fewer neighbours than d_row (missW | missE), resolved one by one.
template <bool FAST>
__device__ __forceinline__ uint32_t update_word(uint32_t a, uint32_t b, uint32_t c, uint32_t s, const LutRegs& L,
                                                const uint32_t* __restrict__ lut, int d_row, uint32_t valid,
                                                uint32_t missW, uint32_t missE, const Coords& q, uint32_t w,
                                                uint32_t row_g) {
  // bit-sliced count of up neighbours: c2 c1 c0
  const uint32_t s1 = a ^ b ^ c;
  const uint32_t m1 = tsu_lop3_maj(a, b, c);
  const uint32_t c0 = s1 ^ s;
  const uint32_t k2 = s1 & s;
  const uint32_t c1 = m1 ^ k2;
  const uint32_t c2 = m1 & k2;

  uint32_t r[8];
  {
    tsu_u32x4 p0 = lattice_call(q, w, row_g, TSU_KIND_PLANE0);
    tsu_u32x4 p1 = lattice_call(q, w, row_g, TSU_KIND_PLANE1);
    r[0] = p0.x; r[1] = p0.y; r[2] = p0.z; r[3] = p0.w;
    r[4] = p1.x; r[5] = p1.y; r[6] = p1.z; r[7] = p1.w;
  }
  uint32_t lt = 0u, eq = 0xffffffffu;
#pragma unroll
  for (int k = 7; k >= 0; --k) {  // least significant of the 8 planes first
    const uint32_t tk = bitsel(c2, L.K[k][4], bitsel(c1, bitsel(c0, L.K[k][3], L.K[k][2]), bitsel(c0, L.K[k][1], L.K[k][0])));
    const uint32_t x = r[k] ^ tk;
    lt = (~r[k] & tk) | (~x & lt);
    eq &= ~x;
  }
  const uint32_t always = L.always;
  if (always) {  // classes with p == 1.0 (threshold 2^32 does not fit 32 bits)
    uint32_t am = 0u;
    if (always & 1u) am |= ~c2 & ~c1 & ~c0;
    if (always & 2u) am |= ~c2 & ~c1 & c0;
    if (always & 4u) am |= ~c2 & c1 & ~c0;
    if (always & 8u) am |= ~c2 & c1 & c0;
    if (always & 16u) am |= c2;
    lt |= am;
    eq &= ~am;
  }
  uint32_t special = 0u;
  if (!FAST) {
    special = (missW | missE) & valid;
    eq &= valid & ~special;
  }
  // lanes whose top 8 bits tie with the threshold: decide on the full 32-bit uniform
  while (eq) {
    const int j = __ffs(eq) - 1;
    eq &= eq - 1u;
    const int up = ((c0 >> j) & 1u) | (((c1 >> j) & 1u) << 1) | (((c2 >> j) & 1u) << 2);
    const uint32_t bit = lane_accept(lut, d_row, up, lane_uniform(r, q, w, row_g, j));
    lt = (lt & ~(1u << j)) | (bit << j);
  }
  if (!FAST) {
    while (special) {
      const int j = __ffs(special) - 1;
      special &= special - 1u;
      const int up = ((c0 >> j) & 1u) | (((c1 >> j) & 1u) << 1) | (((c2 >> j) & 1u) << 2);
      const int d = d_row - (int)((missW >> j) & 1u) - (int)((missE >> j) & 1u);
      const uint32_t bit = lane_accept(lut, d, up, lane_uniform(r, q, w, row_g, j));
      lt = (lt & ~(1u << j)) | (bit << j);
    }
    lt &= valid;
  }
  return lt;
}

#ifdef TSU_LATTICE_TRACE  // diagnosis build (tools/lattice_trace.py): when and where every CTA of the last launches ran
constexpr int kTraceCtas = 8192;
__device__ unsigned long long g_lat_trace[4][3 * kTraceCtas];  // [launch & 3][cta] = start ns, end ns, SM id
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif

template <int W, int MINB>
__global__ void __launch_bounds__(128, MINB) half_sweep_fast_kernel(SweepParams P) {
#ifdef TSU_LATTICE_TRACE
  const unsigned long long t_start = trace_now();
#endif
  tsu_fast::half_sweep_fast_body<W>(P);
#ifdef TSU_LATTICE_TRACE
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x < kTraceCtas) {
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long* t = g_lat_trace[(P.sweep * 2 + P.colour) & 3] + 3 * blockIdx.x;
    t[0] = t_start;
    t[1] = trace_now();
    t[2] = smid;
  }
#endif
}

// One word of (replica rep, colour, local row i): any size, open or periodic edges, ragged last word.
__device__ __forceinline__ void generic_update_one(const SweepParams& P, int colour, uint32_t sweep, int rep, int i, int w) {
  const Geom& g = P.g;
  const size_t plane = (size_t)g.rows * g.wpr;
  uint32_t* own = P.state + ((size_t)rep * 2 + colour) * plane;
  Planes pl;
  pl.opp = P.state + ((size_t)rep * 2 + (1 - colour)) * plane;
  pl.halo_top = P.halo_top ? P.halo_top + (size_t)rep * g.wpr : nullptr;
  pl.halo_bot = P.halo_bot ? P.halo_bot + (size_t)rep * g.wpr : nullptr;
  const uint32_t* lut = P.lut + (P.lut_index ? (size_t)P.lut_index[rep] * 32 : 0);

  const Hood h = load_hood(pl, g, colour, i, w);
  if (h.valid == 0u) return;  // padding word (stays 0)
  const int d_row = 2 + h.has_n + h.has_s;
  LutRegs L;
  load_lut_regs(L, lut, d_row);
  Coords q;
  q.colour = (uint32_t)colour;
  q.sweep = sweep;
  q.replica = P.replica0 + (uint32_t)rep;
  q.k0 = P.k0;
  q.k1 = P.k1;
  own[(size_t)i * g.wpr + w] = update_word<false>(h.n, h.s, h.c, h.side, L, lut, d_row, h.valid, h.missW, h.missE, q,
                                                  (uint32_t)w, (uint32_t)(g.row0 + i));
}

// Generic path: one thread per word, one launch per half-sweep.
__global__ void __launch_bounds__(128) half_sweep_generic_kernel(SweepParams P) {
  // rows [P.row_begin, P.row_end) of every replica
  const Geom& g = P.g;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_rep = (long long)(P.row_end - P.row_begin) * g.wpr;
  if (tid >= per_rep * g.n_replicas) return;
  const int rep = (int)(tid / per_rep);
  const int rem = (int)(tid - (long long)rep * per_rep);
  const int i = rem / g.wpr;
  generic_update_one(P, P.colour, P.sweep, rep, P.row_begin + i, rem - i * g.wpr);
}

// Small lattices (C1: 50 x 50): the whole replica belongs to ONE thread block, which runs both colours of
// n_sweeps sweeps in a single launch with a block barrier between half-sweeps.  Same words, same Philox
// coordinates, same bits as the per-half-sweep launches - it only removes 2 n_sweeps - 1 kernel launches,
// which is all the time there is at this size.
constexpr int kResidentThreads = 256;
__global__ void __launch_bounds__(kResidentThreads) sweeps_resident_kernel(SweepParams P, int n_sweeps) {
  const Geom& g = P.g;
  const int rep = blockIdx.x;
  const int per_rep = g.rows * g.wpr;
  for (int t = 0; t < n_sweeps; ++t) {
    for (int colour = 0; colour < 2; ++colour) {
      for (int rem = threadIdx.x; rem < per_rep; rem += kResidentThreads) {
        const int i = rem / g.wpr;
        generic_update_one(P, colour, P.sweep + (uint32_t)t, rep, i, rem - i * g.wpr);
      }
      __syncthreads();  // the other colour reads what this half-sweep wrote (same SM: L1 is coherent for its own stores)
    }
  }
}

// Open boundaries and / or ragged rows: the wide kernel updates, in every row that has both vertical neighbours, the
// 4-word groups whose lanes all exist, as if the columns wrapped at a word boundary; this pass then recomputes, with the true geometry and degree tables, the rim it
// got wrong or skipped: rows [0, rb) and [re, rows) completely and, for open columns, the first and last word of the
// rows in between.  Both passes read only the other colour, so the order of the two launches is the only dependency.
struct RimRows {
  int a0, a1;  // rows [a0, a1) and [b0, b1) are recomputed completely
  int b0, b1;
  int p0, p1;  // of the rows [p0, p1): word 0 if `head`, and the words tail_begin .. wpr-1
  int head, tail_begin;
};

__global__ void __launch_bounds__(128) half_sweep_rim_kernel(SweepParams P, RimRows R) {
  const Geom& g = P.g;
  const int na = R.a1 - R.a0, nb = R.b1 - R.b0;
  const int full_rows = na + nb;
  const int per_row = R.head + (g.wpr - R.tail_begin);
  const long long per_rep = (long long)full_rows * g.wpr + (long long)per_row * (R.p1 - R.p0);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= per_rep * g.n_replicas) return;
  const int rep = (int)(tid / per_rep);
  const int rem = (int)(tid - (long long)rep * per_rep);
  int i, w;
  if (rem < full_rows * g.wpr) {
    const int r = rem / g.wpr;
    w = rem - r * g.wpr;
    i = r < na ? R.a0 + r : R.b0 + (r - na);
  } else {
    const int e = rem - full_rows * g.wpr;
    const int r = e / per_row, k = e - r * per_row;
    i = R.p0 + r;
    w = (R.head && k == 0) ? 0 : R.tail_begin + (k - R.head);
  }
  generic_update_one(P, P.colour, P.sweep, rep, i, w);
}

// Parity mode: uniforms injected per site.  One thread per word, one lane at a time.
__global__ void __launch_bounds__(128) half_sweep_injected_kernel(SweepParams P, const uint32_t* __restrict__ uniforms) {
  const Geom& g = P.g;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_rep = (long long)g.rows * g.wpr;
  if (tid >= per_rep * g.n_replicas) return;
  const int rep = (int)(tid / per_rep);
  const int rem = (int)(tid - (long long)rep * per_rep);
  const int i = rem / g.wpr;
  const int w = rem - i * g.wpr;
  const size_t plane = (size_t)g.rows * g.wpr;
  uint32_t* own = P.state + ((size_t)rep * 2 + P.colour) * plane;
  Planes pl;
  pl.opp = P.state + ((size_t)rep * 2 + (1 - P.colour)) * plane;
  pl.halo_top = P.halo_top ? P.halo_top + (size_t)rep * g.wpr : nullptr;
  pl.halo_bot = P.halo_bot ? P.halo_bot + (size_t)rep * g.wpr : nullptr;
  const uint32_t* lut = P.lut + (P.lut_index ? (size_t)P.lut_index[rep] * 32 : 0);
  const Hood h = load_hood(pl, g, P.colour, i, w);
  if (h.valid == 0u) return;
  const int p = (g.row0 + i + P.colour) & 1;
  const uint32_t* urow = uniforms + ((size_t)rep * g.rows + i) * g.cols;
  uint32_t out = 0u;
  uint32_t m = h.valid;
  while (m) {
    const int j = __ffs(m) - 1;
    m &= m - 1u;
    const int up = ((h.n >> j) & 1u) + ((h.s >> j) & 1u) + ((h.c >> j) & 1u) + ((h.side >> j) & 1u);
    const int d = 2 + h.has_n + h.has_s - (int)((h.missW >> j) & 1u) - (int)((h.missE >> j) & 1u);
    const int col = 2 * (32 * w + j) + p;
    out |= lane_accept(lut, d, up, urow[col]) << j;
  }
  own[(size_t)i * g.wpr + w] = out;
}

__global__ void init_random_kernel(uint32_t* state, Geom g, uint32_t replica0, uint32_t k0, uint32_t k1) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_rep = 2LL * g.rows * g.wpr;
  if (tid >= per_rep * g.n_replicas) return;
  const int rep = (int)(tid / per_rep);
  long long rem = tid - (long long)rep * per_rep;
  const int colour = (int)(rem / ((long long)g.rows * g.wpr));
  rem -= (long long)colour * g.rows * g.wpr;
  const int i = (int)(rem / g.wpr);
  const int w = (int)(rem - (long long)i * g.wpr);
  const int p = (g.row0 + i + colour) & 1;
  const uint32_t valid = lane_mask_lt(colour_count(g.cols, p) - 32 * w);
  tsu_u32x4 o = tsu_lattice_philox((uint32_t)w, (uint32_t)colour, TSU_KIND_INIT, (uint32_t)(g.row0 + i), 0u,
                                   replica0 + (uint32_t)rep, k0, k1);
  state[tid] = o.x & valid;
}

__global__ void pack_kernel(const int8_t* __restrict__ spins, uint32_t* state, Geom g) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_rep = 2LL * g.rows * g.wpr;
  if (tid >= per_rep * g.n_replicas) return;
  const int rep = (int)(tid / per_rep);
  long long rem = tid - (long long)rep * per_rep;
  const int colour = (int)(rem / ((long long)g.rows * g.wpr));
  rem -= (long long)colour * g.rows * g.wpr;
  const int i = (int)(rem / g.wpr);
  const int w = (int)(rem - (long long)i * g.wpr);
  const int p = (g.row0 + i + colour) & 1;
  const int nk = colour_count(g.cols, p);
  const int8_t* row = spins + ((size_t)rep * g.rows + i) * g.cols;
  uint32_t x = 0u;
  for (int j = 0; j < 32; ++j) {
    int k = 32 * w + j;
    if (k < nk && row[2 * k + p] > 0) x |= 1u << j;
  }
  state[tid] = x;
}

__global__ void unpack_kernel(const uint32_t* __restrict__ state, int8_t* spins, Geom g, int as_pm1) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_rep = 2LL * g.rows * g.wpr;
  if (tid >= per_rep * g.n_replicas) return;
  const int rep = (int)(tid / per_rep);
  long long rem = tid - (long long)rep * per_rep;
  const int colour = (int)(rem / ((long long)g.rows * g.wpr));
  rem -= (long long)colour * g.rows * g.wpr;
  const int i = (int)(rem / g.wpr);
  const int w = (int)(rem - (long long)i * g.wpr);
  const int p = (g.row0 + i + colour) & 1;
  const int nk = colour_count(g.cols, p);
  int8_t* row = spins + ((size_t)rep * g.rows + i) * g.cols;
  const uint32_t x = state[tid];
  for (int j = 0; j < 32; ++j) {
    int k = 32 * w + j;
    if (k < nk) {
      int b = (x >> j) & 1u;
      row[2 * k + p] = (int8_t)(as_pm1 ? 2 * b - 1 : b);
    }
  }
}

// up-spin count and anti-aligned (right + down) bond count per replica
__global__ void __launch_bounds__(256) observables_kernel(const uint32_t* __restrict__ state, Geom g,
                                                         const uint32_t* __restrict__ next_rows,
                                                         unsigned long long* out) {
  const int rep = blockIdx.y;
  const size_t plane = (size_t)g.rows * g.wpr;
  const long long n_words = 2LL * g.rows * g.wpr;
  unsigned long long ups = 0, anti = 0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_words;
       t += (long long)gridDim.x * blockDim.x) {
    const int colour = (int)(t / ((long long)g.rows * g.wpr));
    const long long rem = t - (long long)colour * g.rows * g.wpr;
    const int i = (int)(rem / g.wpr);
    const int w = (int)(rem - (long long)i * g.wpr);
    const uint32_t* own = state + ((size_t)rep * 2 + colour) * plane;
    Planes pl;
    pl.opp = state + ((size_t)rep * 2 + (1 - colour)) * plane;
    pl.halo_top = nullptr;
    pl.halo_bot = next_rows ? next_rows + ((size_t)rep * 2 + (1 - colour)) * g.wpr : nullptr;
    const Hood h = load_hood(pl, g, colour, i, w);
    if (h.valid == 0u) continue;
    const uint32_t x = own[(size_t)i * g.wpr + w];
    const int p = (g.row0 + i + colour) & 1;
    ups += __popc(x & h.valid);
    // east neighbour: centre word if own col is even (p == 0), shifted word otherwise
    const uint32_t east = p ? h.side : h.c;
    anti += __popc((x ^ east) & h.valid & ~h.missE);
    if (h.has_s) anti += __popc((x ^ h.s) & h.valid);
  }
  ups = tsu_warp_sum(ups);
  anti = tsu_warp_sum(anti);
  if ((threadIdx.x & 31) == 0) {
    if (ups) atomicAdd(out + 2 * rep, ups);
    if (anti) atomicAdd(out + 2 * rep + 1, anti);
  }
}

// Same counts for lattices with full words (cols % 256 == 0; periodic or open), at HBM speed: a thread owns a 4-word column
// strip of BOTH colour planes and walks down its rows with a rolling (row i, row i+1) window, so every word is
// loaded once as a 16-byte vector.  In a row exactly one colour has odd own columns (east neighbour = next lane of
// the other plane, i.e. a 1-bit funnel shift that needs one extra word); the other colour's east neighbour is the
// same lane.  The south neighbour of (colour, row i, lane) is (other colour, row i+1, same lane).
__global__ void __launch_bounds__(128) observables_fast_kernel(const uint32_t* __restrict__ state, Geom g,
                                                              const uint32_t* __restrict__ next_rows, int strip_rows,
                                                              int n_strips, unsigned long long* out) {
  const int nvec = g.wpr >> 2;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_rep = n_strips * nvec;
  const int per_rep_pad = (per_rep + 31) & ~31;  // warps never straddle replicas
  const int rep = (int)(tid / per_rep_pad);
  if (rep >= g.n_replicas) return;
  const int rem = (int)(tid - (long long)rep * per_rep_pad);
  unsigned ups = 0, anti = 0;
  if (rem < per_rep) {
    const int strip = rem / nvec, w0 = 4 * (rem - strip * nvec);
    const int r_begin = strip * strip_rows, r_end = min(g.rows, r_begin + strip_rows);
    const int w_next = (w0 + 4 == g.wpr) ? 0 : w0 + 4;
    const size_t plane = (size_t)g.rows * g.wpr;
    const uint32_t* pl[2] = {state + (size_t)rep * 2 * plane, state + ((size_t)rep * 2 + 1) * plane};
    Planes opp[2];  // opp[c] = the plane that is NOT colour c (south rows beyond the local rows come from next_rows)
    for (int c = 0; c < 2; ++c) {
      opp[c].opp = pl[1 - c];
      opp[c].halo_top = nullptr;
      opp[c].halo_bot = next_rows ? next_rows + ((size_t)rep * 2 + (1 - c)) * g.wpr : nullptr;
    }
    uint4 a[2];
    a[0] = *reinterpret_cast<const uint4*>(pl[0] + (size_t)r_begin * g.wpr + w0);
    a[1] = *reinterpret_cast<const uint4*>(pl[1] + (size_t)r_begin * g.wpr + w0);
    for (int i = r_begin; i < r_end; ++i) {
      // row i + 1 of both planes (the south neighbours), or nothing below an open last row
      const uint32_t* s1 = opp_row(opp[0], g, i + 1);  // plane 1 at row i+1 = south of colour 0
      const uint32_t* s0 = opp_row(opp[1], g, i + 1);  // plane 0 at row i+1 = south of colour 1
      uint4 b[2] = {a[0], a[1]};
      if (s0) b[0] = *reinterpret_cast<const uint4*>(s0 + w0);
      if (s1) b[1] = *reinterpret_cast<const uint4*>(s1 + w0);
      const int odd = 1 - ((g.row0 + i) & 1);  // colour `odd` has odd own columns in this row: col = 2 lane + 1
      const uint32_t extra = pl[1 - odd][(size_t)i * g.wpr + w_next];  // next word of the plane the odd colour looks at
      const uint32_t x0[4] = {a[0].x, a[0].y, a[0].z, a[0].w}, x1[4] = {a[1].x, a[1].y, a[1].z, a[1].w};
      const uint32_t* own_odd = odd ? x1 : x0;   // colour with the shifted east neighbour
      const uint32_t* own_even = odd ? x0 : x1;
      // (own_even is also the plane the odd colour looks at, and vice versa)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t nxt = k < 3 ? own_even[k + 1] : extra;
        const uint32_t east_odd = __funnelshift_r(own_even[k], nxt, 1);
        // open columns: the last site of a row with odd own columns has no east neighbour
        const uint32_t has_east = (k == 3 && !g.wrap_cols && w0 + 4 == g.wpr) ? 0x7fffffffu : 0xffffffffu;
        ups += __popc(x0[k]) + __popc(x1[k]);
        anti += __popc((own_odd[k] ^ east_odd) & has_east) + __popc(own_even[k] ^ own_odd[k]);
      }
      if (s1) anti += __popc(a[0].x ^ b[1].x) + __popc(a[0].y ^ b[1].y) + __popc(a[0].z ^ b[1].z) + __popc(a[0].w ^ b[1].w);
      if (s0) anti += __popc(a[1].x ^ b[0].x) + __popc(a[1].y ^ b[0].y) + __popc(a[1].z ^ b[0].z) + __popc(a[1].w ^ b[0].w);
      a[0] = b[0];
      a[1] = b[1];
    }
  }
  const unsigned long long u64 = tsu_warp_sum((unsigned long long)ups), a64 = tsu_warp_sum((unsigned long long)anti);
  if ((threadIdx.x & 31) == 0) {
    if (u64) atomicAdd(out + 2 * rep, u64);
    if (a64) atomicAdd(out + 2 * rep + 1, a64);
  }
}

__global__ void energy_from_obs_kernel(const unsigned long long* __restrict__ obs, int n, double J, double h,
                                       long long n_bonds, long long n_sites, double* energy) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const double up = (double)obs[2 * r];
  const double anti = (double)obs[2 * r + 1];
  energy[r] = -J * ((double)n_bonds - 2.0 * anti) - h * (2.0 * up - (double)n_sites);
}

// ---- row slabs with peer-mapped halos (tsu_ising2d_slab_sweeps_p2p) ---------------------------------------
// copy one row of `colour` of every replica into a halo buffer that may live on another GPU (NVLink peer mapping)
__global__ void __launch_bounds__(256) slab_send_rows_kernel(const uint32_t* __restrict__ state, int n_replicas, int rows,
                                                           int wpr, int colour, uint32_t* up_dst, uint32_t* down_dst) {
  // up_dst   <- my first row (the row below the upper neighbour's last row)
  // down_dst <- my last row  (the row above the lower neighbour's first row)
  const int nvec = wpr >> 2;
  const long long per_side = (long long)n_replicas * nvec;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < 2 * per_side; t += (long long)gridDim.x * blockDim.x) {
    const int side = t >= per_side;
    uint32_t* dst = side ? down_dst : up_dst;
    if (!dst) continue;
    const long long u = t - side * per_side;
    const int rep = (int)(u / nvec), v = (int)(u - (long long)rep * nvec);
    const size_t row = side ? (size_t)(rows - 1) : 0;
    const uint4 x = *reinterpret_cast<const uint4*>(state + (((size_t)rep * 2 + colour) * rows + row) * wpr + 4 * v);
    *reinterpret_cast<uint4*>(dst + (size_t)rep * wpr + 4 * v) = x;
  }
}

// the rows sent before this kernel (stream order) are complete: publish their message number to the receivers
__global__ void slab_signal_kernel(uint32_t* up_flag, uint32_t* down_flag, uint32_t value) {
  __threadfence_system();
  if (up_flag) *reinterpret_cast<volatile uint32_t*>(up_flag) = value;
  if (down_flag) *reinterpret_cast<volatile uint32_t*>(down_flag) = value;
}

// hold the stream until both neighbours' rows with message number >= need have arrived (they write the counters
// through their peer mapping).  Gives up after ~20 s and raises the status word instead of hanging the GPU.
__global__ void slab_wait_kernel(const uint32_t* flag_a, const uint32_t* flag_b, uint32_t need, uint32_t* status) {
  const long long t0 = clock64();
  const volatile uint32_t* fa = flag_a;
  const volatile uint32_t* fb = flag_b;
  while (true) {
    const bool a_ok = !fa || (int32_t)(*fa - need) >= 0;
    const bool b_ok = !fb || (int32_t)(*fb - need) >= 0;
    if (a_ok && b_ok) return;
    if (clock64() - t0 > 40000000000LL) {
      *status = 1u;
      return;
    }
    __nanosleep(200);
  }
}

// streams the slab driver launches its extra interior row ranges on: created once per device, never destroyed
std::mutex g_slab_stream_mutex;
std::map<int, std::vector<cudaStream_t>> g_slab_streams;
cudaStream_t slab_stream(int dev, int idx) {
  std::lock_guard<std::mutex> lock(g_slab_stream_mutex);
  std::vector<cudaStream_t>& v = g_slab_streams[dev];
  while ((int)v.size() <= idx) {
    cudaStream_t s = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    v.push_back(s);
  }
  return v[idx];
}

bool geom_ok(int n_replicas, int rows, int cols) {
  // counter word 1 of the lattice stream keeps the global row in 24 bits, counter word 0 the 4-word group in 24
  return n_replicas > 0 && rows > 0 && cols > 0 && cols < (1 << 26) && rows <= TSU_LATTICE_MAX_ROWS;
}

bool rows_ok(int rows, int row0) { return row0 >= 0 && (long long)row0 + rows <= TSU_LATTICE_MAX_ROWS; }

// Tuning knobs (never change results), read once per process:
//   TSU_LATTICE_STRIP  rows per thread strip of the wide kernel (default: 128, shortened until the grid fills the GPU)
//   TSU_LATTICE_SPLIT  row ranges a half-sweep is cut into (split_ranges(); 1 = one launch per half-sweep)
//   TSU_LATTICE_W      words per thread of the prebuilt wide kernel (2 or 4)
//   TSU_LATTICE_OPEN_GENERIC / TSU_LATTICE_OBS_GENERIC / TSU_LATTICE_RESIDENT=0  force the one-thread-per-word kernels
//   TSU_JIT_THREADS (32 / 64 / 128 threads per CTA),
//   TSU_JIT_W, TSU_JIT_MINB, TSU_JIT_UNROLL   words per thread, CTAs per SM and row-loop unrolling the run-time
//                                             specialised kernel is compiled for
struct Tuning {
  int strip = 0, w = 0, open_generic = 0, obs_generic = 0, resident = 1, jit_w = 0, jit_minb = 0, jit_unroll = 0, jit_threads = 0, split = -1;
  Tuning() {
    auto num = [](const char* name, int dflt) {
      const char* e = getenv(name);
      return e ? atoi(e) : dflt;
    };
    strip = num("TSU_LATTICE_STRIP", 0);
    w = num("TSU_LATTICE_W", 0);
    open_generic = getenv("TSU_LATTICE_OPEN_GENERIC") != nullptr;
    obs_generic = getenv("TSU_LATTICE_OBS_GENERIC") != nullptr;
    resident = num("TSU_LATTICE_RESIDENT", 1);
    jit_w = num("TSU_JIT_W", 0);
    jit_minb = num("TSU_JIT_MINB", 0);
    jit_unroll = num("TSU_JIT_UNROLL", 0);
    jit_threads = num("TSU_JIT_THREADS", 0);
    split = num("TSU_LATTICE_SPLIT", -1);
  }
};
Tuning g_tuning;  // read when the library is loaded; tsu_ising2d_reload_tuning() reads the environment again
const Tuning& tuning() { return g_tuning; }
constexpr int kDefaultW = 4, kDefaultJitW = 4, kDefaultJitMinB = 5, kDefaultJitThreads = 128;
constexpr int kDefaultSplit = 8, kMaxSplit = 8;  // row ranges of one half-sweep launched on streams of their own  // 5 CTAs/SM (93 registers): +1 % over 4, measured

Geom make_geom(int n_replicas, int rows, int cols, int wrap_rows, int wrap_cols, int row0) {
  Geom g;
  g.rows = rows;
  g.cols = cols;
  g.wpr = words_per_row(cols);
  g.wrap_rows = wrap_rows ? 1 : 0;
  g.wrap_cols = wrap_cols ? 1 : 0;
  g.row0 = row0;
  g.n_replicas = n_replicas;
  return g;
}

unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// ---- run-time specialisation of the fast kernel (NVRTC) ------------------------------------------------
// For a launch whose replicas all share one threshold table, ising2d_fast.cuh is compiled once per table set
// with the eight 5-bit truth tables as literals.  libnvrtc and libcuda are opened with dlopen so that the
// library has no link-time dependency on them; any failure simply leaves the prebuilt jump-table kernel in use.
std::mutex g_jit_mutex;
std::map<std::string, int> g_jit_cache;  // "device:t0,..,t7" -> handle
struct JitKernel {
  void* fn;
  int w;        // words per thread it was compiled for
  int threads;  // threads per CTA it was compiled for
};
std::vector<JitKernel> g_jit_functions;  // handle - 1 -> kernel

JitKernel jit_function(int handle) {
  std::lock_guard<std::mutex> lock(g_jit_mutex);
  if (handle < 1 || handle > (int)g_jit_functions.size()) return JitKernel{nullptr, 0, 0};
  return g_jit_functions[handle - 1];
}

int launch_half_sweep(uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_rows, int wrap_cols, int colour,
                      const uint32_t* d_lut, const int32_t* d_lut_index, uint64_t seed, uint32_t sweep,
                      uint32_t replica0, int row0, const uint32_t* d_halo_top, const uint32_t* d_halo_bot,
                      cudaStream_t st, JitKernel jit = JitKernel{nullptr, 0, 0}, int upd_begin = 0, int upd_end = -1,
                      int pieces = 1) {
  // pieces: this launch is one of `pieces` row ranges of a half-sweep that run concurrently (split_ranges())
  if (upd_end < 0) upd_end = rows;  // local rows [upd_begin, upd_end) are updated (default: all)
  if (upd_begin >= upd_end) return TSU_OK;
  SweepParams P;
  P.state = d_state;
  P.lut = d_lut;
  P.lut_index = d_lut_index;
  P.halo_top = d_halo_top;
  P.halo_bot = d_halo_bot;
  P.g = make_geom(n_replicas, rows, cols, wrap_rows, wrap_cols, row0);
  P.colour = colour;
  P.sweep = sweep;
  P.replica0 = replica0;
  P.k0 = (uint32_t)seed;
  P.k1 = (uint32_t)(seed >> 32);
  P.keys = tsu_philox_key_schedule(P.k0, P.k1);
  P.row_begin = 0;
  P.row_end = rows;
  P.nvec_fast = P.g.wpr / 4;
  // rows that have a north / south neighbour (exactly the cases opp_row() resolves): the wide kernel takes those,
  // columns treated as periodic; what that gets wrong on open lattices is redone by the rim pass
  const bool north_ok = d_halo_top || wrap_rows, south_ok = d_halo_bot || wrap_rows;
  const int rb0 = north_ok ? 0 : 1, re0 = south_ok ? rows : rows - 1;
  const int rb = rb0 > upd_begin ? rb0 : upd_begin, re = re0 < upd_end ? re0 : upd_end;  // rows of the wide kernel
  // words that are full for both row parities: floor(cols / 2) lanes exist in every row of either colour
  const int full_words = (cols / 2) / 32;
  const bool ragged = cols % 256 != 0;
  const int nvec_f = ragged ? full_words / 4 : P.g.wpr / 4;
  const bool need_rim = rb > upd_begin || re < upd_end || !wrap_cols || ragged;
  bool fast = nvec_f >= 1 && re - rb >= 1;
  if (need_rim && tuning().open_generic) fast = false;
  if (fast) {
    const bool use_jit = jit.fn && !d_lut_index;
    const int W = use_jit ? jit.w : (tuning().w == 2 ? 2 : kDefaultW);
    const int nvec = nvec_f * (4 / W);  // W-word groups per row
    const int frows = re - rb;
    P.nvec_fast = nvec_f;
    // strips long enough to amortise the two halo rows and the warp prologue, short enough to fill 148 SMs x 16 warps
    const long long target_threads = 148LL * 2048;
    int strip = 128;
    const long long srows = (long long)frows * pieces;
    while (strip > 1 && (long long)n_replicas * nvec * ((srows + strip - 1) / strip) < target_threads) strip >>= 1;
    // the tail of one range is filled by the next: longer strips (less set-up per row) cost nothing there
    if (pieces >= 4) strip = strip * 4 < 128 ? strip * 4 : 128;
    if (tuning().strip > 0) strip = tuning().strip;
    P.strip_rows = strip;
    P.n_strips = (frows + strip - 1) / strip;
    P.row_begin = rb;
    P.row_end = re;
    const long long per_rep_pad = ((long long)P.n_strips * nvec + 31) / 32 * 32;  // warps never straddle replicas
    const long long total = (long long)n_replicas * per_rep_pad;
    const unsigned grid = blocks_for(total, 128);
    if (use_jit) {  // table-specialised build of the same kernel body
      void* args[] = {&P};
      if (tsu_jit::launch(jit.fn, blocks_for(total, jit.threads), jit.threads, 0, (void*)st, args) != 0) return 999;  // driver-API launch failure
    } else if (W == 2) {
      half_sweep_fast_kernel<2, 8><<<grid, 128, 0, st>>>(P);
    } else {
      half_sweep_fast_kernel<4, 4><<<grid, 128, 0, st>>>(P);
    }
    if (need_rim) {
      RimRows R;
      R.a0 = upd_begin; R.a1 = rb;   // updated rows above / below the wide kernel's range (no north / south neighbour)
      R.b0 = re; R.b1 = upd_end;
      R.p0 = rb; R.p1 = re;
      R.head = (ragged || !wrap_cols) ? 1 : 0;
      R.tail_begin = ragged ? 4 * nvec_f : (wrap_cols ? P.g.wpr : P.g.wpr - 1);
      const long long per_rep = (long long)((R.a1 - R.a0) + (R.b1 - R.b0)) * P.g.wpr +
                                (long long)(R.head + P.g.wpr - R.tail_begin) * frows;
      if (per_rep > 0) half_sweep_rim_kernel<<<blocks_for(per_rep * n_replicas, 128), 128, 0, st>>>(P, R);
    }
  } else {
    P.strip_rows = 1;
    P.n_strips = rows;
    P.row_begin = upd_begin;
    P.row_end = upd_end;
    const long long total = (long long)n_replicas * (upd_end - upd_begin) * P.g.wpr;
    half_sweep_generic_kernel<<<blocks_for(total, 128), 128, 0, st>>>(P);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TSU_OK : (int)e;
}

// Row ranges a half-sweep is cut into.  One launch per half-sweep loses ~15 us to its tail (launch gap, the last warps
// finishing alone: measured, tools/lattice_trace.py); ranges on streams of their own, each depending only on its
// neighbours' previous half-sweep, let the next half-sweep start while this one drains.  Pays between ~0.1 and ~5 ms
// per half-sweep (6.8 % on a 16384 x 131072 slab, 2.5 % on 131072^2); TSU_LATTICE_SPLIT overrides (1 = never).
int split_ranges(int n_replicas, int rows, int cols) {
  int K;
  if (tuning().split >= 1) {
    K = tuning().split < kMaxSplit ? tuning().split : kMaxSplit;
  } else {
    const double sites = (double)n_replicas * rows * cols;
    K = (sites >= 1e9 && sites <= 6e10) ? kDefaultSplit : 1;
  }
  while (K > 1 && rows / K < 256) --K;  // short ranges would only add launches
  return K;
}

// Chunks of replicas a batch of lattices is cut into for the same reason (replicas do not depend on each other, so
// every chunk simply runs its half-sweeps back to back on its own stream).  TSU_LATTICE_SPLIT=1 disables.
int split_replicas(int n_replicas, int rows, int cols) {
  if (tuning().split == 1 || n_replicas < 2 * kDefaultSplit) return 1;
  const double sites = (double)n_replicas * rows * cols;
  if (tuning().split < 1 && !(sites >= 5e8 && sites <= 6e10)) return 1;
  return tuning().split > 1 ? (tuning().split < kMaxSplit ? tuning().split : kMaxSplit) : kDefaultSplit;
}

// the whole-lattice sweeps of tsu_ising2d_sweeps / _sweeps_jit (no halos)
int launch_sweeps(uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_rows, int wrap_cols, const uint32_t* d_lut,
                  const int32_t* d_lut_index, uint64_t seed, uint32_t sweep0, int n_sweeps, uint32_t replica0,
                  cudaStream_t st, JitKernel jit) {
  // many replicas: independent chunks of replicas, one stream each, no dependency between them at all
  const int KR = split_replicas(n_replicas, rows, cols);
  if (KR > 1 && n_sweeps > 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaEvent_t ev_join;
    if (cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess) return (int)cudaGetLastError();
    cudaEventRecord(ev_join, st);
    const size_t rep_words = 2 * (size_t)rows * words_per_row(cols);
    int rc = TSU_OK;
    for (int k = 0; k < KR && rc == TSU_OK; ++k) {
      cudaStream_t sk = k == 0 ? st : slab_stream(dev, k - 1);
      if (k > 0 && !sk) {  // (the caller's stream may be the default stream: a null handle)
        rc = (int)cudaErrorUnknown;
        break;
      }
      if (k > 0) cudaStreamWaitEvent(sk, ev_join, 0);
      const int r0 = (int)((long long)n_replicas * k / KR), r1 = (int)((long long)n_replicas * (k + 1) / KR);
      for (int t = 0; t < 2 * n_sweeps && rc == TSU_OK; ++t)
        rc = launch_half_sweep(d_state + (size_t)r0 * rep_words, r1 - r0, rows, cols, wrap_rows, wrap_cols, t & 1, d_lut,
                               d_lut_index ? d_lut_index + r0 : nullptr, seed, sweep0 + (uint32_t)(t >> 1),
                               replica0 + (uint32_t)r0, 0, nullptr, nullptr, sk, jit, 0, -1, KR);
    }
    for (int k = 1; k < KR; ++k) {  // also after an error: the caller's stream must not run ahead of what was launched
      cudaStream_t sk = slab_stream(dev, k - 1);
      if (!sk) break;
      cudaEventRecord(ev_join, sk);
      cudaStreamWaitEvent(st, ev_join, 0);
    }
    cudaEventDestroy(ev_join);
    if (rc != TSU_OK) return rc;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? TSU_OK : (int)e;
  }
  const int K = split_ranges(n_replicas, rows, cols);
  if (K <= 2 || n_sweeps == 0) {
    for (int t = 0; t < 2 * n_sweeps; ++t) {
      int rc = launch_half_sweep(d_state, n_replicas, rows, cols, wrap_rows, wrap_cols, t & 1, d_lut, d_lut_index, seed,
                                 sweep0 + (uint32_t)(t >> 1), replica0, 0, nullptr, nullptr, st, jit);
      if (rc != TSU_OK) return rc;
    }
    return TSU_OK;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaStream_t sub[kMaxSplit];
  sub[0] = st;
  for (int k = 1; k < K; ++k) {
    sub[k] = slab_stream(dev, k - 1);
    if (!sub[k]) return (int)cudaGetLastError();
  }
  int bound[kMaxSplit + 1];
  for (int k = 0; k <= K; ++k) bound[k] = (int)((long long)rows * k / K);
  cudaEvent_t ev[kMaxSplit][2], ev_join;
  bool ev_ok = cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) == cudaSuccess;
  for (int k = 0; k < K && ev_ok; ++k)
    for (int j = 0; j < 2 && ev_ok; ++j) ev_ok = cudaEventCreateWithFlags(&ev[k][j], cudaEventDisableTiming) == cudaSuccess;
  if (!ev_ok) return (int)cudaGetLastError();
  cudaEventRecord(ev_join, st);
  for (int k = 1; k < K; ++k) cudaStreamWaitEvent(sub[k], ev_join, 0);
  int rc = TSU_OK;
  for (int t = 0; t < 2 * n_sweeps && rc == TSU_OK; ++t) {
    const int j = t & 1, jp = j ^ 1;
    // the range launched first changes every half-sweep: with periodic rows range 0 depends on range K - 1, and what
    // the first range of a half-sweep depends on should be what the half-sweep before launched first
    for (int i = 0; i < K && rc == TSU_OK; ++i) {
      const int k = (t + i) % K, above = (k + K - 1) % K, below = (k + 1) % K;
      if (t > 0) {
        if (k > 0 || wrap_rows) cudaStreamWaitEvent(sub[k], ev[above][jp], 0);
        if (k + 1 < K || wrap_rows) cudaStreamWaitEvent(sub[k], ev[below][jp], 0);
      }
      rc = launch_half_sweep(d_state, n_replicas, rows, cols, wrap_rows, wrap_cols, t & 1, d_lut, d_lut_index, seed,
                             sweep0 + (uint32_t)(t >> 1), replica0, 0, nullptr, nullptr, sub[k], jit, bound[k], bound[k + 1], K);
      cudaEventRecord(ev[k][j], sub[k]);
    }
  }
  for (int k = 1; k < K; ++k) {  // the caller's stream continues after everything issued here
    cudaEventRecord(ev_join, sub[k]);
    cudaStreamWaitEvent(st, ev_join, 0);
  }
  for (int k = 0; k < K; ++k)
    for (int j = 0; j < 2; ++j) cudaEventDestroy(ev[k][j]);
  cudaEventDestroy(ev_join);
  if (rc != TSU_OK) return rc;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TSU_OK : (int)e;
}

// true when a replica is small enough that per-half-sweep launches would be pure launch latency
bool resident_eligible(int n_replicas, int rows, int cols) {
  if (tuning().resident == 0) return false;
  const long long per_rep = (long long)rows * words_per_row(cols);
  // <= 16 words per thread and half-sweep, and not enough replicas to fill the GPU with the wide kernels
  return per_rep <= 16LL * kResidentThreads && (long long)n_replicas * per_rep <= 148LL * 2048;
}

int launch_resident_sweeps(uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_rows, int wrap_cols,
                           const uint32_t* d_lut, const int32_t* d_lut_index, uint64_t seed, uint32_t sweep0, int n_sweeps,
                           uint32_t replica0, cudaStream_t st) {
  SweepParams P = {};
  P.state = d_state;
  P.lut = d_lut;
  P.lut_index = d_lut_index;
  P.g = make_geom(n_replicas, rows, cols, wrap_rows, wrap_cols, 0);
  P.sweep = sweep0;
  P.replica0 = replica0;
  P.k0 = (uint32_t)seed;
  P.k1 = (uint32_t)(seed >> 32);
  P.strip_rows = 1;
  P.n_strips = rows;
  sweeps_resident_kernel<<<n_replicas, kResidentThreads, 0, st>>>(P, n_sweeps);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TSU_OK : (int)e;
}

}  // namespace

extern "C" {

void tsu_ising2d_reload_tuning(void) { g_tuning = Tuning(); }

#ifdef TSU_LATTICE_TRACE
int tsu_debug_lattice_trace(unsigned long long* h_out) {  // 4 x 3 x 8192 words
  return (int)cudaMemcpyFromSymbol(h_out, g_lat_trace, sizeof(unsigned long long) * 4 * 3 * kTraceCtas);
}
#endif

int64_t tsu_ising2d_words_per_row(int cols) { return cols > 0 ? words_per_row(cols) : 0; }

int64_t tsu_ising2d_state_words(int rows, int cols) {
  return (rows > 0 && cols > 0) ? 2LL * rows * words_per_row(cols) : 0;
}

int tsu_ising2d_init_random(uint32_t* d_state, int n_replicas, int rows, int cols, uint64_t seed, uint32_t replica0,
                            int row0, uintptr_t stream) {
  TSU_CHECK_ARG(d_state && geom_ok(n_replicas, rows, cols) && rows_ok(rows, row0));
  Geom g = make_geom(n_replicas, rows, cols, 0, 0, row0);
  const long long total = 2LL * rows * g.wpr * n_replicas;
  init_random_kernel<<<blocks_for(total, 256), 256, 0, tsu_stream(stream)>>>(d_state, g, replica0, (uint32_t)seed,
                                                                            (uint32_t)(seed >> 32));
  TSU_RETURN_LAUNCH_STATUS();
}

int tsu_ising2d_pack(const int8_t* d_spins, uint32_t* d_state, int n_replicas, int rows, int cols, uintptr_t stream) {
  TSU_CHECK_ARG(d_spins && d_state && geom_ok(n_replicas, rows, cols));
  Geom g = make_geom(n_replicas, rows, cols, 0, 0, 0);
  const long long total = 2LL * rows * g.wpr * n_replicas;
  pack_kernel<<<blocks_for(total, 256), 256, 0, tsu_stream(stream)>>>(d_spins, d_state, g);
  TSU_RETURN_LAUNCH_STATUS();
}

int tsu_ising2d_unpack(const uint32_t* d_state, int8_t* d_spins, int n_replicas, int rows, int cols, int as_pm1,
                       uintptr_t stream) {
  TSU_CHECK_ARG(d_spins && d_state && geom_ok(n_replicas, rows, cols));
  Geom g = make_geom(n_replicas, rows, cols, 0, 0, 0);
  const long long total = 2LL * rows * g.wpr * n_replicas;
  unpack_kernel<<<blocks_for(total, 256), 256, 0, tsu_stream(stream)>>>(d_state, d_spins, g, as_pm1);
  TSU_RETURN_LAUNCH_STATUS();
}

int tsu_ising2d_half_sweep(uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_rows, int wrap_cols,
                           int colour, const uint32_t* d_lut, const int32_t* d_lut_index, uint64_t seed,
                           uint32_t sweep, uint32_t replica0, int row0, const uint32_t* d_halo_top,
                           const uint32_t* d_halo_bot, uintptr_t stream) {
  TSU_CHECK_ARG(d_state && d_lut && geom_ok(n_replicas, rows, cols) && rows_ok(rows, row0));
  TSU_CHECK_ARG(colour == 0 || colour == 1);
  TSU_CHECK_ARG(!wrap_cols || (cols % 2 == 0 && cols > 2));
  TSU_CHECK_ARG(!wrap_rows || (rows % 2 == 0 && rows > 2) || d_halo_top || d_halo_bot);
  return launch_half_sweep(d_state, n_replicas, rows, cols, wrap_rows, wrap_cols, colour, d_lut, d_lut_index, seed,
                           sweep, replica0, row0, d_halo_top, d_halo_bot, tsu_stream(stream));
}

int tsu_ising2d_half_sweep_rows(int jit_handle, uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_rows,
                                int wrap_cols, int colour, const uint32_t* d_lut, const int32_t* d_lut_index, uint64_t seed,
                                uint32_t sweep, uint32_t replica0, int row0, const uint32_t* d_halo_top,
                                const uint32_t* d_halo_bot, int row_begin, int row_end, uintptr_t stream) {
  TSU_CHECK_ARG(d_state && d_lut && geom_ok(n_replicas, rows, cols) && rows_ok(rows, row0));
  TSU_CHECK_ARG(colour == 0 || colour == 1);
  TSU_CHECK_ARG(!wrap_cols || (cols % 2 == 0 && cols > 2));
  TSU_CHECK_ARG(!wrap_rows || (rows % 2 == 0 && rows > 2) || d_halo_top || d_halo_bot);
  TSU_CHECK_ARG(row_begin >= 0 && row_begin <= row_end && row_end <= rows);
  const JitKernel jit = (jit_handle > 0 && !d_lut_index) ? jit_function(jit_handle) : JitKernel{nullptr, 0, 0};
  return launch_half_sweep(d_state, n_replicas, rows, cols, wrap_rows, wrap_cols, colour, d_lut, d_lut_index, seed, sweep,
                           replica0, row0, d_halo_top, d_halo_bot, tsu_stream(stream), jit, row_begin, row_end);
}

int tsu_ising2d_slab_sweeps_p2p(int jit_handle, uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_cols,
                                const uint32_t* d_lut, const int32_t* d_lut_index, uint64_t seed, uint32_t sweep0,
                                int n_sweeps, uint32_t replica0, int row0, uint32_t* d_halo, uint32_t* d_flags,
                                uint32_t* d_up_halo, uint32_t* d_up_flags, uint32_t* d_down_halo, uint32_t* d_down_flags,
                                uint32_t msgs_colour0, uint32_t msgs_colour1, uintptr_t main_stream,
                                uintptr_t side_stream) {
  TSU_CHECK_ARG(d_state && d_lut && d_halo && d_flags && geom_ok(n_replicas, rows, cols) && rows_ok(rows, row0));
  TSU_CHECK_ARG(rows >= 4 && n_sweeps >= 0 && main_stream != side_stream);
  TSU_CHECK_ARG(!wrap_cols || (cols % 2 == 0 && cols > 2));
  TSU_CHECK_ARG((d_up_halo == nullptr) == (d_up_flags == nullptr) && (d_down_halo == nullptr) == (d_down_flags == nullptr));
  const JitKernel jit = (jit_handle > 0 && !d_lut_index) ? jit_function(jit_handle) : JitKernel{nullptr, 0, 0};
  cudaStream_t main = tsu_stream(main_stream), side = tsu_stream(side_stream);
  const int wpr = words_per_row(cols);
  const size_t side_words = (size_t)n_replicas * wpr;  // one halo row set
  auto halo = [&](uint32_t* base, int colour, int which) { return base ? base + ((size_t)colour * 2 + which) * side_words : nullptr; };
  auto flag = [&](uint32_t* base, int colour, int which) { return base ? base + colour * 2 + which : nullptr; };
  uint32_t msgs[2] = {msgs_colour0, msgs_colour1};
  // The interior rows are cut into row ranges launched on streams of their own: range k of a half-sweep depends only on
  // ranges k - 1, k, k + 1 (and, at the ends, on the boundary rows) of the half-sweep before, so the first ranges of
  // the next half-sweep fill the SMs while the last ranges of this one drain.  A single launch per half-sweep loses
  // ~15 us to its tail (launch gap + the last warps finishing alone), 6 % of a 16384 x 131072 slab.
  int dev = 0;
  cudaGetDevice(&dev);
  const int K = split_ranges(n_replicas, rows - 2, cols);
  cudaStream_t sub[kMaxSplit];
  sub[0] = main;
  for (int k = 1; k < K; ++k) {
    sub[k] = slab_stream(dev, k - 1);
    if (!sub[k]) return (int)cudaGetLastError();
  }
  int bound[kMaxSplit + 1];  // range k = local rows [bound[k], bound[k + 1])
  for (int k = 0; k <= K; ++k) bound[k] = 1 + (int)((long long)(rows - 2) * k / K);
  // strips as long as one launch over the whole interior would use
  cudaEvent_t ev_sub[kMaxSplit][2], ev_bnd[2], ev_join;
  bool ev_ok = cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) == cudaSuccess;
  for (int j = 0; j < 2 && ev_ok; ++j) {
    ev_ok = cudaEventCreateWithFlags(&ev_bnd[j], cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < K && ev_ok; ++k) ev_ok = cudaEventCreateWithFlags(&ev_sub[k][j], cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ev_ok) return (int)cudaGetLastError();
  const unsigned send_grid = blocks_for(2LL * n_replicas * (wpr / 4), 256) < 64u ? blocks_for(2LL * n_replicas * (wpr / 4), 256) : 64u;
  // what a half-sweep of `colour` sends afterwards: my first row to the rank above (its "below" halo), my last row
  // to the rank below (its "above" halo), then the message number
  auto send = [&](int colour) {
    ++msgs[colour];
    slab_send_rows_kernel<<<send_grid, 256, 0, side>>>(d_state, n_replicas, rows, wpr, colour, halo(d_up_halo, colour, 1),
                                                      halo(d_down_halo, colour, 0));
    slab_signal_kernel<<<1, 1, 0, side>>>(flag(d_up_flags, colour, 1), flag(d_down_flags, colour, 0), msgs[colour]);
  };
  int rc = TSU_OK;
  // everything the caller queued on the main stream (e.g. a changed state) precedes the first rows sent and updated
  cudaEventRecord(ev_join, main);
  cudaStreamWaitEvent(side, ev_join, 0);
  for (int k = 1; k < K; ++k) cudaStreamWaitEvent(sub[k], ev_join, 0);
  send(1);  // colour 0 goes first and reads colour 1
  for (int t = 0; t < 2 * n_sweeps && rc == TSU_OK; ++t) {
    const int colour = t & 1, opp = 1 - colour, j = t & 1, jp = j ^ 1;
    const uint32_t sweep = sweep0 + (uint32_t)(t >> 1);
    // interior ranges: each reads and overwrites rows next to what its neighbours updated in the half-sweep before
    for (int k = 0; k < K && rc == TSU_OK; ++k) {
      if (t > 0) {
        if (k > 0) cudaStreamWaitEvent(sub[k], ev_sub[k - 1][jp], 0);
        if (k + 1 < K) cudaStreamWaitEvent(sub[k], ev_sub[k + 1][jp], 0);
        if (k == 0 || k + 1 == K) cudaStreamWaitEvent(sub[k], ev_bnd[jp], 0);
      }
      rc = launch_half_sweep(d_state, n_replicas, rows, cols, 0, wrap_cols, colour, d_lut, d_lut_index, seed, sweep, replica0,
                             row0, nullptr, nullptr, sub[k], jit, bound[k], bound[k + 1], K);
      cudaEventRecord(ev_sub[k][j], sub[k]);
    }
    if (rc != TSU_OK) break;
    // boundary rows on the side stream: after the previous update of the rows next to them and the arrival of the
    // neighbours' rows of the other colour
    if (t > 0) {
      cudaStreamWaitEvent(side, ev_sub[0][jp], 0);
      if (K > 1) cudaStreamWaitEvent(side, ev_sub[K - 1][jp], 0);
    }
    const uint32_t* top = d_up_halo ? halo(d_halo, opp, 0) : nullptr;
    const uint32_t* bot = d_down_halo ? halo(d_halo, opp, 1) : nullptr;
    if (top || bot)
      slab_wait_kernel<<<1, 1, 0, side>>>(top ? flag(d_flags, opp, 0) : nullptr, bot ? flag(d_flags, opp, 1) : nullptr,
                                          msgs[opp], d_flags + 8);
    rc = launch_half_sweep(d_state, n_replicas, rows, cols, 0, wrap_cols, colour, d_lut, d_lut_index, seed, sweep, replica0,
                           row0, top, bot, side, jit, 0, 1);
    if (rc == TSU_OK)
      rc = launch_half_sweep(d_state, n_replicas, rows, cols, 0, wrap_cols, colour, d_lut, d_lut_index, seed, sweep, replica0,
                             row0, top, bot, side, jit, rows - 1, rows);
    cudaEventRecord(ev_bnd[j], side);
    send(colour);
  }
  // the caller's stream continues after everything issued here
  cudaEventRecord(ev_join, side);
  cudaStreamWaitEvent(main, ev_join, 0);
  for (int k = 1; k < K; ++k) {
    cudaEventRecord(ev_join, sub[k]);
    cudaStreamWaitEvent(main, ev_join, 0);
  }
  for (int j = 0; j < 2; ++j) {
    cudaEventDestroy(ev_bnd[j]);
    for (int k = 0; k < K; ++k) cudaEventDestroy(ev_sub[k][j]);
  }
  cudaEventDestroy(ev_join);
  if (rc != TSU_OK) return rc;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TSU_OK : (int)e;
}

int tsu_ising2d_sweeps(uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_rows, int wrap_cols,
                       const uint32_t* d_lut, const int32_t* d_lut_index, uint64_t seed, uint32_t sweep0, int n_sweeps,
                       uint32_t replica0, uintptr_t stream) {
  TSU_CHECK_ARG(d_state && d_lut && geom_ok(n_replicas, rows, cols) && n_sweeps >= 0);
  TSU_CHECK_ARG(!wrap_cols || (cols % 2 == 0 && cols > 2));
  TSU_CHECK_ARG(!wrap_rows || (rows % 2 == 0 && rows > 2));
  if (n_sweeps > 0 && resident_eligible(n_replicas, rows, cols))
    return launch_resident_sweeps(d_state, n_replicas, rows, cols, wrap_rows, wrap_cols, d_lut, d_lut_index, seed, sweep0,
                                  n_sweeps, replica0, tsu_stream(stream));
  return launch_sweeps(d_state, n_replicas, rows, cols, wrap_rows, wrap_cols, d_lut, d_lut_index, seed, sweep0, n_sweeps,
                       replica0, tsu_stream(stream), JitKernel{nullptr, 0, 0});
}

int tsu_ising2d_half_sweep_injected(uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_rows,
                                    int wrap_cols, int colour, const uint32_t* d_lut, const int32_t* d_lut_index,
                                    const uint32_t* d_uniforms, int row0, const uint32_t* d_halo_top,
                                    const uint32_t* d_halo_bot, uintptr_t stream) {
  TSU_CHECK_ARG(d_state && d_lut && d_uniforms && geom_ok(n_replicas, rows, cols) && rows_ok(rows, row0));
  TSU_CHECK_ARG(colour == 0 || colour == 1);
  TSU_CHECK_ARG(!wrap_cols || (cols % 2 == 0 && cols > 2));
  TSU_CHECK_ARG(!wrap_rows || (rows % 2 == 0 && rows > 2) || d_halo_top || d_halo_bot);
  SweepParams P;
  P.state = d_state;
  P.lut = d_lut;
  P.lut_index = d_lut_index;
  P.halo_top = d_halo_top;
  P.halo_bot = d_halo_bot;
  P.g = make_geom(n_replicas, rows, cols, wrap_rows, wrap_cols, row0);
  P.colour = colour;
  P.sweep = 0;
  P.replica0 = 0;
  P.k0 = P.k1 = 0;
  P.strip_rows = 1;
  P.n_strips = rows;
  const long long total = (long long)n_replicas * rows * P.g.wpr;
  half_sweep_injected_kernel<<<blocks_for(total, 128), 128, 0, tsu_stream(stream)>>>(P, d_uniforms);
  TSU_RETURN_LAUNCH_STATUS();
}

int tsu_ising2d_observables(const uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_rows, int wrap_cols,
                            int row0, const uint32_t* d_next_rows, unsigned long long* d_out, uintptr_t stream) {
  TSU_CHECK_ARG(d_state && d_out && geom_ok(n_replicas, rows, cols) && rows_ok(rows, row0));
  Geom g = make_geom(n_replicas, rows, cols, wrap_rows, wrap_cols, row0);
  cudaStream_t st = tsu_stream(stream);
  cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(unsigned long long) * 2 * (size_t)n_replicas, st);
  if (e != cudaSuccess) return (int)e;
  if (cols % 256 == 0 && !tuning().obs_generic) {
    const int nvec = g.wpr / 4;
    int strip = 64;  // long strips re-read one row in `strip`; short ones fill the GPU for small batches
    while (strip > 1 && (long long)n_replicas * nvec * ((rows + strip - 1) / strip) < 148LL * 2048) strip >>= 1;
    const int n_strips = (rows + strip - 1) / strip;
    const long long per_rep_pad = ((long long)n_strips * nvec + 31) / 32 * 32;
    observables_fast_kernel<<<blocks_for((long long)n_replicas * per_rep_pad, 128), 128, 0, st>>>(d_state, g, d_next_rows,
                                                                                                strip, n_strips, d_out);
    TSU_RETURN_LAUNCH_STATUS();
  }
  TSU_CHECK_ARG(n_replicas <= 65535);  // gridDim.y of the word-by-word kernel
  const long long n_words = 2LL * rows * g.wpr;
  long long bx = (n_words + 255) / 256;
  const long long cap = (148LL * 8 + n_replicas - 1) / n_replicas;  // about 8 CTAs per SM in total
  if (bx > cap) bx = cap < 1 ? 1 : cap;
  dim3 grid((unsigned)bx, (unsigned)n_replicas);
  observables_kernel<<<grid, 256, 0, st>>>(d_state, g, d_next_rows, d_out);
  TSU_RETURN_LAUNCH_STATUS();
}

int tsu_ising2d_energy_from_observables(const unsigned long long* d_obs, int n_replicas, double J, double h,
                                        int64_t n_bonds, int64_t n_sites, double* d_energy, uintptr_t stream) {
  TSU_CHECK_ARG(d_obs && d_energy && n_replicas > 0);
  energy_from_obs_kernel<<<blocks_for(n_replicas, 128), 128, 0, tsu_stream(stream)>>>(d_obs, n_replicas, J, h, n_bonds,
                                                                                    n_sites, d_energy);
  TSU_RETURN_LAUNCH_STATUS();
}


int tsu_ising2d_jit_prepare(const uint32_t* h_lut, const char* src_dir, char* log_buf, int log_len) {
  TSU_CHECK_ARG(h_lut && src_dir);
  if (log_buf && log_len > 0) log_buf[0] = 0;
  std::lock_guard<std::mutex> lock(g_jit_mutex);
  if (!tsu_jit::available()) {
    if (log_buf && log_len > 0) snprintf(log_buf, log_len, "libnvrtc.so / libcuda.so.1 not available");
    return 0;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  unsigned tab[8];
  for (int k = 0; k < 8; ++k) {
    unsigned m = 0;
    for (int u = 0; u < 5; ++u) m |= ((h_lut[20 + u] >> (31 - k)) & 1u) << u;
    tab[k] = m;
  }
  unsigned fz = 0;  // classes whose threshold has zero low 24 bits and is not "always accept"
  for (int u = 0; u < 5; ++u)
    if ((h_lut[20 + u] & 0x00ffffffu) == 0u && !((h_lut[25] >> (20 + u)) & 1u)) fz |= 1u << u;
  char key[160];
  snprintf(key, sizeof key, "%d:%u,%u,%u,%u,%u,%u,%u,%u;%u;%u", dev, tab[0], tab[1], tab[2], tab[3], tab[4], tab[5], tab[6],
           tab[7], fz, (h_lut[25] >> 20) & 31u);
  auto it = g_jit_cache.find(key);
  if (it != g_jit_cache.end()) return it->second;
  std::string src;
  for (int k = 0; k < 8; ++k) src += "#define TSU_FT" + std::to_string(k) + " " + std::to_string(tab[k]) + "\n";
  src += "#define TSU_FZ " + std::to_string(fz) + "\n";
  const int jw = tuning().jit_w == 2 ? 2 : (tuning().jit_w == 4 ? 4 : kDefaultJitW);
  const int jt = (tuning().jit_threads == 32 || tuning().jit_threads == 64) ? tuning().jit_threads : kDefaultJitThreads;
  const int minb = tuning().jit_minb > 0 ? tuning().jit_minb : (jw == 2 ? 8 : kDefaultJitMinB) * (128 / jt);
  src += "#define TSU_ALWAYS " + std::to_string((h_lut[25] >> 20) & 31u) + "u\n";
  src += "#define TSU_JIT_MINB " + std::to_string(minb) + "\n";
  src += "#define TSU_JIT_W " + std::to_string(jw) + "\n";
  src += "#define TSU_JIT_THREADS " + std::to_string(jt) + "\n";
  src += "#define TSU_FAST_WARPS " + std::to_string(jt / 32) + "\n";
  if (tuning().jit_unroll > 1) src += "#define TSU_ROW_UNROLL " + std::to_string(tuning().jit_unroll) + "\n";
  src +=
      "#include \"ising2d_fast.cuh\"\n"
      "extern \"C\" __global__ void __launch_bounds__(TSU_JIT_THREADS, TSU_JIT_MINB) tsu_jit_half_sweep(tsu_fast::SweepParams P) {\n"
      "  tsu_fast::half_sweep_fast_body<TSU_JIT_W>(P);\n}\n";
  std::string log;
  void* fn = tsu_jit::compile(src, "tsu_jit.cu", "tsu_jit_half_sweep", src_dir, log);
  if (!fn) {
    if (log_buf && log_len > 0) snprintf(log_buf, log_len, "%s", log.c_str());
    g_jit_cache[key] = 0;
    return 0;
  }
  g_jit_functions.push_back(JitKernel{fn, jw, jt});
  const int handle = (int)g_jit_functions.size();
  g_jit_cache[key] = handle;
  return handle;
}

int tsu_ising2d_half_sweep_jit(int jit_handle, uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_rows,
                               int wrap_cols, int colour, const uint32_t* d_lut, uint64_t seed, uint32_t sweep,
                               uint32_t replica0, int row0, const uint32_t* d_halo_top, const uint32_t* d_halo_bot,
                               uintptr_t stream) {
  TSU_CHECK_ARG(d_state && d_lut && geom_ok(n_replicas, rows, cols) && rows_ok(rows, row0));
  TSU_CHECK_ARG(colour == 0 || colour == 1);
  TSU_CHECK_ARG(!wrap_cols || (cols % 2 == 0 && cols > 2));
  TSU_CHECK_ARG(!wrap_rows || (rows % 2 == 0 && rows > 2) || d_halo_top || d_halo_bot);
  return launch_half_sweep(d_state, n_replicas, rows, cols, wrap_rows, wrap_cols, colour, d_lut, nullptr, seed, sweep,
                           replica0, row0, d_halo_top, d_halo_bot, tsu_stream(stream), jit_function(jit_handle));
}

int tsu_ising2d_sweeps_jit(int jit_handle, uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_rows,
                           int wrap_cols, const uint32_t* d_lut, uint64_t seed, uint32_t sweep0, int n_sweeps,
                           uint32_t replica0, uintptr_t stream) {
  TSU_CHECK_ARG(d_state && d_lut && geom_ok(n_replicas, rows, cols) && n_sweeps >= 0);
  TSU_CHECK_ARG(!wrap_cols || (cols % 2 == 0 && cols > 2));
  TSU_CHECK_ARG(!wrap_rows || (rows % 2 == 0 && rows > 2));
  if (n_sweeps > 0 && resident_eligible(n_replicas, rows, cols))  // launch latency dominates: one launch for everything
    return launch_resident_sweeps(d_state, n_replicas, rows, cols, wrap_rows, wrap_cols, d_lut, nullptr, seed, sweep0,
                                  n_sweeps, replica0, tsu_stream(stream));
  return launch_sweeps(d_state, n_replicas, rows, cols, wrap_rows, wrap_cols, d_lut, nullptr, seed, sweep0, n_sweeps, replica0,
                       tsu_stream(stream), jit_function(jit_handle));
}

}  // extern "C"

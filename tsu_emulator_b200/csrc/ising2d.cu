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

#include <cstdlib>

#include "common.cuh"
#include "philox.cuh"

namespace {

struct Geom {
  int rows, cols, wpr;
  int wrap_rows, wrap_cols;
  int row0;
  int n_replicas;
};

__host__ __device__ __forceinline__ int colour_count(int cols, int p) { return (cols - p + 1) >> 1; }

__host__ __device__ __forceinline__ int words_per_row(int cols) {
  int ck = (cols + 1) / 2;
  int w = (ck + 31) / 32;
  return (w + 3) / 4 * 4;
}

__device__ __forceinline__ uint32_t lane_mask_lt(int n) {  // lanes [0, n), n clamped to [0, 32]
  return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u));
}

// pointers to the two colour planes of one replica
struct Planes {
  const uint32_t* opp;       // plane of the colour NOT being updated: [rows][wpr]
  const uint32_t* halo_top;  // opposite-colour row above local row 0, or nullptr
  const uint32_t* halo_bot;  // opposite-colour row below local row rows-1, or nullptr
};

__device__ __forceinline__ const uint32_t* opp_row(const Planes& P, const Geom& g, int i) {
  if (i < 0) return P.halo_top ? P.halo_top : (g.wrap_rows ? P.opp + (size_t)(g.rows - 1) * g.wpr : nullptr);
  if (i >= g.rows) return P.halo_bot ? P.halo_bot : (g.wrap_rows ? P.opp : nullptr);
  return P.opp + (size_t)i * g.wpr;
}

// Neighbourhood of word w of (colour, local row i) with all boundary cases.
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

struct Coords {
  uint32_t c0_base;  // w | colour << 20   (kind added per call)
  uint32_t row_g, sweep, replica, k0, k1;
};

__device__ __forceinline__ tsu_u32x4 lattice_call(const Coords& q, uint32_t kind) {
  return tsu_philox4x32_10(q.c0_base | (kind << 21), q.row_g, q.sweep, q.replica, q.k0, q.k1);
}

// full 32-bit uniform of lane j (top 8 bits from the planes, low 24 bits from the lane-group call)
__device__ __forceinline__ uint32_t lane_uniform(const uint32_t r[8], const Coords& q, int j) {
  uint32_t u = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) u |= ((r[k] >> j) & 1u) << (31 - k);
  tsu_u32x4 lo = lattice_call(q, TSU_KIND_LOW0 + (uint32_t)(j >> 2));
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
                                                uint32_t missW, uint32_t missE, const Coords& q) {
  // bit-sliced count of up neighbours: c2 c1 c0
  const uint32_t s1 = a ^ b ^ c;
  const uint32_t m1 = tsu_lop3_maj(a, b, c);
  const uint32_t c0 = s1 ^ s;
  const uint32_t k2 = s1 & s;
  const uint32_t c1 = m1 ^ k2;
  const uint32_t c2 = m1 & k2;

  uint32_t r[8];
  {
    tsu_u32x4 p0 = lattice_call(q, TSU_KIND_PLANE0);
    tsu_u32x4 p1 = lattice_call(q, TSU_KIND_PLANE1);
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
    const uint32_t bit = lane_accept(lut, d_row, up, lane_uniform(r, q, j));
    lt = (lt & ~(1u << j)) | (bit << j);
  }
  if (!FAST) {
    while (special) {
      const int j = __ffs(special) - 1;
      special &= special - 1u;
      const int up = ((c0 >> j) & 1u) | (((c1 >> j) & 1u) << 1) | (((c2 >> j) & 1u) << 2);
      const int d = d_row - (int)((missW >> j) & 1u) - (int)((missE >> j) & 1u);
      const uint32_t bit = lane_accept(lut, d, up, lane_uniform(r, q, j));
      lt = (lt & ~(1u << j)) | (bit << j);
    }
    lt &= valid;
  }
  return lt;
}

struct SweepParams {
  uint32_t* state;
  const uint32_t* lut;
  const int32_t* lut_index;
  const uint32_t* halo_top;
  const uint32_t* halo_bot;
  Geom g;
  int colour;
  uint32_t sweep, replica0, k0, k1;
  int strip_rows;  // rows per thread strip (fast path)
  int n_strips;    // strips per replica (fast path)
  int debug_flags; // experiments only (TSU_LATTICE_DEBUG): bit0 = skip tie resolution (WRONG results)
};

// ---- FAST path ---------------------------------------------------------------------------
// lop3 with a compile-time truth table: out bit = (LUT >> (4a + 2b + c)) & 1
template <int LUT>
__device__ __forceinline__ uint32_t lop3_imm(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
  return d;
}

// Threshold bit of every lane for one bit-plane: tk = table[up-count], up-count = 4 c2 + 2 c1 + c0.
// The 5-entry truth table T (bit u = threshold bit of class u) is warp-uniform in practice, so instead
// of four bitwise selects per word (ALU pipe) an indexed branch (brx.idx -> SASS BRX) picks the lop3
// immediate: one LOP3 per word plus one jump per plane shared by the thread's four words.
#define TSU_TK_CASE(I)                                                                    \
  "L" #I ": lop3.b32 %0, %4, %8, %12, " #I "; lop3.b32 %1, %5, %9, %13, " #I ";"          \
  " lop3.b32 %2, %6, %10, %14, " #I "; lop3.b32 %3, %7, %11, %15, " #I "; bra.uni LDONE;\n"

__device__ __forceinline__ void tk_select4(uint32_t T, const uint32_t c2[4], const uint32_t c1[4],
                                           const uint32_t c0[4], uint32_t tk[4]) {
  asm("{\n"
      "LTAB: .branchtargets L0, L1, L2, L3, L4, L5, L6, L7, L8, L9, L10, L11, L12, L13, L14, L15, L16, L17, "
      "L18, L19, L20, L21, L22, L23, L24, L25, L26, L27, L28, L29, L30, L31;\n"
      "brx.idx %16, LTAB;\n"
      TSU_TK_CASE(0) TSU_TK_CASE(1) TSU_TK_CASE(2) TSU_TK_CASE(3) TSU_TK_CASE(4) TSU_TK_CASE(5) TSU_TK_CASE(6)
      TSU_TK_CASE(7) TSU_TK_CASE(8) TSU_TK_CASE(9) TSU_TK_CASE(10) TSU_TK_CASE(11) TSU_TK_CASE(12)
      TSU_TK_CASE(13) TSU_TK_CASE(14) TSU_TK_CASE(15) TSU_TK_CASE(16) TSU_TK_CASE(17) TSU_TK_CASE(18)
      TSU_TK_CASE(19) TSU_TK_CASE(20) TSU_TK_CASE(21) TSU_TK_CASE(22) TSU_TK_CASE(23) TSU_TK_CASE(24)
      TSU_TK_CASE(25) TSU_TK_CASE(26) TSU_TK_CASE(27) TSU_TK_CASE(28) TSU_TK_CASE(29) TSU_TK_CASE(30)
      TSU_TK_CASE(31)
      "LDONE:\n"
      "}\n"
      : "=r"(tk[0]), "=r"(tk[1]), "=r"(tk[2]), "=r"(tk[3])
      : "r"(c2[0]), "r"(c2[1]), "r"(c2[2]), "r"(c2[3]), "r"(c1[0]), "r"(c1[1]), "r"(c1[2]), "r"(c1[3]),
        "r"(c0[0]), "r"(c0[1]), "r"(c0[2]), "r"(c0[3]), "r"(T & 31u));
}

// Warp-cooperative, software-pipelined resolution of "tie" lanes (top byte of the uniform equals the
// threshold's, 2^-8 per lane, ~12 % of the words have one).  Row i: owners store the row with tie lanes
// cleared and push one descriptor per word that has ties into a per-warp shared-memory queue
// (ballot-allocated slots, no atomics, no per-tie loop).  Row i+1: every lane of the warp takes one
// queued word of row i, draws the low 24 bits of its tie lanes (one Philox call each) and ORs the accepted
// lanes into the stored word with a global RED.  A warp row (128 words) has ~15 such words: one pass of
// useful work per lane instead of max-over-lanes(#ties) serial passes with one or two active lanes.
constexpr int kTieCap = 96;
struct TieQueue {
  uint32_t count;
  uint32_t pad[3];
  uint32_t word[kTieCap];  // word index in the row
  uint32_t row[kTieCap];   // local row
  uint32_t mask[kTieCap];  // tie lanes
  uint32_t c0[kTieCap], c1[kTieCap], c2[kTieCap];  // bit-sliced up-neighbour count of the word
};

// low 24 bits of the uniforms of the tie lanes `mask` of one word; returns the accepted lanes
__device__ __forceinline__ uint32_t resolve_word_ties(uint32_t mask, uint32_t c0, uint32_t c1, uint32_t c2,
                                                      uint32_t c0_word, uint32_t row_g,
                                                      const uint32_t* __restrict__ lut, const Coords& q) {
  uint32_t acc = 0u;
  while (mask) {
    const int j = __ffs(mask) - 1;
    mask &= mask - 1u;
    const uint32_t up = ((c0 >> j) & 1u) | (((c1 >> j) & 1u) << 1) | (((c2 >> j) & 1u) << 2);
    const tsu_u32x4 lo = tsu_philox4x32_10(c0_word | ((TSU_KIND_LOW0 + (uint32_t)(j >> 2)) << 21), row_g, q.sweep,
                                           q.replica, q.k0, q.k1);
    const int sel = j & 3;
    const uint32_t vv = sel == 0 ? lo.x : (sel == 1 ? lo.y : (sel == 2 ? lo.z : lo.w));
    if ((vv >> 8) < (__ldg(lut + 20 + up) & 0x00ffffffu)) acc |= 1u << j;
  }
  return acc;
}

__device__ __forceinline__ void drain_tie_queue(const TieQueue& tp, uint32_t lane, uint32_t* own, const Geom& g,
                                                uint32_t colour_bits, const uint32_t* __restrict__ lut,
                                                const Coords& q, int debug_flags) {
  if (debug_flags & 2) return;
  const uint32_t n_tie = tp.count;
  for (uint32_t t = lane; t < n_tie; t += 32u) {
    const uint32_t w = tp.word[t], row_l = tp.row[t];
    const uint32_t acc = resolve_word_ties(tp.mask[t], tp.c0[t], tp.c1[t], tp.c2[t], w | colour_bits,
                                           (uint32_t)g.row0 + row_l, lut, q);
    if (acc && !(debug_flags & 4)) atomicOr(own + (size_t)row_l * g.wpr + w, acc);
    if ((debug_flags & 4) && acc == 0xdeadbeefu) own[0] = acc;
  }
}

// Periodic columns, every word full (cols % 256 == 0), a neighbour row above and below every local row
// (wrap or halo).  One thread owns a 4-word (128 spin) column strip of `strip_rows` rows and keeps a
// rolling window (north, centre, south, next south) of 128-bit loads; the four words of a row are
// processed together.  Threads of a warp always belong to the same replica (thread index space padded
// to a multiple of 32 per replica), so the threshold tables are warp-uniform.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) half_sweep_fast_kernel(SweepParams P) {
  __shared__ TieQueue tie_queues[4][2];
  const Geom& g = P.g;
  const int nvec = g.wpr >> 2;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_rep = P.n_strips * nvec;
  const int per_rep_pad = (per_rep + 31) & ~31;
  const int rep = (int)(tid / per_rep_pad);
  if (rep >= g.n_replicas) return;  // whole warps only: per_rep_pad is a multiple of 32
  const int rem = (int)(tid - (long long)rep * per_rep_pad);
  const bool active = rem < per_rep;
  const int strip = active ? rem / nvec : 0;
  const int v = active ? rem - strip * nvec : 0;
  const int r_begin = strip * P.strip_rows;
  const int r_end = active ? min(g.rows, r_begin + P.strip_rows) : r_begin;
  const unsigned lane = threadIdx.x & 31u;
  TieQueue* tqs = tie_queues[threadIdx.x >> 5];
  if (lane == 0) {
    tqs[0].count = 0u;
    tqs[1].count = 0u;
  }
  __syncwarp();

  const size_t plane = (size_t)g.rows * g.wpr;
  uint32_t* own = P.state + ((size_t)rep * 2 + P.colour) * plane;
  Planes pl;
  pl.opp = P.state + ((size_t)rep * 2 + (1 - P.colour)) * plane;
  pl.halo_top = P.halo_top ? P.halo_top + (size_t)rep * g.wpr : nullptr;
  pl.halo_bot = P.halo_bot ? P.halo_bot + (size_t)rep * g.wpr : nullptr;
  const uint32_t* __restrict__ lut = P.lut + (P.lut_index ? (size_t)P.lut_index[rep] * 32 : 0);

  // per-plane truth tables of the degree-4 classes (5 bits each, planes 0-5 in Tlo, 6-7 in Thi)
  uint32_t Tlo = 0u, Thi = 0u;
  {
    uint32_t t[5];
#pragma unroll
    for (int u = 0; u < 5; ++u) t[u] = __ldg(lut + 20 + u);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint32_t m = 0u;
#pragma unroll
      for (int u = 0; u < 5; ++u) m |= ((t[u] >> (31 - k)) & 1u) << u;
      if (k < 6)
        Tlo |= m << (5 * k);
      else
        Thi |= m << (5 * (k - 6));
    }
  }
  const uint32_t always = (__ldg(lut + 25) >> 20) & 31u;

  Coords q;
  q.sweep = P.sweep;
  q.replica = P.replica0 + (uint32_t)rep;
  q.k0 = P.k0;
  q.k1 = P.k1;
  const uint32_t colour_bits = (uint32_t)P.colour << 20;

  const int w0 = v * 4;
  const int w_prev = (w0 == 0) ? g.wpr - 1 : w0 - 1;
  const int w_next = (w0 + 4 == g.wpr) ? 0 : w0 + 4;

  uint4 n = make_uint4(0, 0, 0, 0), c = n, s = n;
  uint32_t side_c = 0u;  // neighbour word of the centre row needed by the funnel shift
  if (active) {
    n = *reinterpret_cast<const uint4*>(opp_row(pl, g, r_begin - 1) + w0);
    const uint32_t* rc0 = opp_row(pl, g, r_begin);
    c = *reinterpret_cast<const uint4*>(rc0 + w0);
    side_c = rc0[((g.row0 + r_begin + P.colour) & 1) ? w_next : w_prev];
    s = *reinterpret_cast<const uint4*>(opp_row(pl, g, r_begin + 1) + w0);
  }
  for (int it = 0; it < P.strip_rows; ++it) {
    const int i = r_begin + it;
    const bool row_valid = i < r_end;
    const int row_g = g.row0 + i;
    const int p = (row_g + P.colour) & 1;
    // resolve the words with ties queued by the previous row: one queued word per lane
    // (two rows are batched per queue so that ~30 of the 32 lanes have a word to resolve)
    if (it > 0 && (it & 1) == 0) drain_tie_queue(tqs[((it >> 1) & 1) ^ 1], lane, own, g, colour_bits, lut, q, P.debug_flags);
    // prefetch the row after next (and the side word of the next row) while this row is computed
    uint4 s2 = s;
    uint32_t side_s = 0u;
    if (i + 1 < r_end) {
      const uint32_t* rs = opp_row(pl, g, i + 1);
      side_s = rs[p ? w_prev : w_next];  // the next row has the opposite parity
      s2 = *reinterpret_cast<const uint4*>(opp_row(pl, g, i + 2) + w0);
    }
    uint32_t lt[4] = {0u, 0u, 0u, 0u};
    uint32_t eq[4] = {0u, 0u, 0u, 0u};
    uint32_t c0[4], c1[4], c2[4];
    if (row_valid) {
      uint32_t sd[4];
      if (p) {
        sd[0] = __funnelshift_r(c.x, c.y, 1);
        sd[1] = __funnelshift_r(c.y, c.z, 1);
        sd[2] = __funnelshift_r(c.z, c.w, 1);
        sd[3] = __funnelshift_r(c.w, side_c, 1);
      } else {
        sd[0] = __funnelshift_l(side_c, c.x, 1);
        sd[1] = __funnelshift_l(c.x, c.y, 1);
        sd[2] = __funnelshift_l(c.y, c.z, 1);
        sd[3] = __funnelshift_l(c.z, c.w, 1);
      }
      q.row_g = (uint32_t)row_g;
      // bit-sliced up-neighbour count of the four words
      const uint32_t an[4] = {n.x, n.y, n.z, n.w}, as[4] = {s.x, s.y, s.z, s.w}, ac[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t s1 = an[k] ^ as[k] ^ ac[k];
        const uint32_t m1 = tsu_lop3_maj(an[k], as[k], ac[k]);
        c0[k] = s1 ^ sd[k];
        const uint32_t k2 = s1 & sd[k];
        c1[k] = m1 ^ k2;
        c2[k] = m1 & k2;
      }
      // borrow-chain compare of the top 8 bits of the uniforms against the thresholds, least significant
      // plane first; planes 4-7 (second Philox call) are consumed before planes 0-3 are generated
#pragma unroll
      for (int k = 0; k < 4; ++k) eq[k] = 0xffffffffu;
#pragma unroll
      for (int half = 1; half >= 0; --half) {
        uint32_t r[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          q.c0_base = (uint32_t)(w0 + k) | colour_bits;
          const tsu_u32x4 pp = lattice_call(q, half ? TSU_KIND_PLANE1 : TSU_KIND_PLANE0);
          r[k][0] = pp.x; r[k][1] = pp.y; r[k][2] = pp.z; r[k][3] = pp.w;
        }
#pragma unroll
        for (int kk = 3; kk >= 0; --kk) {
          const int k = half * 4 + kk;
          const uint32_t T = k < 6 ? (Tlo >> (5 * k)) : (Thi >> (5 * (k - 6)));
          uint32_t tk[4];
          tk_select4(T, c2, c1, c0, tk);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t x = r[j][kk] ^ tk[j];
            lt[j] = (~r[j][kk] & tk[j]) | (~x & lt[j]);
            eq[j] &= ~x;
          }
        }
      }
      if (always) {  // classes with p == 1.0 (threshold 2^32 does not fit 32 bits)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t am = 0u;
          if (always & 1u) am |= ~c2[j] & ~c1[j] & ~c0[j];
          if (always & 2u) am |= ~c2[j] & ~c1[j] & c0[j];
          if (always & 4u) am |= ~c2[j] & c1[j] & ~c0[j];
          if (always & 8u) am |= ~c2[j] & c1[j] & c0[j];
          if (always & 16u) am |= c2[j];
          lt[j] |= am;
          eq[j] &= ~am;
        }
      }
    }
    // ---- tie lanes of this row: store the row without them, queue them for the next iteration ----
    if (row_valid) {
      uint4 o;
      o.x = lt[0]; o.y = lt[1]; o.z = lt[2]; o.w = lt[3];
      *reinterpret_cast<uint4*>(own + (size_t)i * g.wpr + w0) = o;
    }
    {
      TieQueue& tq = tqs[(it >> 1) & 1];
      uint32_t base = (it & 1) ? tq.count : 0u;  // second row of the pair appends
      if (!(P.debug_flags & 1)) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool has = eq[k] != 0u;
          const unsigned m = __ballot_sync(0xffffffffu, has);
          if (has) {
            const uint32_t slot = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
            if (slot < (uint32_t)kTieCap) {
              tq.word[slot] = (uint32_t)(w0 + k);
              tq.row[slot] = (uint32_t)i;
              tq.mask[slot] = eq[k];
              tq.c0[slot] = c0[k];
              tq.c1[slot] = c1[k];
              tq.c2[slot] = c2[k];
            } else {  // queue overflow (practically never): resolve on the spot
              const uint32_t acc = resolve_word_ties(eq[k], c0[k], c1[k], c2[k], (uint32_t)(w0 + k) | colour_bits,
                                                     (uint32_t)row_g, lut, q);
              if (acc) atomicOr(own + (size_t)i * g.wpr + (w0 + k), acc);
            }
          }
          base += (uint32_t)__popc(m);
        }
      }
      if (lane == 0) tq.count = min(base, (uint32_t)kTieCap);
    }
    __syncwarp();  // queue of this row and the stored words are visible to the warp
    n = c;
    c = s;
    s = s2;
    side_c = side_s;
  }
  drain_tie_queue(tqs[((P.strip_rows - 1) >> 1) & 1], lane, own, g, colour_bits, lut, q, P.debug_flags);  // last pair of rows
}

// Generic path: any size, open or periodic edges, ragged last word.  One thread per word.
__global__ void __launch_bounds__(128) half_sweep_generic_kernel(SweepParams P) {
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
  if (h.valid == 0u) return;  // padding word (stays 0)
  const int d_row = 2 + h.has_n + h.has_s;
  LutRegs L;
  load_lut_regs(L, lut, d_row);
  Coords q;
  q.c0_base = (uint32_t)w | ((uint32_t)P.colour << 20);
  q.row_g = (uint32_t)(g.row0 + i);
  q.sweep = P.sweep;
  q.replica = P.replica0 + (uint32_t)rep;
  q.k0 = P.k0;
  q.k1 = P.k1;
  own[(size_t)i * g.wpr + w] = update_word<false>(h.n, h.s, h.c, h.side, L, lut, d_row, h.valid, h.missW, h.missE, q);
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
  tsu_u32x4 o = tsu_philox4x32_10(TSU_LATTICE_C0(w, colour, TSU_KIND_INIT), (uint32_t)(g.row0 + i), 0u,
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

__global__ void energy_from_obs_kernel(const unsigned long long* __restrict__ obs, int n, double J, double h,
                                       long long n_bonds, long long n_sites, double* energy) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const double up = (double)obs[2 * r];
  const double anti = (double)obs[2 * r + 1];
  energy[r] = -J * ((double)n_bonds - 2.0 * anti) - h * (2.0 * up - (double)n_sites);
}

bool geom_ok(int n_replicas, int rows, int cols) {
  return n_replicas > 0 && rows > 0 && cols > 0 && cols < (1 << 26) && rows < (1 << 30);
}

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

int launch_half_sweep(uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_rows, int wrap_cols, int colour,
                      const uint32_t* d_lut, const int32_t* d_lut_index, uint64_t seed, uint32_t sweep,
                      uint32_t replica0, int row0, const uint32_t* d_halo_top, const uint32_t* d_halo_bot,
                      cudaStream_t st) {
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
  P.debug_flags = 0;
  if (const char* e = getenv("TSU_LATTICE_DEBUG")) P.debug_flags = atoi(e);
  const bool rows_closed = (wrap_rows && !d_halo_top && !d_halo_bot) || (d_halo_top && d_halo_bot);
  const bool fast = wrap_cols && (cols % 256 == 0) && rows_closed;
  if (fast) {
    const int nvec = P.g.wpr / 4;
    // strips long enough to amortise the two halo rows, short enough to fill 148 SMs x 16 warps
    const long long target_threads = 148LL * 2048;
    int strip = 64;
    while (strip > 1 && (long long)n_replicas * nvec * ((rows + strip - 1) / strip) < target_threads) strip >>= 1;
    if (const char* e = getenv("TSU_LATTICE_STRIP")) strip = atoi(e) > 0 ? atoi(e) : strip;
    P.strip_rows = strip;
    P.n_strips = (rows + strip - 1) / strip;
    const long long per_rep_pad = ((long long)P.n_strips * nvec + 31) / 32 * 32;  // warps never straddle replicas
    const long long total = (long long)n_replicas * per_rep_pad;
    int minb = 4;
    if (const char* e = getenv("TSU_LATTICE_MINB")) minb = atoi(e);
    const unsigned grid = blocks_for(total, 128);
    if (minb == 3) half_sweep_fast_kernel<3><<<grid, 128, 0, st>>>(P);
    else if (minb == 5) half_sweep_fast_kernel<5><<<grid, 128, 0, st>>>(P);
    else if (minb == 6) half_sweep_fast_kernel<6><<<grid, 128, 0, st>>>(P);
    else half_sweep_fast_kernel<4><<<grid, 128, 0, st>>>(P);
  } else {
    P.strip_rows = 1;
    P.n_strips = rows;
    const long long total = (long long)n_replicas * rows * P.g.wpr;
    half_sweep_generic_kernel<<<blocks_for(total, 128), 128, 0, st>>>(P);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TSU_OK : (int)e;
}

}  // namespace

extern "C" {

int64_t tsu_ising2d_words_per_row(int cols) { return cols > 0 ? words_per_row(cols) : 0; }

int64_t tsu_ising2d_state_words(int rows, int cols) {
  return (rows > 0 && cols > 0) ? 2LL * rows * words_per_row(cols) : 0;
}

int tsu_ising2d_init_random(uint32_t* d_state, int n_replicas, int rows, int cols, uint64_t seed, uint32_t replica0,
                            int row0, uintptr_t stream) {
  TSU_CHECK_ARG(d_state && geom_ok(n_replicas, rows, cols) && row0 >= 0);
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
  TSU_CHECK_ARG(d_state && d_lut && geom_ok(n_replicas, rows, cols) && row0 >= 0);
  TSU_CHECK_ARG(colour == 0 || colour == 1);
  TSU_CHECK_ARG(!wrap_cols || (cols % 2 == 0 && cols > 2));
  TSU_CHECK_ARG(!wrap_rows || (rows % 2 == 0 && rows > 2) || d_halo_top || d_halo_bot);
  return launch_half_sweep(d_state, n_replicas, rows, cols, wrap_rows, wrap_cols, colour, d_lut, d_lut_index, seed,
                           sweep, replica0, row0, d_halo_top, d_halo_bot, tsu_stream(stream));
}

int tsu_ising2d_sweeps(uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_rows, int wrap_cols,
                       const uint32_t* d_lut, const int32_t* d_lut_index, uint64_t seed, uint32_t sweep0, int n_sweeps,
                       uint32_t replica0, uintptr_t stream) {
  TSU_CHECK_ARG(d_state && d_lut && geom_ok(n_replicas, rows, cols) && n_sweeps >= 0);
  TSU_CHECK_ARG(!wrap_cols || (cols % 2 == 0 && cols > 2));
  TSU_CHECK_ARG(!wrap_rows || (rows % 2 == 0 && rows > 2));
  for (int t = 0; t < n_sweeps; ++t) {
    for (int colour = 0; colour < 2; ++colour) {
      int rc = launch_half_sweep(d_state, n_replicas, rows, cols, wrap_rows, wrap_cols, colour, d_lut, d_lut_index,
                                 seed, sweep0 + (uint32_t)t, replica0, 0, nullptr, nullptr, tsu_stream(stream));
      if (rc != TSU_OK) return rc;
    }
  }
  return TSU_OK;
}

int tsu_ising2d_half_sweep_injected(uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_rows,
                                    int wrap_cols, int colour, const uint32_t* d_lut, const int32_t* d_lut_index,
                                    const uint32_t* d_uniforms, int row0, const uint32_t* d_halo_top,
                                    const uint32_t* d_halo_bot, uintptr_t stream) {
  TSU_CHECK_ARG(d_state && d_lut && d_uniforms && geom_ok(n_replicas, rows, cols) && row0 >= 0);
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
  P.debug_flags = 0;
  P.strip_rows = 1;
  P.n_strips = rows;
  const long long total = (long long)n_replicas * rows * P.g.wpr;
  half_sweep_injected_kernel<<<blocks_for(total, 128), 128, 0, tsu_stream(stream)>>>(P, d_uniforms);
  TSU_RETURN_LAUNCH_STATUS();
}

int tsu_ising2d_observables(const uint32_t* d_state, int n_replicas, int rows, int cols, int wrap_rows, int wrap_cols,
                            int row0, const uint32_t* d_next_rows, unsigned long long* d_out, uintptr_t stream) {
  TSU_CHECK_ARG(d_state && d_out && geom_ok(n_replicas, rows, cols) && row0 >= 0);
  TSU_CHECK_ARG(n_replicas <= 65535);
  Geom g = make_geom(n_replicas, rows, cols, wrap_rows, wrap_cols, row0);
  cudaStream_t st = tsu_stream(stream);
  cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(unsigned long long) * 2 * (size_t)n_replicas, st);
  if (e != cudaSuccess) return (int)e;
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

}  // extern "C"

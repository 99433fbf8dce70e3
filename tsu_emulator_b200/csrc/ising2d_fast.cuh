// Fast path of the bit-packed checkerboard heat-bath update (see ising2d.cu for the algorithm notes).
//
// This header is compiled twice: by nvcc as part of libtsu_b200.so (threshold truth tables read at run
// time, threshold-bit select through a brx.idx jump table) and by NVRTC at run time with the eight 5-bit
// truth tables of one temperature as compile-time constants (TSU_FT0..TSU_FT7, TSU_JIT_ALWAYS): the select
// becomes one lop3 with a literal immediate, the jump tables disappear and the loop body shrinks from 48 KB
// to 29 KB of SASS.  Keep it free of host headers.
#pragma once
#ifdef __CUDACC_RTC__
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef unsigned char uint8_t;
typedef int int32_t;
typedef unsigned long size_t;
#else
#include <stdint.h>
#include <stddef.h>
#endif
#include "philox.cuh"

namespace tsu_fast {

__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (a & c) | (b & c); }

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

struct Coords {
  uint32_t c0_base;  // w | colour << 20   (kind added per call)
  uint32_t row_g, sweep, replica, k0, k1;
};

__device__ __forceinline__ tsu_u32x4 lattice_call(const Coords& q, uint32_t kind) {
  return tsu_philox4x32_10(q.c0_base | (kind << 21), q.row_g, q.sweep, q.replica, q.k0, q.k1);
}

// full 32-bit uniform of lane j (top 8 bits from the planes, low 24 bits from the lane-group call)

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
  int row_begin, row_end;  // local rows the fast path updates (rows without a north / south neighbour are left to the rim pass)
  int nvec_fast;           // 4-word groups per row the fast path updates (all their lanes exist); the rest is rim
  int debug_flags; // experiments only (TSU_LATTICE_DEBUG): bit0 = skip tie resolution (WRONG results)
};


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
__device__ __forceinline__ void half_sweep_fast_body(const SweepParams& P) {
  __shared__ TieQueue tie_queues[4][2];
  const Geom& g = P.g;
  const int nvec = P.nvec_fast;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_rep = P.n_strips * nvec;
  const int per_rep_pad = (per_rep + 31) & ~31;
  const int rep = (int)(tid / per_rep_pad);
  if (rep >= g.n_replicas) return;  // whole warps only: per_rep_pad is a multiple of 32
  const int rem = (int)(tid - (long long)rep * per_rep_pad);
  const bool active = rem < per_rep;
  const int strip = active ? rem / nvec : 0;
  const int v = active ? rem - strip * nvec : 0;
  const int r_begin = P.row_begin + strip * P.strip_rows;
  const int r_end = active ? min(P.row_end, r_begin + P.strip_rows) : r_begin;
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
        const uint32_t m1 = lop3_maj(an[k], as[k], ac[k]);
        c0[k] = s1 ^ sd[k];
        const uint32_t k2 = s1 & sd[k];
        c1[k] = m1 ^ k2;
        c2[k] = m1 & k2;
      }
      // borrow-chain compare of the top 8 bits of the uniforms against the thresholds, least significant
      // plane first; planes 4-7 (second Philox call) are consumed before planes 0-3 are generated
#pragma unroll
      for (int k = 0; k < 4; ++k) eq[k] = 0xffffffffu;
#ifdef TSU_JIT_WIDE
      // all eight Philox calls of the row are issued before any plane is consumed (more independent chains)
      uint32_t rw[2][4][4];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          q.c0_base = (uint32_t)(w0 + k) | colour_bits;
          const tsu_u32x4 pp = lattice_call(q, hh ? TSU_KIND_PLANE1 : TSU_KIND_PLANE0);
          rw[hh][k][0] = pp.x; rw[hh][k][1] = pp.y; rw[hh][k][2] = pp.z; rw[hh][k][3] = pp.w;
        }
#endif
#pragma unroll
      for (int half = 1; half >= 0; --half) {
        uint32_t r[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#ifdef TSU_JIT_WIDE
          r[k][0] = rw[half][k][0]; r[k][1] = rw[half][k][1]; r[k][2] = rw[half][k][2]; r[k][3] = rw[half][k][3];
#else
          q.c0_base = (uint32_t)(w0 + k) | colour_bits;
          const tsu_u32x4 pp = lattice_call(q, half ? TSU_KIND_PLANE1 : TSU_KIND_PLANE0);
          r[k][0] = pp.x; r[k][1] = pp.y; r[k][2] = pp.z; r[k][3] = pp.w;
#endif
        }
#pragma unroll
        for (int kk = 3; kk >= 0; --kk) {
          const int k = half * 4 + kk;
          const uint32_t T = k < 6 ? (Tlo >> (5 * k)) : (Thi >> (5 * (k - 6)));
          uint32_t tk[4];
#ifdef TSU_FT0
          {
            constexpr int kTab[8] = {TSU_FT0, TSU_FT1, TSU_FT2, TSU_FT3, TSU_FT4, TSU_FT5, TSU_FT6, TSU_FT7};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              switch (k) {  // k is a compile-time constant after unrolling
                case 0: tk[j] = lop3_imm<kTab[0]>(c2[j], c1[j], c0[j]); break;
                case 1: tk[j] = lop3_imm<kTab[1]>(c2[j], c1[j], c0[j]); break;
                case 2: tk[j] = lop3_imm<kTab[2]>(c2[j], c1[j], c0[j]); break;
                case 3: tk[j] = lop3_imm<kTab[3]>(c2[j], c1[j], c0[j]); break;
                case 4: tk[j] = lop3_imm<kTab[4]>(c2[j], c1[j], c0[j]); break;
                case 5: tk[j] = lop3_imm<kTab[5]>(c2[j], c1[j], c0[j]); break;
                case 6: tk[j] = lop3_imm<kTab[6]>(c2[j], c1[j], c0[j]); break;
                default: tk[j] = lop3_imm<kTab[7]>(c2[j], c1[j], c0[j]); break;
              }
            }
            (void)T;
          }
#else
          tk_select4(T, c2, c1, c0, tk);
#endif
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
#ifdef TSU_FZ
    // classes whose threshold has zero low 24 bits (e.g. p = 1/2 exactly): a tie can never be accepted by the
    // low bits, so those lanes need no second draw
    if (TSU_FZ != 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) eq[j] &= ~lop3_imm<TSU_FZ>(c2[j], c1[j], c0[j]);
    }
#endif
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
}  // namespace tsu_fast

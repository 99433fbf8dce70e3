// Dense-coupling Gibbs sampler on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces, for a batch of chains that share one coupling matrix (BASELINE config 3: N = 4096 spins,
// 2048 chains), the local-field evaluation of the reference
//     h_i = np.dot(coupling[i, :], state) + bias[i]                 tsu/gibbs.py:79-100
// inside the sequential sweep of tsu/gibbs.py:128-162.
//
// Exact sequential Gibbs, blocked: the N sites are visited in index order in blocks of 64.  For a block
// the fields of its 64 sites for 128 chains are one 128 x 64 x N GEMM  H = S . J[blk, :]^T  (S: current bits
// as bf16 0/1, J: bf16, fp32 accumulation in TMEM) issued as tcgen05.mma instructions by one thread; the
// epilogue thread of each chain then walks the 64 sites in order, draws the heat-bath bit from
// sigmoid(h/T) and applies the rank-1 correction h_i' += J[i', i] * (new - old) to the not yet visited
// sites of the block, which makes the result identical to a site-by-site sweep with the same fields.
//
// One CTA owns 128 chains (TMEM lane = chain).  The chain states stay resident in shared memory as bits
// for the whole sweep (64 KB).  Each K-chunk of 128 sites is expanded to bf16 by the thread that owns the
// chain and written straight into TENSOR MEMORY (tcgen05.st, lane = chain, column = K pair): the spin
// operand A never touches shared memory.  A spin is encoded as 0.0 / 2.0: bf16 2.0 = 0x4000 has ONE set bit,
// so a packed pair of spins is (word << s) & 0x40004000 - two integer instructions per register - and the
// accumulated field is halved (exactly) in the epilogue.  The matching J tile (operand B, K-major, no
// swizzle) is streamed from L2 with cp.async.
//
// The K-chunks of a block alternate between two independent pipelines (producer group of 4 warps + one MMA
// issuer warp + own operand rings + own accumulator), so that the per-chunk synchronisation latencies
// (mbarrier wake-ups, TMEM store drain, commit) of one pipeline hide behind the other.

#include <cstdlib>
#include <cuda_bf16.h>

#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr int kChains = 128;  // chains per CTA = UMMA M = TMEM lanes
constexpr int kBlk = 32;      // sites per block = UMMA N
constexpr int kKC = 128;      // K-chunk (sites) per pipeline stage
constexpr int kPipes = 2;     // independent producer/issuer pipelines; chunk cc of a block goes to pipeline cc % kPipes
constexpr int kASlots = 3;    // per pipeline: expanded spin tiles in tensor memory (64 columns each)
constexpr int kLook = 4;      // per pipeline: J tiles are requested 4 of its own chunks ahead
constexpr int kBSlots = kLook + kASlots;  // per pipeline: J tile ring (8 KB each).  With this depth the J slot of chunk
                                          // m + kLook is the one chunk m - kASlots used: ONE "empty" barrier frees both
constexpr int kAccCols = 2 * kPipes * kBlk;  // two buffers x one 32-column fp32 accumulator per pipeline = 128 columns
constexpr int kACols = kKC / 2;              // 32-bit columns of one expanded spin tile
static_assert(kAccCols + kPipes * kASlots * kACols <= 512, "tensor memory budget");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 bytes (contiguous 128 B);
// SBO = byte distance between 8-row groups, LBO = byte distance between the two 8-element K halves
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base offset 0, layout type 0 = SWIZZLE_NONE
}

// instruction descriptor: D = F32, A = B = BF16, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// same with the A operand in tensor memory (lane = row, one 32-bit column = two consecutive K elements)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// for long waits (the epilogue idles for a whole block's GEMM): back off so that the polling does not compete
// with the producers for shared-memory / issue bandwidth
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(256);
  }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

#ifdef TSU_TC_TIMING  // per-role stall accounting, see tools/tc_timing.py
__device__ unsigned long long g_tc_timing[32];
#define TC_T0() long long t0__ = clock64()
#define TC_ACC(i) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) atomicAdd(&g_tc_timing[i], (unsigned long long)(clock64() - t0__)); } while (0)
#define TC_NEXT(i) do { const long long t1__ = clock64(); if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) atomicAdd(&g_tc_timing[i], (unsigned long long)(t1__ - t0__)); t0__ = t1__; } while (0)
#else
#define TC_NEXT(i) do {} while (0)
#define TC_T0() do {} while (0)
#define TC_ACC(i) do {} while (0)
#endif

struct TcParams {
  const __nv_bfloat16* J;   // [N][N] row-major coupling matrix (row i = couplings INTO site i)
  const float* bias;        // [N] or nullptr
  uint8_t* state;           // [n_chains][N] bits, updated in place
  float* fields_out;        // diagnostics: [n_chains][N] field of every site at the time it was visited
  const double* T_chain;    // [n_chains] or nullptr
  int n_chains, N, n_sweeps;
  float T;
  uint32_t k0, k1, sweep0, chain0;
  int gemm_only;            // diagnostics: no spin update (fields of the initial state for every site)
  int dbg;                  // diagnostics (TSU_TC_DEBUG, timing experiments only): 1 no MMA, 2 no A expansion, 4 no J loads
};

constexpr int kProducers = 128 * kPipes;            // warps 0-7: 4 warps (128 chains) per pipeline
constexpr int kEpilogue0 = kProducers / 32;         // warps 8-11: epilogue (thread = chain = TMEM lane)
constexpr int kIssuer0 = kEpilogue0 + 4;            // warps 12-13: MMA issuers
constexpr int kThreads = 32 * (kIssuer0 + kPipes);  // 448

// position of site i (0-31 of a block) inside a state word: even sites in the low half, odd sites in the high
// half, so that (word >> j) & 0x00010001 is the pair (2j, 2j+1)
__host__ __device__ constexpr int site_bit(int i) { return (i >> 1) + 16 * (i & 1); }

// shared memory carve-up (~182 KB)
struct TcSmem {
  uint32_t sbits[4096 / 32][kChains];        // chain states, word-major: sbits[w][chain], bit order = site_bit()
  __align__(128) __nv_bfloat16 b[kPipes][kBSlots][kKC / 8][kBlk / 8][8][8];
  __align__(16) float jblk[kBlk][kBlk + 4];  // J[blk, blk] as fp32, transposed: jblk[i][i'] = J[i0+i'][i0+i]
  __align__(8) uint64_t full[kPipes][kASlots];   // producers -> issuer: A slot written, J tile landed  (one arrival per warp)
  __align__(8) uint64_t empty[kPipes][kASlots];  // issuer -> producers: the MMAs reading the A slot (and its J slot) are done (commit)
  __align__(8) uint64_t acc_full[2];         // issuers -> epilogue: accumulator buffer complete          (one commit per pipeline)
  __align__(8) uint64_t acc_free[2];         // epilogue -> issuers: accumulator buffer read out          (one arrival per warp)
  __align__(8) uint64_t state_ready[4];      // epilogue -> producers: bits of block gb written (ring, one arrival per warp)
  uint32_t tmem_base;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// all lanes have done their part: one lane arrives for the warp (__syncwarp orders the lanes' writes before it)
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) dense_tc_kernel(TcParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = P.N;
  const int n_blocks = N / kBlk, n_chunks = N / kKC;
  const int total_blocks = n_blocks * P.n_sweeps;
  const int n_active = n_chunks < kPipes ? n_chunks : kPipes;  // pipelines that ever get a chunk
  static_assert(kKC == 4 * kBlk, "a K-chunk holds four blocks");

  // ---- one-time setup -------------------------------------------------------------------------
  if (tid < kChains) {  // pack this chain's bits
    const int chain = blockIdx.x * kChains + tid;
    for (int w = 0; w < N / 32; ++w) {
      uint32_t x = 0;
      if (chain < P.n_chains) {
        const uint8_t* src = P.state + (size_t)chain * N + 32 * w;
#pragma unroll
        for (int b = 0; b < 32; b += 4) {
          const uint32_t v = *reinterpret_cast<const uint32_t*>(src + b);  // sites b .. b+3
          x |= ((v & 1u) | ((v >> 15) & 2u)) << (b >> 1);                  // even sites b, b+2
          x |= (((v >> 8) & 1u) | ((v >> 23) & 2u)) << (16 + (b >> 1));    // odd sites b+1, b+3
        }
      }
      sm.sbits[w][tid] = x;
    }
  }
  if (tid == 0) {
    for (int q = 0; q < kPipes; ++q)
      for (int s = 0; s < kASlots; ++s) {
        mbar_init(&sm.full[q][s], 4);
        mbar_init(&sm.empty[q][s], 1);
      }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm.acc_full[s], n_active);
      mbar_init(&sm.acc_free[s], 4);
    }
    for (int s = 0; s < 4; ++s) mbar_init(&sm.state_ready[s], 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kIssuer0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&sm.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = sm.tmem_base;

  if (warp < kEpilogue0) {
    // ===================== producers: expand spins to bf16 A tiles, stream J tiles ======================
    // pipeline q = warp / 4 handles the chunks cc = q, q + kPipes, ... of every block; thread = chain (TMEM lane)
    const int q = warp >> 2, row = tid & (kChains - 1);
    if (q < n_active) {
      // J tile requests run kLook of the pipeline's own chunks ahead (across blocks and sweeps)
      int ld_gb = 0, ld_blk = 0, ld_cc = q, ld_slot = 0;
      auto load_b = [&]() {
        if (ld_gb < total_blocks) {
          int lkc = ld_cc + (ld_blk >> 2) + 1;
          if (lkc >= n_chunks) lkc -= n_chunks;
          const __nv_bfloat16* src = P.J + (size_t)(ld_blk * kBlk) * N + lkc * kKC;
          // 16-byte pieces: a warp instruction covers 16 rows x one 32-byte sector (full sectors from L2) and lands
          // in two 256-byte runs of the tile (4 shared-memory wavefronts, the minimum for 512 bytes)
          if (!(P.dbg & 4)) {
#pragma unroll
            for (int p = 0; p < kBlk * (kKC / 8) / kChains; ++p) {
              const int n = ((row & 31) >> 1) + 16 * (p & 1), k16 = 2 * ((row >> 5) + 4 * (p >> 1)) + (row & 1);
              cp_async16(&sm.b[q][ld_slot][k16][n >> 3][n & 7][0], src + (size_t)n * N + 8 * k16);
            }
          }
          if (++ld_slot == kBSlots) ld_slot = 0;
          ld_cc += kPipes;
          if (ld_cc >= n_chunks) {
            ld_cc = q;
            ++ld_gb;
            if (++ld_blk == n_blocks) ld_blk = 0;
          }
        }
        cp_async_commit();  // (an empty group at the tail keeps the group count uniform)
      };
      for (int i = 0; i < kLook; ++i) load_b();
      uint32_t empty_phase = 0;
      int ready_seen = 0;  // number of state_ready phases consumed (block gb needs gb of them for its last two chunks)
      int sa = 0;
      long long m = 0;     // own chunks done
      for (int gb = 0; gb < total_blocks; ++gb) {
        const int blk = gb % n_blocks;
        for (int cc = q; cc < n_chunks; cc += kPipes, ++m) {
          // chunk order: own + 1, ..., own - 1, own.  The two last chunks may hold the previous block's sites,
          // so they wait for that block's update; everything earlier only needs older state.
          int kc = cc + (blk >> 2) + 1;
          if (kc >= n_chunks) kc -= n_chunks;
          if (m >= kASlots) {  // the MMAs of own chunk m - kASlots are done: A slot sa and the J slot of chunk m + kLook are free
            TC_T0();
            mbar_wait(&sm.empty[q][sa], (empty_phase >> sa) & 1u);
            empty_phase ^= 1u << sa;
            if ((warp & 3) == 0) TC_ACC(3 + 4 * q);
          }
#ifdef TSU_TC_TIMING
          long long tl0__ = clock64();
#endif
          load_b();
#ifdef TSU_TC_TIMING
          if (warp == 0 && blockIdx.x == 0 && (tid & 31) == 0) atomicAdd(&g_tc_timing[16], (unsigned long long)(clock64() - tl0__));
#endif
          const int need = (cc >= n_chunks - 2) ? gb : gb - 1;
          if (ready_seen < need) {
            TC_T0();
            while (ready_seen < need) {
              mbar_wait(&sm.state_ready[ready_seen & 3], (uint32_t)((ready_seen >> 2) & 1));
              ++ready_seen;
            }
            if ((warp & 3) == 0) TC_ACC(4 + 4 * q);
          }
          TC_T0();
          // 128 bits of this chain -> 64 packed bf16 pairs -> 64 TMEM columns of the chain's lane
          const uint32_t a_col = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(kAccCols + (q * kASlots + sa) * kACols);
#pragma unroll
          for (int w4 = 0; w4 < kKC / 32 && !(P.dbg & 2); ++w4) {
            const uint32_t w = sm.sbits[(kKC / 32) * kc + w4][row];
            uint32_t r[16];
#pragma unroll
            for (int j = 0; j < 15; ++j) r[j] = (w << (14 - j)) & 0x40004000u;  // sites 2j (low half), 2j+1 (high half)
            r[15] = (w >> 1) & 0x40004000u;
            tmem_st16(a_col + 16 * w4, r);
          }
          if (warp == 0) TC_NEXT(17);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          if (warp == 0) TC_NEXT(18);
          cp_async_wait<kLook>();  // this thread's pieces of the J tile of this chunk have landed
          if (warp == 0) TC_NEXT(19);
          // NOTE: no fence.proxy.async here.  It lowers to MEMBAR.ALL.CTA, which waits for ALL of the thread's
          // outstanding memory operations - including the J-tile cp.asyncs just issued for kLook chunks ahead -
          // i.e. one full L2/HBM latency per chunk.  The issuer warp executes the proxy fence after its acquire.
          tc_fence_before();
          warp_arrive(&sm.full[q][sa]);
          if (warp == 0) TC_NEXT(20);
          if (warp == 4) TC_ACC(9);
          if (++sa == kASlots) sa = 0;
        }
      }
    }
  } else if (warp >= kIssuer0) {
    // ===================== MMA issuers: one elected lane per pipeline feeds the tensor core ================
    // The whole warp runs the loop (uniform control flow keeps counters and descriptors in uniform registers);
    // only the tcgen05.mma / commit instructions are issued by the elected lane.
    const int q = warp - kIssuer0;
    if (q < n_active) {
      const uint32_t idesc = umma_idesc(kChains, kBlk);
      constexpr uint32_t kLbo = (kBlk / 8) * 128, kSbo = 128;
      const uint64_t b_desc0 = umma_desc(smem_u32(&sm.b[q][0][0][0][0][0]), kLbo, kSbo);
      constexpr uint32_t kSlotUnits = (uint32_t)(sizeof(sm.b[0][0]) >> 4);  // 16-byte units per ring slot
      constexpr uint32_t kStepUnits = (2 * kLbo) >> 4;                       // per K = 16 step
      uint32_t full_phase = 0, free_phase = 0;
      int sa = 0, sb = 0;
      for (int gb = 0; gb < total_blocks; ++gb) {
        const int buf = gb & 1;
        if (gb >= 2) {  // the epilogue has read this accumulator buffer out
          TC_T0();
          mbar_wait(&sm.acc_free[buf], (free_phase >> buf) & 1u);
          free_phase ^= 1u << buf;
          TC_ACC(2);
        }
        const uint32_t d0 = tmem_d + (uint32_t)((buf * kPipes + q) * kBlk);
        for (int cc = q; cc < n_chunks; cc += kPipes) {
          {
            TC_T0();
            mbar_wait(&sm.full[q][sa], (full_phase >> sa) & 1u);
            TC_ACC(q);
          }
          full_phase ^= 1u << sa;
          TC_T0();
          fence_async_smem();  // producers' cp.async (generic proxy) writes, acquired above -> async proxy reads
          tc_fence_after();
          if (q == 0) TC_NEXT(21);
          if (elect_one()) {
            const uint64_t bd0 = b_desc0 + (uint64_t)((uint32_t)sb * kSlotUnits);
            const uint32_t a0 = tmem_d + (uint32_t)(kAccCols + (q * kASlots + sa) * kACols);
#pragma unroll
            for (int j = 0; j < kKC / 16 && !(P.dbg & 1); ++j)
              umma_bf16_ts(d0, a0 + 8u * j, bd0 + (uint64_t)(j * kStepUnits), idesc, (cc >= kPipes || j > 0) ? 1u : 0u);
            umma_commit(&sm.empty[q][sa]);
            if (cc + kPipes >= n_chunks) umma_commit(&sm.acc_full[buf]);  // this pipeline's share of the block is in
          }
          __syncwarp();
          if (q == 0) TC_NEXT(22);
          if (++sa == kASlots) sa = 0;
          if (++sb == kBSlots) sb = 0;
        }
      }
    }
  } else {
    // ===================== epilogue: fields out of TMEM, sequential update of the block ================
    const int row = tid - kProducers;                    // TMEM lane = chain within the tile
    const int chain = blockIdx.x * kChains + row;
    const bool chain_ok = chain < P.n_chains;
    const uint32_t tmem_lane = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    const float T = P.T_chain ? (float)P.T_chain[chain_ok ? chain : 0] : P.T;
    const uint32_t chain_g = P.chain0 + (uint32_t)chain;
    uint32_t accf_phase = 0;
    // diagonal block J[blk, blk]: 8 bf16 per thread (row i0 + row/4, columns i0 + 8 (row%4) ..), fetched one
    // block ahead so that the load latency hides behind the previous block's update
    auto load_diag = [&](int blk) {
      const int i0 = blk * kBlk;
      return __ldg(reinterpret_cast<const uint4*>(P.J + (size_t)(i0 + (row >> 2)) * N + i0 + 8 * (row & 3)));
    };
    uint4 jd = load_diag(0);
    for (int gb = 0; gb < total_blocks; ++gb) {
      const int blk = gb % n_blocks, sweep = gb / n_blocks, buf = gb & 1;
      const int i0 = blk * kBlk;
      // diagonal block as fp32, transposed: jblk[i][i'] = J[i0 + i', i0 + i]
      named_bar_sync(1, kChains);  // everybody is done with the previous block's jblk
      {
        const uint32_t jw[4] = {jd.x, jd.y, jd.z, jd.w};
        const int r = row >> 2, c0 = 8 * (row & 3);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          sm.jblk[c0 + 2 * e][r] = __uint_as_float(jw[e] << 16);  // bf16 -> fp32 is a 16-bit shift
          sm.jblk[c0 + 2 * e + 1][r] = __uint_as_float(jw[e] & 0xffff0000u);
        }
      }
      if (gb + 1 < total_blocks) jd = load_diag((gb + 1) % n_blocks);
      named_bar_sync(1, kChains);
      // Acceptance thresholds of the block, computed while the tensor core is still accumulating:
      //   u < sigmoid(h / T)  <=>  h > T * logit(u)          (gibbs.py:61-77,126; strict <)
      // and the clamp of gibbs.py:65-70 (|h/T| > 20 -> p = 1 / 0) is the clamp of the threshold to +-20 T.
      // This takes exp, the division and the Philox call off the site-to-site dependency chain: per site the
      // chain is compare -> select -> fma.
      float thr[kBlk];
      if (!P.gemm_only) {
#pragma unroll
        for (int i = 0; i < kBlk; i += 4) {
          const tsu_u32x4 o = tsu_philox4x32_10((uint32_t)((i0 + i) >> 2), chain_g, P.sweep0 + (uint32_t)sweep,
                                                TSU_STREAM_DENSE_TC, P.k0, P.k1);
          const uint32_t r4[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float u = (float)(r4[k] >> 8) * (1.0f / 16777216.0f);  // 24-bit uniform, exact in fp32
            const float lg = (__log2f(u) - __log2f(1.0f - u)) * 0.69314718056f;
            thr[i + k] = fminf(fmaxf(lg, -20.0f), 20.0f) * T;
          }
        }
      }
      {
        TC_T0();
        mbar_wait_backoff(&sm.acc_full[buf], (accf_phase >> buf) & 1u);
        if (warp == kEpilogue0) TC_ACC(11);
      }
      accf_phase ^= 1u << buf;
      TC_T0();
      tc_fence_after();
      float h[kBlk];
      tmem_ld32(tmem_lane + (uint32_t)(buf * kPipes * kBlk), h);
      if (n_active > 1) {
        float hp[kBlk];
        tmem_ld32(tmem_lane + (uint32_t)((buf * kPipes + 1) * kBlk), hp);
#pragma unroll
        for (int i = 0; i < kBlk; ++i) h[i] += hp[i];
      }
      tc_fence_before();
      warp_arrive(&sm.acc_free[buf]);  // the tensor core may overwrite this buffer (block gb + 2)
#pragma unroll
      for (int i = 0; i < kBlk; ++i) h[i] = fmaf(h[i], 0.5f, P.bias ? __ldg(P.bias + i0 + i) : 0.0f);  // spins were 0 / 2
      if (P.gemm_only) {
        if (P.fields_out && chain_ok) {
#pragma unroll
          for (int i = 0; i < kBlk; ++i) P.fields_out[(size_t)chain * N + i0 + i] = h[i];
        }
      } else {
        // sequential heat-bath update of the 32 sites of this block for this thread's chain (gibbs.py:153-160)
        static_assert(kBlk == 32, "one state word per block");
        const uint32_t w_old = sm.sbits[blk][row];
        uint32_t w_new = 0;
#pragma unroll
        for (int i = 0; i < kBlk; ++i) {
          if (P.fields_out && chain_ok) P.fields_out[(size_t)chain * N + i0 + i] = h[i];  // field at visit time
          const float d_up = ((w_old >> site_bit(i)) & 1u) ? 0.0f : 1.0f;   // new - old if the site comes out 1 ...
          const float d_dn = d_up - 1.0f;                                    // ... or 0 (both known before the chain)
          const bool up = h[i] > thr[i];
          const float delta = up ? d_up : d_dn;
          w_new |= up ? (1u << site_bit(i)) : 0u;
          // not yet visited sites of the block see the new value (rank-1 correction, branch free)
#pragma unroll
          for (int ip = i + 1; ip < kBlk; ++ip) h[ip] = fmaf(sm.jblk[i][ip], delta, h[ip]);
        }
        sm.sbits[blk][row] = w_new;
      }
      warp_arrive(&sm.state_ready[gb & 3]);  // release: the producers may expand chunks holding this block
      if (warp == kEpilogue0) TC_ACC(12);
    }
    if (!P.gemm_only && chain_ok) {  // unpack the final bits of this chain
      for (int w = 0; w < N / 32; ++w) {
        const uint32_t x = sm.sbits[w][row];
        uint8_t* dst = P.state + (size_t)chain * N + 32 * w;
#pragma unroll
        for (int b = 0; b < 32; b += 4) {
          const uint32_t e2 = x >> (b >> 1), o2 = x >> (16 + (b >> 1));
          *reinterpret_cast<uint32_t*>(dst + b) = (e2 & 1u) | ((o2 & 1u) << 8) | ((e2 & 2u) << 15) | ((o2 & 2u) << 23);
        }
      }
    }
  }
  // ---- teardown ----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kIssuer0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_d) : "memory");
  }
}

}  // namespace

static int launch_tc(const TcParams& P, cudaStream_t st) {
  const size_t smem = sizeof(TcSmem);
  cudaError_t e = cudaFuncSetAttribute(dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dense_tc_kernel<<<(P.n_chains + kChains - 1) / kChains, kThreads, smem, st>>>(P);
  e = cudaGetLastError();
  return e == cudaSuccess ? TSU_OK : (int)e;
}

extern "C" int tsu_dense_gibbs_tc_run(const void* d_J_bf16, const float* d_bias, uint8_t* d_state, int n_chains, int N,
                                      double T, const double* d_T_chain, int n_sweeps, uint64_t seed, uint32_t sweep0,
                                      uint32_t chain0, float* d_fields_or_null, uintptr_t stream) {
  TSU_CHECK_ARG(d_J_bf16 && d_state && n_chains > 0 && N > 0 && N % 128 == 0 && N <= 4096 && n_sweeps >= 0);
  TSU_CHECK_ARG(d_T_chain || T > 0);
  TcParams P = {};
  P.J = reinterpret_cast<const __nv_bfloat16*>(d_J_bf16);
  P.bias = d_bias;
  P.state = d_state;
  P.fields_out = d_fields_or_null;
  P.T_chain = d_T_chain;
  P.n_chains = n_chains;
  P.N = N;
  P.n_sweeps = n_sweeps;
  P.T = (float)T;
  P.k0 = (uint32_t)seed;
  P.k1 = (uint32_t)(seed >> 32);
  P.sweep0 = sweep0;
  P.chain0 = chain0;
  P.gemm_only = 0;
  if (const char* e = getenv("TSU_TC_DEBUG")) P.dbg = atoi(e);
  return launch_tc(P, tsu_stream(stream));
}

#ifdef TSU_TC_TIMING
extern "C" int tsu_dense_tc_debug_timing(unsigned long long* h_out16, int reset) {
  if (reset) {
    unsigned long long z[32] = {0};
    return (int)cudaMemcpyToSymbol(g_tc_timing, z, sizeof z);
  }
  return (int)cudaMemcpyFromSymbol(h_out16, g_tc_timing, sizeof(unsigned long long) * 32);
}
#endif

extern "C" int tsu_dense_tc_debug_fields(const void* d_J_bf16, const uint8_t* d_state, int n_chains, int N,
                                         float* d_fields, uintptr_t stream) {
  TSU_CHECK_ARG(d_J_bf16 && d_state && d_fields && n_chains > 0 && N > 0 && N % 128 == 0 && N <= 4096);
  TcParams P = {};
  P.J = reinterpret_cast<const __nv_bfloat16*>(d_J_bf16);
  P.state = const_cast<uint8_t*>(d_state);
  P.fields_out = d_fields;
  P.n_chains = n_chains;
  P.N = N;
  P.n_sweeps = 1;
  P.T = 1.0f;
  P.gemm_only = 1;
  if (const char* e = getenv("TSU_TC_DEBUG")) P.dbg = atoi(e);
  return launch_tc(P, tsu_stream(stream));
}

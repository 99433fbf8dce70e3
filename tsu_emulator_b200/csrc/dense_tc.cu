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
// for the whole sweep (64 KB); each K-chunk of 64 sites is expanded to a bf16 operand tile in the
// canonical no-swizzle K-major UMMA layout, the matching J tile is streamed from L2 with cp.async.

#include <cuda_bf16.h>

#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr int kChains = 128;  // chains per CTA = UMMA M = TMEM lanes
constexpr int kBlk = 32;      // sites per block = UMMA N
constexpr int kKC = 64;       // K-chunk (sites) per pipeline stage
constexpr int kAStages = 4;   // expanded spin tiles (a stage is reused 4 chunks later, when its MMAs are long done)
constexpr int kBStages = 12;  // J tile ring
constexpr int kLook = 8;      // J tiles are requested 8 chunks ahead (L2/HBM latency); ring slack = 4 chunks

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

struct TcParams {
  const __nv_bfloat16* J;   // [N][N] row-major coupling matrix (row i = couplings INTO site i)
  const float* bias;        // [N] or nullptr
  uint8_t* state;           // [n_chains][N] bits, updated in place
  float* fields_out;        // debug: [n_chains][N] fields seen at visit time (nullptr in production)
  const double* T_chain;    // [n_chains] or nullptr
  double* energy;           // [n_chains] or nullptr
  int n_chains, N, n_sweeps;
  float T;
  uint32_t k0, k1, sweep0, chain0;
  int gemm_only;            // debug: skip the spin update (fields of the initial state for every site)
};

// shared memory carve-up
struct TcSmem {
  uint32_t sbits[4096 / 32][kChains];               // chain states, word-major: sbits[w][chain]   (N <= 4096)
  __align__(128) __nv_bfloat16 a[kAStages][kKC / 8][kChains / 8][8][8];  // [k16B][row group][row][8 elems]
  __align__(128) __nv_bfloat16 b[kBStages][kKC / 8][kBlk / 8][8][8];
  __align__(16) float jblk[kBlk][kBlk + 4];           // J[blk, blk] as fp32, transposed: jblk[i][i'] = J[i0+i'][i0+i]
  __align__(16) uint4 lut[256];                       // byte -> 8 bf16 (0.0 / 1.0)
  __align__(8) uint64_t a_done[kAStages];             // MMAs that read the A stage have completed
  __align__(8) uint64_t b_done[kBStages];             // MMAs that read the B stage have completed
  __align__(8) uint64_t acc_done;                     // accumulator complete
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 1) dense_tc_kernel(TcParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = P.N;
  const int chain = blockIdx.x * kChains + tid;          // TMEM lane tid <-> chain
  const bool chain_ok = chain < P.n_chains;
  const int n_blocks = N / kBlk, n_chunks = N / kKC;

  // ---- one-time setup -------------------------------------------------------------------------
  for (int i = tid; i < 256; i += 128) {
    uint32_t w[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) w[p] = ((i >> (2 * p)) & 1 ? 0x3F80u : 0u) | ((i >> (2 * p + 1)) & 1 ? 0x3F800000u : 0u);
    sm.lut[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (int w = 0; w < N / 32; ++w) {  // pack this chain's bits
    uint32_t x = 0;
    if (chain_ok) {
      const uint8_t* src = P.state + (size_t)chain * N + 32 * w;
#pragma unroll
      for (int b = 0; b < 32; b += 4) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(src + b);
        x |= ((v & 1u) | ((v >> 7) & 2u) | ((v >> 14) & 4u) | ((v >> 21) & 8u)) << b;
      }
    }
    sm.sbits[w][tid] = x;
  }
  if (tid == 0) {
    for (int s = 0; s < kAStages; ++s) mbar_init(&sm.a_done[s], 1);
    for (int s = 0; s < kBStages; ++s) mbar_init(&sm.b_done[s], 1);
    mbar_init(&sm.acc_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&sm.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = sm.tmem_base;
  const uint32_t tmem_lane = tmem_d + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc = umma_idesc(kChains, kBlk);
  const float T = P.T_chain ? (float)P.T_chain[chain_ok ? chain : 0] : P.T;
  const float invT = 1.0f / T;

  uint32_t a_phase = 0, b_phase = 0;   // bit s = parity to wait for on a_done[s] / b_done[s]
  uint32_t acc_phase = 0;
  // producer position (J tile requests run kLook chunks ahead of consumption, across blocks and sweeps);
  // all indices are kept incrementally - no divisions in the chunk loop
  int ld_sweep = 0, ld_blk = 0, ld_cc = 0, ld_stage = 0;
  long long ld_count = 0;
  auto load_b = [&]() {
    if (ld_sweep < P.n_sweeps) {
      int lkc = ld_cc + (ld_blk >> 1) + 1;  // (ld_blk * kBlk) / kKC == ld_blk / 2
      if (lkc >= n_chunks) lkc -= n_chunks;
      if (ld_count >= kBStages) {  // the MMAs of the chunk that used this ring slot kBStages chunks ago
        mbar_wait(&sm.b_done[ld_stage], (b_phase >> ld_stage) & 1u);
        b_phase ^= 1u << ld_stage;
      }
#pragma unroll
      for (int p = 0; p < kBlk * 8 / 128; ++p) {
        const int piece = tid + 128 * p;
        const int n = piece >> 3, k16 = piece & 7;
        cp_async16(&sm.b[ld_stage][k16][n >> 3][n & 7][0], P.J + (size_t)(ld_blk * kBlk + n) * N + lkc * kKC + 8 * k16);
      }
      ++ld_count;
      if (++ld_stage == kBStages) ld_stage = 0;
      if (++ld_cc == n_chunks) {
        ld_cc = 0;
        if (++ld_blk == n_blocks) {
          ld_blk = 0;
          ++ld_sweep;
        }
      }
    }
    cp_async_commit();  // (an empty group at the tail keeps the group count uniform)
  };
  static_assert(kKC == 2 * kBlk, "chunk index of a block is blk / 2");

  for (int i = 0; i < kLook; ++i) load_b();
  long long g = 0;  // chunks consumed so far
  int sa = 0, sb = 0;

  for (int sweep = 0; sweep < P.n_sweeps; ++sweep) {
    for (int blk = 0; blk < n_blocks; ++blk) {
      const int i0 = blk * kBlk;
      // diagonal block J[blk, blk]: 8 bf16 per thread now (row i0 + tid/4, columns i0 + 8 (tid%4) ..), used after
      // the GEMM, so the load latency hides behind the chunk loop
      const uint4 jd = __ldg(reinterpret_cast<const uint4*>(P.J + (size_t)(i0 + (tid >> 2)) * N + i0 + 8 * (tid & 3)));
      // ---- GEMM: H[chain, i] = sum_k S[chain, k] * J[i0 + i, k] -----------------------------------
      // chunk order: the chunk holding this block's own sites goes last (it is the one the previous
      // block's update has just modified); all other chunks only need older state
      for (int cc = 0; cc < n_chunks; ++cc, ++g) {
        int kc = cc + (blk >> 1) + 1;  // ends with the chunk holding this block
        if (kc >= n_chunks) kc -= n_chunks;
        load_b();
        if (g >= kAStages) {  // the MMAs of chunk g - kAStages read this A stage
          mbar_wait(&sm.a_done[sa], (a_phase >> sa) & 1u);
          a_phase ^= 1u << sa;
        }
        // A tile: this chain's 64 bits of the chunk -> 64 bf16 (8 x 16 B, one per 8-element K group)
        {
          const uint32_t w0 = sm.sbits[2 * kc][tid], w1 = sm.sbits[2 * kc + 1][tid];
#pragma unroll
          for (int k16 = 0; k16 < 8; ++k16) {
            const uint32_t byte = ((k16 < 4 ? w0 : w1) >> (8 * (k16 & 3))) & 0xFFu;
            *reinterpret_cast<uint4*>(&sm.a[sa][k16][tid >> 3][tid & 7][0]) = sm.lut[byte];
          }
        }
        cp_async_wait<kLook>();         // J tile of chunk g has landed (this thread's pieces)
        fence_async_smem();             // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < kKC / 16; ++j) {
            const uint64_t ad = umma_desc(smem_u32(&sm.a[sa][2 * j][0][0][0]), (kChains / 8) * 128, 128);
            const uint64_t bd = umma_desc(smem_u32(&sm.b[sb][2 * j][0][0][0]), (kBlk / 8) * 128, 128);
            umma_bf16(tmem_d, ad, bd, idesc, (cc > 0 || j > 0) ? 1u : 0u);
          }
          umma_commit(&sm.a_done[sa]);
          umma_commit(&sm.b_done[sb]);
          if (cc == n_chunks - 1) umma_commit(&sm.acc_done);
        }
        if (++sa == kAStages) sa = 0;
        if (++sb == kBStages) sb = 0;
      }
      // diagonal block as fp32, transposed: jblk[i][i'] = J[i0 + i', i0 + i] (what a flip of site i adds to the
      // field of i')
      {
        const uint32_t jw[4] = {jd.x, jd.y, jd.z, jd.w};
        const int r = tid >> 2, c0 = 8 * (tid & 3);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          sm.jblk[c0 + 2 * e][r] = __uint_as_float(jw[e] << 16);              // bf16 -> fp32 is a 16-bit shift
          sm.jblk[c0 + 2 * e + 1][r] = __uint_as_float(jw[e] & 0xffff0000u);
        }
      }
      __syncthreads();
      // ---- epilogue: fields out of TMEM, sequential update of the block -------------------------
      mbar_wait(&sm.acc_done, acc_phase);
      acc_phase ^= 1u;
      tc_fence_after();
      float h[kBlk];
      tmem_ld32(tmem_lane + 0, h);
      if (kBlk > 32) tmem_ld32(tmem_lane + 32, h + 32);
      tc_fence_before();
      if (P.bias) {
#pragma unroll
        for (int i = 0; i < kBlk; ++i) h[i] += __ldg(P.bias + i0 + i);
      }
      if (P.fields_out && chain_ok && P.gemm_only) {  // fields as accumulated by the tensor core
#pragma unroll
        for (int i = 0; i < kBlk; ++i) P.fields_out[(size_t)chain * N + i0 + i] = h[i];
      }
      if (!P.gemm_only) {
        // sequential heat-bath update of the 64 sites of this block for this thread's chain (gibbs.py:153-160)
        static_assert(kBlk == 32, "one state word per block");
        uint32_t w[1] = {sm.sbits[blk][tid]};
        const uint32_t chain_g = P.chain0 + (uint32_t)chain;
        tsu_u32x4 o = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int i = 0; i < kBlk; ++i) {
          if ((i & 3) == 0)  // one Philox block serves 4 consecutive sites of a chain
            o = tsu_philox4x32_10((uint32_t)((i0 + i) >> 2), chain_g, P.sweep0 + (uint32_t)sweep, TSU_STREAM_DENSE_TC,
                                  P.k0, P.k1);
          const uint32_t r32 = (i & 3) == 0 ? o.x : ((i & 3) == 1 ? o.y : ((i & 3) == 2 ? o.z : o.w));
          const float u = (float)(r32 >> 8) * (1.0f / 16777216.0f);   // 24-bit uniform, exact in fp32
          if (P.fields_out && chain_ok) P.fields_out[(size_t)chain * N + i0 + i] = h[i];  // field at visit time
          const float x = h[i] * invT;
          float pacc = 1.0f / (1.0f + __expf(-x));                     // gibbs.py:61-77 incl. the clamp
          pacc = x > 20.0f ? 1.0f : (x < -20.0f ? 0.0f : pacc);
          const uint32_t nb = u < pacc ? 1u : 0u;                      // gibbs.py:126 (strict <)
          const uint32_t ob = (w[i >> 5] >> (i & 31)) & 1u;
          const float delta = (float)nb - (float)ob;
          w[i >> 5] = (w[i >> 5] & ~(1u << (i & 31))) | (nb << (i & 31));
          // not yet visited sites of the block see the new value (rank-1 correction, branch free)
#pragma unroll
          for (int ip = i + 1; ip < kBlk; ++ip) h[ip] = fmaf(sm.jblk[i][ip], delta, h[ip]);
        }
        sm.sbits[blk][tid] = w[0];
      }
      __syncthreads();  // jblk reuse, sbits of this block final before the next block's last chunk
    }
  }
  if (!P.gemm_only && chain_ok) {  // unpack the final bits of this chain
    for (int w = 0; w < N / 32; ++w) {
      const uint32_t x = sm.sbits[w][tid];
      uint8_t* dst = P.state + (size_t)chain * N + 32 * w;
#pragma unroll
      for (int b = 0; b < 32; b += 4) {
        const uint32_t n4 = (x >> b) & 15u;
        *reinterpret_cast<uint32_t*>(dst + b) = (n4 & 1u) | ((n4 & 2u) << 7) | ((n4 & 4u) << 14) | ((n4 & 8u) << 21);
      }
    }
  }
  // ---- teardown ----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_d) : "memory");
  }
}

}  // namespace

static int launch_tc(const TcParams& P, cudaStream_t st) {
  const size_t smem = sizeof(TcSmem);
  cudaError_t e = cudaFuncSetAttribute(dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dense_tc_kernel<<<(P.n_chains + kChains - 1) / kChains, 128, smem, st>>>(P);
  e = cudaGetLastError();
  return e == cudaSuccess ? TSU_OK : (int)e;
}

extern "C" int tsu_dense_gibbs_tc_run(const void* d_J_bf16, const float* d_bias, uint8_t* d_state, int n_chains, int N,
                                      double T, const double* d_T_chain, int n_sweeps, uint64_t seed, uint32_t sweep0,
                                      uint32_t chain0, float* d_fields_or_null, uintptr_t stream) {
  TSU_CHECK_ARG(d_J_bf16 && d_state && n_chains > 0 && N > 0 && N % 64 == 0 && N <= 4096 && n_sweeps >= 0);
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
  return launch_tc(P, tsu_stream(stream));
}

extern "C" int tsu_dense_tc_debug_fields(const void* d_J_bf16, const uint8_t* d_state, int n_chains, int N,
                                         float* d_fields, uintptr_t stream) {
  TSU_CHECK_ARG(d_J_bf16 && d_state && d_fields && n_chains > 0 && N > 0 && N % 64 == 0 && N <= 4096);
  TcParams P = {};
  P.J = reinterpret_cast<const __nv_bfloat16*>(d_J_bf16);
  P.state = const_cast<uint8_t*>(d_state);
  P.fields_out = d_fields;
  P.n_chains = n_chains;
  P.N = N;
  P.n_sweeps = 1;
  P.T = 1.0f;
  P.gemm_only = 1;
  return launch_tc(P, tsu_stream(stream));
}

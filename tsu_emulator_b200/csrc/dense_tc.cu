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
constexpr int kBlk = 64;      // sites per block = UMMA N
constexpr int kKC = 64;       // K-chunk (sites) per pipeline stage
constexpr int kStages = 3;

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
  __align__(128) __nv_bfloat16 a[kStages][kKC / 8][kChains / 8][8][8];  // [k16B][row group][row][8 elems]
  __align__(128) __nv_bfloat16 b[kStages][kKC / 8][kBlk / 8][8][8];
  __align__(16) float jblk[kBlk][kBlk + 4];           // J[blk, blk] as fp32: jblk[i'][i]
  __align__(16) uint4 lut[256];                       // byte -> 8 bf16 (0.0 / 1.0)
  __align__(8) uint64_t mma_done[kStages];            // stage buffers free again
  __align__(8) uint64_t acc_done;                     // accumulator complete
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 1) dense_tc_kernel(TcParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
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
    for (int s = 0; s < kStages; ++s) mbar_init(&sm.mma_done[s], 1);
    mbar_init(&sm.acc_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&sm.tmem_base)) : "memory");
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

  uint32_t stage_phase = 0;   // bit s = parity to wait for on mma_done[s]
  uint32_t stage_used = 0;    // bit s = stage s has an MMA group in flight
  uint32_t acc_phase = 0;
  int issue = 0;              // running chunk counter (stage = issue % kStages)

  for (int sweep = 0; sweep < P.n_sweeps; ++sweep) {
    for (int blk = 0; blk < n_blocks; ++blk) {
      const int i0 = blk * kBlk;
      // J[blk, blk] as fp32 for the in-block rank-1 corrections
      for (int e = tid; e < kBlk * kBlk; e += 128) {
        const int r = e / kBlk, c = e - r * kBlk;
        sm.jblk[r][c] = __bfloat162float(P.J[(size_t)(i0 + r) * N + i0 + c]);
      }
      // ---- GEMM: H[chain, i] = sum_k S[chain, k] * J[i0 + i, k] -----------------------------------
      // chunk order: the chunk holding this block's own sites goes last (it is the one the previous
      // block's update has just modified); all other chunks only need older state
      for (int cc = 0; cc < n_chunks; ++cc) {
        const int kc = (cc + blk + 1) % n_chunks;  // ends with kc == blk
        const int st = issue % kStages;
        if ((stage_used >> st) & 1u) {  // wait until the MMAs that read this stage have completed
          mbar_wait(&sm.mma_done[st], (stage_phase >> st) & 1u);
          stage_phase ^= 1u << st;
        }
        // B tile: rows i0 .. i0+63 of J, columns kc*64 .. +63  (4 x 16 B per thread, coalesced)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int piece = tid + 128 * p;
          const int n = piece >> 3, k16 = piece & 7;
          cp_async16(&sm.b[st][k16][n >> 3][n & 7][0], P.J + (size_t)(i0 + n) * N + kc * kKC + 8 * k16);
        }
        cp_async_commit();
        // A tile: this chain's 64 bits of the chunk -> 64 bf16 (8 x 16 B, one per 8-element K group)
        {
          const uint32_t w0 = sm.sbits[2 * kc][tid], w1 = sm.sbits[2 * kc + 1][tid];
#pragma unroll
          for (int k16 = 0; k16 < 8; ++k16) {
            const uint32_t byte = ((k16 < 4 ? w0 : w1) >> (8 * (k16 & 3))) & 0xFFu;
            *reinterpret_cast<uint4*>(&sm.a[st][k16][tid >> 3][tid & 7][0]) = sm.lut[byte];
          }
        }
        cp_async_wait<0>();
        fence_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < kKC / 16; ++j) {
            const uint64_t ad = umma_desc(smem_u32(&sm.a[st][2 * j][0][0][0]), (kChains / 8) * 128, 128);
            const uint64_t bd = umma_desc(smem_u32(&sm.b[st][2 * j][0][0][0]), (kBlk / 8) * 128, 128);
            umma_bf16(tmem_d, ad, bd, idesc, (cc > 0 || j > 0) ? 1u : 0u);
          }
          umma_commit(&sm.mma_done[st]);
          if (cc == n_chunks - 1) umma_commit(&sm.acc_done);
        }
        stage_used |= 1u << st;
        ++issue;
      }
      // ---- epilogue: fields out of TMEM, sequential update of the block -------------------------
      mbar_wait(&sm.acc_done, acc_phase);
      acc_phase ^= 1u;
      tc_fence_after();
      float h[kBlk];
      tmem_ld32(tmem_lane + 0, h);
      tmem_ld32(tmem_lane + 32, h + 32);
      tc_fence_before();
      if (P.bias) {
#pragma unroll
        for (int i = 0; i < kBlk; ++i) h[i] += __ldg(P.bias + i0 + i);
      }
      if (P.fields_out && chain_ok) {
#pragma unroll
        for (int i = 0; i < kBlk; ++i) P.fields_out[(size_t)chain * N + i0 + i] = h[i];
      }
      (void)invT;
      __syncthreads();  // jblk reuse, sbits of this block final before the next block's last chunk
    }
  }
  // ---- teardown ----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_d) : "memory");
  }
}

}  // namespace

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
  const size_t smem = sizeof(TcSmem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dense_tc_kernel<<<(n_chains + kChains - 1) / kChains, 128, smem, tsu_stream(stream)>>>(P);
  TSU_RETURN_LAUNCH_STATUS();
}

// Dense-coupling heat-bath Gibbs sampler: exact sequential single-site updates, one CTA per chain,
// incrementally maintained local fields (sm_100a).
//
// Replaces the hot loop of the reference
//   GibbsSampler.gibbs_sweep / sample_conditional / _compute_local_field  tsu/gibbs.py:79-162
//   GibbsSampler.sample_boltzmann (burn-in + n_samples x n_sweeps)        tsu/gibbs.py:164-213
//   GibbsSampler.compute_energy                                           tsu/gibbs.py:215-236
// for a batch of independent chains that share one coupling matrix (the "parallel chains" that
// HardwareEmulator.sample_parallel, gibbs.py:450-487, and parallel_tempering, gibbs.py:284-303,
// run one after the other).
//
// The reference recomputes h_i = J[i,:] . s + b_i (O(N)) for every visited site.  Here every chain
// keeps all N local fields in shared memory; a site visit is O(1) (read h_i, compare) and only an
// actual flip costs O(N): h_j += J[j,i] * (new - old) for all j, one coalesced row of the transposed
// coupling matrix read from L2 with 2-4 fields per thread, so that a flip costs about one L2 round
// trip.  With float64 fields the comparison u < sigmoid(h_i / T) is made as logit(u) < h_i / T, the
// logits of a sweep's uniforms being computed in parallel beforehand; whenever the two sides are
// within 1e-4 of each other, or near the +-20 clamp, the reference's own expression decides, so the
// bits are the reference's (tests/test_dense_gpu.py, uniforms placed on and next to p).  The self
// term J_ii s_i stays part of h_i exactly as in gibbs.py:97.  One __syncthreads per visited site:
// writes to h_i / s_i of the site just decided are deferred by one step so that slow readers of the
// same step never race with them.  Per site visit, 296 chains: 0.49 us at N = 512, 0.8 us at 2048,
// 1.7 us at 4096 (was 2.1 / 2.9 / 5.7 us with 16 fields per thread and the sigmoid in the chain).

#include "common.cuh"
#include "philox.cuh"

namespace {

struct DenseParams {
  const void* Jt;  // [N][N] row-major transpose of J: Jt[i*N + j] = J[j][i]
  const void* bias;
  uint8_t* state;
  const double* T_chain;
  const double* T_sweep;
  const int32_t* order;
  const double* uniforms;
  uint8_t* samples;
  double* energy;
  uint8_t* best_state;
  double* best_energy;
  double T;
  int n_chains, N;
  int n_visit;  // sites visited per sweep (N for a full sweep)
  int n_burnin, n_samples, sweeps_per_sample;
  int track_best;
  int refresh_every;  // recompute fields from scratch every this many sweeps (float32 fields)
  uint32_t k0, k1, sweep0, chain0;
};

template <typename AT>
__device__ __forceinline__ AT sigmoid_clamped(AT x);

// tsu/gibbs.py:61-77: x > 20 -> 1, x < -20 -> 0 (strict), else 1 / (1 + exp(-x))
template <>
__device__ __forceinline__ double sigmoid_clamped<double>(double x) {
  if (x > 20.0) return 1.0;
  if (x < -20.0) return 0.0;
  return 1.0 / (1.0 + exp(-x));
}
template <>
__device__ __forceinline__ float sigmoid_clamped<float>(float x) {
  if (x > 20.0f) return 1.0f;
  if (x < -20.0f) return 0.0f;
  return 1.0f / (1.0f + expf(-x));
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nw; ++w) t += red[w];  // same order in every thread: deterministic
  return t;
}

template <typename JT, typename AT>
__device__ void compute_fields(const DenseParams& P, const uint8_t* s, AT* h) {
  const JT* Jt = reinterpret_cast<const JT*>(P.Jt);
  const JT* b = reinterpret_cast<const JT*>(P.bias);
  const int N = P.N;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    AT acc = (AT)0;
    for (int k = 0; k < N; ++k)
      if (s[k]) acc += (AT)Jt[(size_t)k * N + j];  // h_j = sum_k J[j][k] s_k  (ascending k)
    if (b) acc += (AT)b[j];
    h[j] = acc;
  }
}

// E = -1/2 s^T J s - b^T s = -1/2 sum_i s_i h_i - 1/2 sum_i b_i s_i   (h includes the bias)
template <typename JT, typename AT>
__device__ double chain_energy(const DenseParams& P, const uint8_t* s, const AT* h, double* red) {
  const JT* b = reinterpret_cast<const JT*>(P.bias);
  double acc = 0.0;
  for (int j = threadIdx.x; j < P.N; j += blockDim.x)
    if (s[j]) acc += -0.5 * (double)h[j] - (b ? 0.5 * (double)b[j] : 0.0);
  return block_sum(acc, red);
}

template <typename JT, typename AT>
__global__ void __launch_bounds__(1024) dense_gibbs_kernel(DenseParams P) {  // pick_threads() goes up to 1024
  extern __shared__ double smem_d[];
  const int N = P.N;
  constexpr bool kLogit = sizeof(AT) == sizeof(double);  // float64 fields: decide by logit(u) < h / T where that is safe
  double* u = smem_d;                                   // [N] uniforms of the current sweep (visit order)
  double* lg = u + N;                                   // [N] logit(u) = log(u) - log(1 - u)   (float64 fields only)
  double* red = lg + (kLogit ? N : 0);                  // [32]
  AT* h = reinterpret_cast<AT*>(red + 32);              // [N] local fields
  AT* diag = h + N;                                     // [N] self couplings J_ii
  uint8_t* s = reinterpret_cast<uint8_t*>(diag + N);    // [N] bits
  const int chain = blockIdx.x;
  const JT* Jt = reinterpret_cast<const JT*>(P.Jt);
  uint8_t* gstate = P.state + (size_t)chain * N;

  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    s[j] = gstate[j] ? 1 : 0;
    diag[j] = (AT)Jt[(size_t)j * N + j];
  }
  __syncthreads();
  compute_fields<JT, AT>(P, s, h);
  __syncthreads();

  double best_e = 0.0;
  if (P.track_best) {
    best_e = chain_energy<JT, AT>(P, s, h, red);
    for (int j = threadIdx.x; j < N; j += blockDim.x) P.best_state[(size_t)chain * N + j] = s[j];
  }

  const int total = P.n_burnin + P.n_samples * P.sweeps_per_sample;
  int next_sample_at = P.n_burnin + P.sweeps_per_sample;
  int sample_idx = 0;
  for (int sw = 0; sw < total; ++sw) {
    const double Td = P.T_chain ? P.T_chain[chain] : (P.T_sweep ? P.T_sweep[sw] : P.T);
    const AT T = (AT)Td;
    if (P.refresh_every > 0 && sw > 0 && (sw % P.refresh_every) == 0) {
      compute_fields<JT, AT>(P, s, h);
    }
    const int NV = P.n_visit;
    const int32_t* ord = P.order ? P.order + (size_t)sw * NV : nullptr;
    // uniforms of this sweep, in visiting order
    for (int idx = threadIdx.x; idx < NV; idx += blockDim.x) {
      if (P.uniforms) {
        u[idx] = P.uniforms[((size_t)sw * P.n_chains + chain) * NV + idx];
      } else {
        const int site = ord ? ord[idx] : idx;
        tsu_u32x4 o = tsu_philox4x32_10((uint32_t)site, P.chain0 + (uint32_t)chain, P.sweep0 + (uint32_t)sw,
                                        TSU_STREAM_DENSE, P.k0, P.k1);
        const unsigned long long m = (((unsigned long long)o.x << 32) | o.y) >> 11;
        u[idx] = (double)m * (1.0 / 9007199254740992.0);
      }
      if (kLogit) lg[idx] = log(u[idx]) - log1p(-u[idx]);  // in parallel, off the site-by-site chain (-inf at u = 0)
    }
    __syncthreads();
    const double invT = 1.0 / Td;

    int pend_i = -1;       // site decided in the previous step: its s / self-term writes are deferred
    int pend_bit = 0;
    AT pend_delta = (AT)0;
    for (int idx = 0; idx < NV; ++idx) {
      const int i = ord ? ord[idx] : idx;
      const AT hi = h[i];
      const int si = s[i];
      const double ui = u[idx];
      // deferred writes of the previous step (all threads have finished reading that site)
      if (pend_i >= 0) {
        if (threadIdx.x == (pend_i % blockDim.x)) {
          s[pend_i] = (uint8_t)pend_bit;
          if (pend_delta != (AT)0) h[pend_i] += diag[pend_i] * pend_delta;
        }
      }
      // gibbs.py:124-126: p = sigmoid(h_i / T) with the +-20 clamp, new bit = u < p (strict).  u < sigmoid(x) <=>
      // logit(u) < x: the exp and the division leave the sequential chain whenever the two sides are further apart
      // than anything rounding can do (float64 sigmoid: <= 2e-7 in logit units for |x| < 20; logit(u) and h * (1 / T):
      // 1e-14); otherwise - and next to the clamps - the reference's expression itself decides.  Same bits.
      int nb;
      if (kLogit) {
        const double xq = (double)hi * invT, t = lg[idx];
        if (xq > 20.000001)
          nb = ui < 1.0 ? 1 : 0;                         // x > 20: p = 1.0
        else if (xq < -20.000001)
          nb = 0;                                        // x < -20: p = 0.0
        else if (fabs(xq) < 19.999999 && fabs(xq - t) > 1e-4)
          nb = t < xq ? 1 : 0;
        else
          nb = (ui < (double)sigmoid_clamped<AT>(hi / T)) ? 1 : 0;
      } else {
        nb = (ui < (double)sigmoid_clamped<AT>(hi / T)) ? 1 : 0;
      }
      const AT delta = (AT)(nb - si);
      if (delta != (AT)0) {
        const JT* row = Jt + (size_t)i * N;               // column i of J
        // few fields per thread and the loads of one thread batched: a flip costs about one L2 round trip
#pragma unroll 4
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
          const AT v = (AT)__ldg(row + j);
          if (j != i) h[j] += v * delta;
        }
      }
      pend_i = i;
      pend_bit = nb;
      pend_delta = delta;
      __syncthreads();
    }
    if (pend_i >= 0 && threadIdx.x == (pend_i % blockDim.x)) {
      s[pend_i] = (uint8_t)pend_bit;
      if (pend_delta != (AT)0) h[pend_i] += diag[pend_i] * pend_delta;
    }
    __syncthreads();

    if (P.track_best) {  // gibbs.py:387-391
      const double e = chain_energy<JT, AT>(P, s, h, red);
      if (e < best_e) {
        best_e = e;
        for (int j = threadIdx.x; j < N; j += blockDim.x) P.best_state[(size_t)chain * N + j] = s[j];
      }
    }
    if (P.samples && sw + 1 == next_sample_at && sample_idx < P.n_samples) {
      uint8_t* dst = P.samples + ((size_t)sample_idx * P.n_chains + chain) * N;
      for (int j = threadIdx.x; j < N; j += blockDim.x) dst[j] = s[j];
      ++sample_idx;
      next_sample_at += P.sweeps_per_sample;
    } else if (sw + 1 == next_sample_at) {
      ++sample_idx;
      next_sample_at += P.sweeps_per_sample;
    }
  }
  for (int j = threadIdx.x; j < N; j += blockDim.x) gstate[j] = s[j];
  if (P.energy) {
    const double e = chain_energy<JT, AT>(P, s, h, red);
    if (threadIdx.x == 0) P.energy[chain] = e;
  }
  if (P.track_best && threadIdx.x == 0) P.best_energy[chain] = best_e;
}

// ---- models of at most 32 bits: one WARP per chain ---------------------------------------------------------
// The reference's published benchmarks (tsu/benchmarks/sampling.py, optimization.py: 1-20 bits, one chain, up to 10^5
// sweeps per call) are pure latency: a visit of the kernel above is a block barrier and shared-memory round trips
// (~340 ns).  Here lane j keeps field h_j in a register, the bits are a warp-uniform mask, the couplings sit in shared
// memory and a visit is one shuffle, one compare and one FMA.  Same arithmetic in the same order as the kernel above
// (initial fields summed over ascending k, one J * delta added per flip, the energy reduced by the same xor shuffles),
// so the bits, energies and best states are identical (goldens dense_*_n8 ... n14, tempering_n10).
constexpr int kWarpChainMaxN = 32;
constexpr int kWarpChainsPerCta = 4;

template <typename JT>
__global__ void __launch_bounds__(32 * kWarpChainsPerCta) dense_gibbs_warp_kernel(DenseParams P) {
  extern __shared__ double smem_d[];
  const int N = P.N;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  double* Jsm = smem_d;                                  // [N][N] Jt as double: Jsm[i * N + j] = J[j][i]
  double* lg = Jsm + N * N + wic * 64;                   // [32] logit(u) of this warp's current sweep (visit order)
  double* u = lg + 32;                                   // [32] the uniforms themselves (exact fallback)
  const JT* Jt = reinterpret_cast<const JT*>(P.Jt);
  const JT* b = reinterpret_cast<const JT*>(P.bias);
  for (int k = threadIdx.x; k < N * N; k += blockDim.x) Jsm[k] = (double)Jt[k];
  __syncthreads();
  const int chain = blockIdx.x * kWarpChainsPerCta + wic;
  if (chain >= P.n_chains) return;
  const bool mine = lane < N;
  uint8_t* gstate = P.state + (size_t)chain * N;
  const uint32_t my_bit = mine && gstate[lane] ? 1u : 0u;
  uint32_t mask = __ballot_sync(0xffffffffu, my_bit != 0u);   // bit j = s_j, the same in every lane
  const double bj = (mine && b) ? (double)b[lane] : 0.0;
  double h = 0.0;
  if (mine) {
    for (int k = 0; k < N; ++k)
      if ((mask >> k) & 1u) h += Jsm[k * N + lane];      // h_j = sum_k J[j][k] s_k (ascending k)
    if (b) h += bj;
  }
  auto energy_now = [&]() {                               // -1/2 sum_j s_j h_j - 1/2 sum_j b_j s_j, as chain_energy()
    double acc = 0.0;
    if (mine && ((mask >> lane) & 1u)) acc += -0.5 * h - (b ? 0.5 * bj : 0.0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return 0.0 + acc;
  };
  double best_e = 0.0;
  if (P.track_best) {
    best_e = energy_now();
    if (mine) P.best_state[(size_t)chain * N + lane] = (uint8_t)((mask >> lane) & 1u);
  }
  const int total = P.n_burnin + P.n_samples * P.sweeps_per_sample;
  int next_sample_at = P.n_burnin + P.sweeps_per_sample;
  int sample_idx = 0;
  const int NV = P.n_visit;
  // visiting order and uniforms (with their logits: two float64 logarithms, most of a short sweep's time) are produced
  // for as many sweeps at once as fit the 32 lanes: lane l serves visit l % NV of sweep sw + l / NV
  const int batch = 32 / NV > 0 ? 32 / NV : 1;
  const int l_sweep = lane / NV, l_visit = lane - l_sweep * NV;
  int site_l = lane;
  for (int sw = 0; sw < total; ++sw) {
    const double T = P.T_chain ? P.T_chain[chain] : (P.T_sweep ? P.T_sweep[sw] : P.T);
    const double invT = 1.0 / T;
    const int in_batch = sw % batch;
    if (in_batch == 0) {
      const int my_sw = sw + l_sweep;
      if (l_sweep < batch && my_sw < total) {
        site_l = P.order ? P.order[(size_t)my_sw * NV + l_visit] : l_visit;
        double ui;
        if (P.uniforms) {
          ui = P.uniforms[((size_t)my_sw * P.n_chains + chain) * NV + l_visit];
        } else {
          tsu_u32x4 o = tsu_philox4x32_10((uint32_t)site_l, P.chain0 + (uint32_t)chain, P.sweep0 + (uint32_t)my_sw,
                                          TSU_STREAM_DENSE, P.k0, P.k1);
          const unsigned long long m = (((unsigned long long)o.x << 32) | o.y) >> 11;
          ui = (double)m * (1.0 / 9007199254740992.0);
        }
        u[lane] = ui;
        lg[lane] = log(ui) - log1p(-ui);
      }
      __syncwarp();
    }
    const int l0 = in_batch * NV;                        // first lane of this sweep's visits
    for (int idx = 0; idx < NV; ++idx) {
      const int i = __shfl_sync(0xffffffffu, site_l, l0 + idx);
      const double hi = __shfl_sync(0xffffffffu, h, i);
      const int si = (int)((mask >> i) & 1u);
      const double ui = u[l0 + idx], t = lg[l0 + idx];
      const double jv = mine ? Jsm[i * N + lane] : 0.0;  // J[lane][i], fetched before the decision is known
      const double xq = hi * invT;                       // decision exactly as in dense_gibbs_kernel (float64 fields)
      int nb;
      if (xq > 20.000001)
        nb = ui < 1.0 ? 1 : 0;
      else if (xq < -20.000001)
        nb = 0;
      else if (fabs(xq) < 19.999999 && fabs(xq - t) > 1e-4)
        nb = t < xq ? 1 : 0;
      else
        nb = (ui < sigmoid_clamped<double>(hi / T)) ? 1 : 0;
      if (nb != si) {                                    // h_j += J[j][i] * (new - old): +-J exactly, every field
        h += nb ? jv : -jv;                              // incl. the self term J_ii (gibbs.py:97)
        mask ^= 1u << i;
      }
    }
    __syncwarp();                                        // u / lg are rewritten by the next sweep
    if (P.track_best) {                                  // gibbs.py:387-391
      const double e = energy_now();
      if (e < best_e) {
        best_e = e;
        if (mine) P.best_state[(size_t)chain * N + lane] = (uint8_t)((mask >> lane) & 1u);
      }
    }
    if (sw + 1 == next_sample_at) {
      if (P.samples && sample_idx < P.n_samples && mine)
        P.samples[((size_t)sample_idx * P.n_chains + chain) * N + lane] = (uint8_t)((mask >> lane) & 1u);
      ++sample_idx;
      next_sample_at += P.sweeps_per_sample;
    }
  }
  if (mine) gstate[lane] = (uint8_t)((mask >> lane) & 1u);
  if (P.energy) {
    const double e = energy_now();
    if (lane == 0) P.energy[chain] = e;
  }
  if (P.track_best && lane == 0) P.best_energy[chain] = best_e;
}

// E = -1/2 s^T J s - b^T s from scratch, float64 accumulation (tsu/gibbs.py:215-236)
template <typename JT>
__global__ void dense_energy_kernel(const JT* __restrict__ Jt, const JT* __restrict__ bias,
                                    const uint8_t* __restrict__ state, int N, double* energy) {
  __shared__ double red[32];
  const int chain = blockIdx.x;
  const uint8_t* s = state + (size_t)chain * N;
  double acc = 0.0;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    if (!s[j]) continue;
    double hj = 0.0;
    for (int k = 0; k < N; ++k)
      if (s[k]) hj += (double)Jt[(size_t)k * N + j];
    acc += -0.5 * hj - (bias ? (double)bias[j] : 0.0);
  }
  const double e = block_sum(acc, red);
  if (threadIdx.x == 0) energy[chain] = e;
}

__global__ void dense_init_kernel(uint8_t* state, int n_chains, int N, uint32_t k0, uint32_t k1, uint32_t chain0) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 32 sites
  const int wpc = (N + 31) / 32;
  if (t >= (long long)n_chains * wpc) return;
  const int chain = (int)(t / wpc);
  const int w = (int)(t - (long long)chain * wpc);
  tsu_u32x4 o = tsu_philox4x32_10((uint32_t)w, chain0 + (uint32_t)chain, 0u, TSU_STREAM_DENSE_INIT, k0, k1);
  for (int b = 0; b < 32; ++b) {
    const int i = 32 * w + b;
    if (i < N) state[(size_t)chain * N + i] = (uint8_t)((o.x >> b) & 1u);
  }
}

// One thread per ladder walks the pairs in order (gibbs.py:308-323).
__global__ void pt_swap_kernel(const double* __restrict__ energy, const double* __restrict__ T_slot,
                               int32_t* slot_replica, int32_t* lut_index, int n_ladders, int R, uint32_t k0,
                               uint32_t k1, uint32_t step, unsigned long long* stats,
                               const double* __restrict__ uniforms, int criterion) {
  const int ladder = blockIdx.x * blockDim.x + threadIdx.x;
  if (ladder >= n_ladders) return;
  int32_t* sr = slot_replica + (size_t)ladder * R;
  unsigned long long attempts = 0, accepts = 0;
  for (int i = 0; i + 1 < R; ++i) {
    const int ra = sr[i], rb = sr[i + 1];
    const double Ei = energy[ra], Ej = energy[rb];
    // criterion 0: the reference's expression (gibbs.py:317); criterion 1: detailed-balance Metropolis rule,
    // delta = (beta_i - beta_j)(E_i - E_j)
    const double delta = (1.0 / T_slot[i] - 1.0 / T_slot[i + 1]) * (criterion ? (Ei - Ej) : (Ej - Ei));
    ++attempts;
    bool acc = delta >= 0.0;
    if (!acc) {  // the uniform is consumed only when delta < 0 (short-circuit `or`, gibbs.py:320)
      double uu;
      if (uniforms) {
        uu = uniforms[(size_t)ladder * (R - 1) + i];
      } else {
        tsu_u32x4 o = tsu_philox4x32_10((uint32_t)i, (uint32_t)ladder, step, TSU_STREAM_PT_SWAP, k0, k1);
        const unsigned long long m = (((unsigned long long)o.x << 32) | o.y) >> 11;
        uu = (double)m * (1.0 / 9007199254740992.0);
      }
      acc = uu < exp(delta);
    }
    if (acc) {
      sr[i] = rb;
      sr[i + 1] = ra;
      ++accepts;
    }
  }
  if (lut_index)
    for (int i = 0; i < R; ++i) lut_index[sr[i]] = i;
  if (stats) {
    atomicAdd(stats, attempts);
    atomicAdd(stats + 1, accepts);
  }
}

int pick_threads(int N) {
  // fields per thread: a flip adds one row of Jt to the fields and costs about one L2 round trip when every thread has
  // its few loads in flight at once (measured per site visit, 296 chains: N = 512: 2 per thread 490 ns, 4: 615, 16: 1655;
  // N = 2048: 2: 1273, 4: 722, 16: 1640; N = 4096: 4: 1699, 8: 2169, 16: 3360).  TSU_DENSE_FPT overrides.
  static const int forced = getenv("TSU_DENSE_FPT") ? atoi(getenv("TSU_DENSE_FPT")) : 0;
  const int per_thread = forced > 0 ? forced : (N <= 512 ? 2 : 4);
  int t = (N + per_thread - 1) / per_thread;
  t = (t + 31) / 32 * 32;
  if (t < 32) t = 32;
  if (t > 1024) t = 1024;
  return t;
}

template <typename JT, typename AT>
int launch_dense(const DenseParams& P, cudaStream_t st) {
  static const bool no_warp_kernel = getenv("TSU_DENSE_NO_WARP") != nullptr;  // launch shape only: same bits
  if (sizeof(AT) == sizeof(double) && P.N <= kWarpChainMaxN && !no_warp_kernel) {
    const size_t smem = sizeof(double) * ((size_t)P.N * P.N + 64 * kWarpChainsPerCta);
    dense_gibbs_warp_kernel<JT><<<(P.n_chains + kWarpChainsPerCta - 1) / kWarpChainsPerCta, 32 * kWarpChainsPerCta, smem,
                                  st>>>(P);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? TSU_OK : (int)e;
  }
  const size_t smem = sizeof(double) * ((size_t)P.N * (sizeof(AT) == sizeof(double) ? 2 : 1) + 32) +
                      2 * sizeof(AT) * (size_t)P.N + (size_t)P.N + 16;
  if (smem > 227 * 1024) return TSU_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e =
        cudaFuncSetAttribute(dense_gibbs_kernel<JT, AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  dense_gibbs_kernel<JT, AT><<<P.n_chains, pick_threads(P.N), smem, st>>>(P);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TSU_OK : (int)e;
}

}  // namespace

extern "C" {

int tsu_dense_gibbs_run(const void* d_Jt, int j_dtype, const void* d_bias, uint8_t* d_state, int n_chains, int N,
                        double T, const double* d_T_chain, const double* d_T_sweep, int n_burnin, int n_samples,
                        int sweeps_per_sample, const int32_t* d_order, const double* d_uniforms, uint8_t* d_samples,
                        double* d_energy, int track_best, uint8_t* d_best_state, double* d_best_energy, uint64_t seed,
                        uint32_t sweep0, uint32_t chain0, int acc_dtype, int visits_per_sweep, uintptr_t stream) {
  TSU_CHECK_ARG(d_Jt && d_state && n_chains > 0 && N > 0);
  TSU_CHECK_ARG(visits_per_sweep >= 0 && visits_per_sweep <= N && (visits_per_sweep == 0 || d_order));
  TSU_CHECK_ARG(j_dtype == 0 || j_dtype == 1);
  TSU_CHECK_ARG(acc_dtype == 0 || acc_dtype == 1);
  TSU_CHECK_ARG(n_burnin >= 0 && n_samples >= 0 && sweeps_per_sample >= 0);
  TSU_CHECK_ARG(n_samples == 0 || sweeps_per_sample > 0);
  TSU_CHECK_ARG(d_T_chain || d_T_sweep || T > 0);
  TSU_CHECK_ARG(!track_best || (d_best_state && d_best_energy));
  DenseParams P;
  P.Jt = d_Jt;
  P.bias = d_bias;
  P.state = d_state;
  P.T_chain = d_T_chain;
  P.T_sweep = d_T_sweep;
  P.order = d_order;
  P.uniforms = d_uniforms;
  P.samples = d_samples;
  P.energy = d_energy;
  P.best_state = d_best_state;
  P.best_energy = d_best_energy;
  P.T = T;
  P.n_chains = n_chains;
  P.N = N;
  P.n_visit = visits_per_sweep > 0 ? visits_per_sweep : N;
  P.n_burnin = n_burnin;
  P.n_samples = n_samples;
  P.sweeps_per_sample = sweeps_per_sample;
  P.track_best = track_best;
  P.refresh_every = acc_dtype == 0 ? 8 : 0;
  P.k0 = (uint32_t)seed;
  P.k1 = (uint32_t)(seed >> 32);
  P.sweep0 = sweep0;
  P.chain0 = chain0;
  cudaStream_t st = tsu_stream(stream);
  if (j_dtype == 0 && acc_dtype == 0) return launch_dense<float, float>(P, st);
  if (j_dtype == 0 && acc_dtype == 1) return launch_dense<float, double>(P, st);
  if (j_dtype == 1 && acc_dtype == 0) return launch_dense<double, float>(P, st);
  return launch_dense<double, double>(P, st);
}

int tsu_dense_energy(const void* d_Jt, int j_dtype, const void* d_bias, const uint8_t* d_state, int n_chains, int N,
                     double* d_energy, uintptr_t stream) {
  TSU_CHECK_ARG(d_Jt && d_state && d_energy && n_chains > 0 && N > 0);
  TSU_CHECK_ARG(j_dtype == 0 || j_dtype == 1);
  const int threads = pick_threads(N);
  if (j_dtype == 0)
    dense_energy_kernel<float><<<n_chains, threads, 0, tsu_stream(stream)>>>(
        reinterpret_cast<const float*>(d_Jt), reinterpret_cast<const float*>(d_bias), d_state, N, d_energy);
  else
    dense_energy_kernel<double><<<n_chains, threads, 0, tsu_stream(stream)>>>(
        reinterpret_cast<const double*>(d_Jt), reinterpret_cast<const double*>(d_bias), d_state, N, d_energy);
  TSU_RETURN_LAUNCH_STATUS();
}

int tsu_dense_init_random(uint8_t* d_state, int n_chains, int N, uint64_t seed, uint32_t chain0, uintptr_t stream) {
  TSU_CHECK_ARG(d_state && n_chains > 0 && N > 0);
  const long long total = (long long)n_chains * ((N + 31) / 32);
  dense_init_kernel<<<(unsigned)((total + 127) / 128), 128, 0, tsu_stream(stream)>>>(
      d_state, n_chains, N, (uint32_t)seed, (uint32_t)(seed >> 32), chain0);
  TSU_RETURN_LAUNCH_STATUS();
}

int tsu_pt_swap(const double* d_energy, const double* d_T_slot, int32_t* d_slot_replica, int32_t* d_lut_index,
                int n_ladders, int R, uint64_t seed, uint32_t step, unsigned long long* d_stats,
                const double* d_uniforms, int criterion, uintptr_t stream) {
  TSU_CHECK_ARG(d_energy && d_T_slot && d_slot_replica && n_ladders > 0 && R > 0);
  TSU_CHECK_ARG(criterion == 0 || criterion == 1);
  pt_swap_kernel<<<(n_ladders + 63) / 64, 64, 0, tsu_stream(stream)>>>(d_energy, d_T_slot, d_slot_replica,
                                                                      d_lut_index, n_ladders, R, (uint32_t)seed,
                                                                      (uint32_t)(seed >> 32), step, d_stats,
                                                                      d_uniforms, criterion);
  TSU_RETURN_LAUNCH_STATUS();
}

}  // extern "C"

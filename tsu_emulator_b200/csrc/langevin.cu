// Fused overdamped-Langevin sampler for built-in analytic energies (sm_100a).
//
// Replaces the Python loop of ThermalSamplingUnit.sample_from_energy (tsu/core.py:100-162):
//   for every sample: restart at x_init (+ 0.1 N(0,I) for sample_idx > 0, core.py:142-143),
//   n_burnin + n_steps Euler-Maruyama steps (core.py:64-80)
//       x <- x - grad E(x) * dt / gamma + sqrt(2 T dt / gamma) * N(0, I)
//   with the central-difference gradient (core.py:82-98) replaced by the analytic gradient of the
//   built-in energy.  Every sample is an independent chain, so one thread runs one chain with the
//   state in registers: gradient + noise + update are one fused loop, HBM is touched only for the
//   final state (and the optional trajectory).
//
// Noise: Philox4x32-10 + Box-Muller, counter = (chain lo, chain hi | call << 16, step, 'LANG').
// Parity mode reads the N(0,1) draws from a caller tensor so that the reference (with
// numpy.random.randn patched to the same draws) can be compared step for step.

#include "common.cuh"
#include "philox.cuh"

namespace {

enum { ENERGY_QUADRATIC = 0, ENERGY_MIXTURE = 1, ENERGY_DOUBLE_WELL = 2, ENERGY_QUADRATIC_FORM = 3 };

constexpr int kMaxDynDim = 64;

struct LangevinParams {
  void* x;
  const void* x_init;
  const void* normals;
  void* traj;
  const double* params;
  long long n_chains;
  unsigned long long chain0;
  int dim, energy_kind, n_params;
  int n_burnin, n_steps;
  int first_chain_exact;
  double jitter, drift, noise;  // drift = dt / gamma, noise = sqrt(2 T dt / gamma)
  uint32_t k0, k1;
};

__device__ __forceinline__ float lg2_fast(float x) {  // x in [2^-25, 1): no denormal handling needed
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename real>
struct BoxMuller;

template <>
struct BoxMuller<float> {
  static constexpr int kPerCall = 4;
  // 4 normals from one Philox block (24-bit uniforms, exactly representable in float)
  __device__ static __forceinline__ void draw(const tsu_u32x4& o, float z[4]) {
    const float u1 = ((float)(o.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((float)(o.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u3 = ((float)(o.z >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u4 = ((float)(o.w >> 8) + 0.5f) * (1.0f / 16777216.0f);
    // fast-math forms (MUFU lg2 / rsq / sin / cos, absolute error ~2^-21): the float32 kernel is instruction bound and
    // the library logf / sincospif cost three times as many instructions as the rest of the step
    const float a1 = -1.3862943611f * lg2_fast(u1), a2 = -1.3862943611f * lg2_fast(u3);  // -2 ln u > 0 (u < 1)
    // (u rounds to 1.0f for the top uniforms: a = 0 -> r = 0, guarded against 0 * inf)
    const float r1 = a1 * rsqrtf(fmaxf(a1, 1e-30f)), r2 = a2 * rsqrtf(fmaxf(a2, 1e-30f));
    const float t1 = 6.2831853072f * u2, t2 = 6.2831853072f * u4;
    z[0] = r1 * __cosf(t1);
    z[1] = r1 * __sinf(t1);
    z[2] = r2 * __cosf(t2);
    z[3] = r2 * __sinf(t2);
  }
};

template <>
struct BoxMuller<double> {
  static constexpr int kPerCall = 2;
  // 2 normals from one Philox block (53-bit uniforms)
  __device__ static __forceinline__ void draw(const tsu_u32x4& o, double z[2]) {
    const unsigned long long a = (((unsigned long long)o.x << 32) | o.y) >> 11;
    const unsigned long long b = (((unsigned long long)o.z << 32) | o.w) >> 11;
    const double u1 = ((double)a + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)b + 0.5) * (1.0 / 9007199254740992.0);
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z[0] = r * c;
    z[1] = r * s;
  }
};

// N(0,1) vector for (chain, step): injected rows or Philox + Box-Muller
template <typename real, int DIM>
__device__ __forceinline__ void normal_vector(const LangevinParams& P, unsigned long long chain_g, long long chain_l,
                                              int step, int dim, real* z) {
  if (P.normals) {
    const long long rows = 1LL + P.n_burnin + P.n_steps;
    const real* src = reinterpret_cast<const real*>(P.normals) + ((size_t)chain_l * rows + step) * dim;
#pragma unroll
    for (int i = 0; i < (DIM > 0 ? DIM : kMaxDynDim); ++i)
      if (i < dim) z[i] = src[i];
    return;
  }
  constexpr int PER = BoxMuller<real>::kPerCall;
  constexpr int MAXD = DIM > 0 ? DIM : kMaxDynDim;
#pragma unroll
  for (int b = 0; b < (MAXD + PER - 1) / PER; ++b) {
    if (b * PER < dim) {
      tsu_u32x4 o = tsu_philox4x32_10((uint32_t)chain_g, ((uint32_t)(chain_g >> 32) & 0xFFFFu) | ((uint32_t)b << 16),
                                      (uint32_t)step, TSU_STREAM_LANGEVIN, P.k0, P.k1);
      real t[PER];
      BoxMuller<real>::draw(o, t);
#pragma unroll
      for (int q = 0; q < PER; ++q)
        if (b * PER + q < MAXD && b * PER + q < dim) z[b * PER + q] = t[q];
    }
  }
}

// gradient of the built-in energies; sp = parameters staged in shared memory as `real`
template <typename real, int DIM>
__device__ __forceinline__ void gradient(int kind, int dim, const real* __restrict__ sp, const real* x, real* g) {
  constexpr int MAXD = DIM > 0 ? DIM : kMaxDynDim;
  if (kind == ENERGY_QUADRATIC) {
    // E = a sum (x - mu)^2 w ; sp = [a, mu[dim], w[dim]]
    const real two_a = (real)2 * sp[0];
#pragma unroll
    for (int i = 0; i < MAXD; ++i)
      if (i < dim) g[i] = two_a * sp[1 + dim + i] * (x[i] - sp[1 + i]);
  } else if (kind == ENERGY_DOUBLE_WELL) {
    // E = sum a (x^2 - b)^2
    const real a4 = (real)4 * sp[0], b = sp[1];
#pragma unroll
    for (int i = 0; i < MAXD; ++i)
      if (i < dim) g[i] = a4 * x[i] * (x[i] * x[i] - b);
  } else if (kind == ENERGY_QUADRATIC_FORM) {
    // E = 1/2 x^T A x - b^T x ; sp = [A[dim][dim], b[dim]] ; grad = A x - b
#pragma unroll
    for (int i = 0; i < MAXD; ++i)
      if (i < dim) {
        real acc = -sp[dim * dim + i];
#pragma unroll
        for (int j = 0; j < MAXD; ++j)
          if (j < dim) acc += sp[i * dim + j] * x[j];
        g[i] = acc;
      }
  } else {
    // E = -log(sum_k p_k exp(-|x - c_k|^2 / 2) + 1e-10); grad = sum_k p_k e_k (x - c_k) / (sum_k p_k e_k + 1e-10)
    const int K = (int)sp[0];
    real tot = (real)0;
#pragma unroll
    for (int i = 0; i < MAXD; ++i)
      if (i < dim) g[i] = (real)0;
    for (int k = 0; k < K; ++k) {
      const real* c = sp + 1 + K + (size_t)k * dim;
      real d2 = (real)0;
#pragma unroll
      for (int i = 0; i < MAXD; ++i)
        if (i < dim) {
          real d = x[i] - c[i];
          d2 += d * d;
        }
      const real e = sp[1 + k] * (real)exp((real)-0.5 * d2);
      tot += e;
#pragma unroll
      for (int i = 0; i < MAXD; ++i)
        if (i < dim) g[i] += e * (x[i] - c[i]);
    }
    const real inv = (real)1 / (tot + (real)1e-10);
#pragma unroll
    for (int i = 0; i < MAXD; ++i)
      if (i < dim) g[i] *= inv;
  }
}

template <typename real, int DIM>
__global__ void __launch_bounds__(128) langevin_kernel(LangevinParams P) {
  extern __shared__ double smem_raw[];
  real* sp = reinterpret_cast<real*>(smem_raw);
  for (int i = threadIdx.x; i < P.n_params; i += blockDim.x) sp[i] = (real)P.params[i];
  __syncthreads();

  const long long chain = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= P.n_chains) return;
  const unsigned long long chain_g = P.chain0 + (unsigned long long)chain;
  const int dim = DIM > 0 ? DIM : P.dim;
  constexpr int MAXD = DIM > 0 ? DIM : kMaxDynDim;
  real x[MAXD], z[MAXD], g[MAXD];

  const real* xi = reinterpret_cast<const real*>(P.x_init);
#pragma unroll
  for (int i = 0; i < MAXD; ++i)
    if (i < dim) x[i] = xi ? xi[i] : (real)0;
  // every chain but the first of a call starts at x_init + jitter * N(0, I)  (core.py:142-143)
  if (!(chain == 0 && P.first_chain_exact) && P.jitter != 0.0) {
    normal_vector<real, DIM>(P, chain_g, chain, 0, dim, z);
#pragma unroll
    for (int i = 0; i < MAXD; ++i)
      if (i < dim) x[i] = x[i] + (real)P.jitter * z[i];
  }
  const real drift = (real)P.drift, noise = (real)P.noise;
  const int total = P.n_burnin + P.n_steps;
  real* traj = reinterpret_cast<real*>(P.traj);
  for (int s = 0; s < total; ++s) {
    gradient<real, DIM>(P.energy_kind, dim, sp, x, g);
    normal_vector<real, DIM>(P, chain_g, chain, s + 1, dim, z);
#pragma unroll
    for (int i = 0; i < MAXD; ++i)
      if (i < dim) x[i] = x[i] + (-g[i] * drift) + noise * z[i];  // core.py:74-80 order of operations
    if (traj && s >= P.n_burnin) {
      real* dst = traj + ((size_t)chain * P.n_steps + (s - P.n_burnin)) * dim;
#pragma unroll
      for (int i = 0; i < MAXD; ++i)
        if (i < dim) dst[i] = x[i];
    }
  }
  real* out = reinterpret_cast<real*>(P.x) + (size_t)chain * dim;
#pragma unroll
  for (int i = 0; i < MAXD; ++i)
    if (i < dim) out[i] = x[i];
}

template <typename real, int DIM>
int launch(const LangevinParams& P, cudaStream_t st) {
  const unsigned grid = (unsigned)((P.n_chains + 127) / 128);
  const size_t smem = sizeof(double) * (size_t)((P.n_params + 1) & ~1);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(langevin_kernel<real, DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  langevin_kernel<real, DIM><<<grid, 128, smem, st>>>(P);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TSU_OK : (int)e;
}

template <typename real>
int dispatch_dim(const LangevinParams& P, cudaStream_t st) {
  switch (P.dim) {
    case 1: return launch<real, 1>(P, st);
    case 2: return launch<real, 2>(P, st);
    case 3: return launch<real, 3>(P, st);
    case 4: return launch<real, 4>(P, st);
    case 8: return launch<real, 8>(P, st);
    case 10: return launch<real, 10>(P, st);
    case 16: return launch<real, 16>(P, st);
    default: return launch<real, 0>(P, st);
  }
}

}  // namespace

extern "C" int tsu_langevin_run(void* d_x, int dtype, int64_t n_chains, int dim, int energy_kind,
                                const double* d_params, int n_params, const void* d_x_init, double jitter,
                                int first_chain_exact, double T, double dt, double gamma, int n_burnin, int n_steps, uint64_t seed, uint64_t chain0,
                                const void* d_normals, void* d_traj, uintptr_t stream) {
  TSU_CHECK_ARG(d_x && d_params && n_chains > 0 && dim > 0 && dim <= kMaxDynDim);
  TSU_CHECK_ARG(dtype == 0 || dtype == 1);
  TSU_CHECK_ARG(T > 0 && dt > 0 && gamma > 0 && n_burnin >= 0 && n_steps >= 0);
  TSU_CHECK_ARG(energy_kind >= 0 && energy_kind <= 3);
  if (energy_kind == ENERGY_QUADRATIC_FORM) TSU_CHECK_ARG(n_params == dim * dim + dim);
  if (energy_kind == ENERGY_QUADRATIC) TSU_CHECK_ARG(n_params == 1 + 2 * dim);
  if (energy_kind == ENERGY_DOUBLE_WELL) TSU_CHECK_ARG(n_params == 2);
  if (energy_kind == ENERGY_MIXTURE) TSU_CHECK_ARG(n_params >= 2 + dim && n_params <= 20000);
  LangevinParams P;
  P.x = d_x;
  P.x_init = d_x_init;
  P.normals = d_normals;
  P.traj = d_traj;
  P.params = d_params;
  P.n_chains = n_chains;
  P.chain0 = chain0;
  P.dim = dim;
  P.energy_kind = energy_kind;
  P.n_params = n_params;
  P.n_burnin = n_burnin;
  P.n_steps = n_steps;
  P.jitter = jitter;
  P.first_chain_exact = first_chain_exact;
  P.drift = dt / gamma;
  P.noise = sqrt(2.0 * T * dt / gamma);
  P.k0 = (uint32_t)seed;
  P.k1 = (uint32_t)(seed >> 32);
  return dtype == 0 ? dispatch_dim<float>(P, tsu_stream(stream)) : dispatch_dim<double>(P, tsu_stream(stream));
}

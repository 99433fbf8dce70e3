// Counter-based Philox4x32-10 (Salmon et al., SC'11), device + host.
//
// Replaces the reference's global NumPy MT19937 stream (tsu/gibbs.py:126,157,201,270,320,368;
// tsu/core.py:78,143).  Every random number in this library is a pure function of
// (seed, stream coordinates), so results do not depend on launch geometry or on how a
// lattice is split across GPUs.
//
// One call = 10 rounds of { 2 x IMAD.WIDE.U32, 2 x LOP3 } : 20 fma-pipe + 20 alu-pipe
// instructions for 128 random bits, all in registers.
#pragma once
#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

#define TSU_PHILOX_M0 0xD2511F53u
#define TSU_PHILOX_M1 0xCD9E8D57u
#define TSU_PHILOX_W0 0x9E3779B9u
#define TSU_PHILOX_W1 0xBB67AE85u

struct tsu_u32x4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ void tsu_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
  lo = a * b;
  hi = __umulhi(a, b);
#else
  uint64_t p = (uint64_t)a * (uint64_t)b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}

__host__ __device__ __forceinline__ tsu_u32x4 tsu_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    tsu_mulhilo(TSU_PHILOX_M0, c0, hi0, lo0);
    tsu_mulhilo(TSU_PHILOX_M1, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += TSU_PHILOX_W0;
    k1 += TSU_PHILOX_W1;
  }
  tsu_u32x4 o;
  o.x = c0;
  o.y = c1;
  o.z = c2;
  o.w = c3;
  return o;
}

// ---- stream coordinates ---------------------------------------------------------------
// Lattice stream (ising2d.cu): counter = (w | colour<<20 | kind<<21, global_row, sweep, replica)
//   kind 0,1  : bit-plane calls (planes 0-3 / 4-7 = top 8 bits of every lane's uniform)
//   kind 2    : initial configuration
//   kind 8-15 : low 24 bits, one call per group of 4 lanes (lane b uses word b&3 of call 8+(b>>2))
#define TSU_KIND_PLANE0 0u
#define TSU_KIND_PLANE1 1u
#define TSU_KIND_INIT 2u
#define TSU_KIND_LOW0 8u
#define TSU_LATTICE_C0(w, colour, kind) ((uint32_t)(w) | ((uint32_t)(colour) << 20) | ((uint32_t)(kind) << 21))

// Generic streams (dense Gibbs / Langevin / fill): counter = (index_lo, index_hi, step, stream_tag)
#define TSU_STREAM_FILL 0x46494C4Cu      // 'FILL'
#define TSU_STREAM_DENSE 0x44454E53u     // 'DENS'
#define TSU_STREAM_DENSE_TC 0x44454E54u  // 'DENT': tensor-core dense path, counter = (site>>2, chain, sweep, tag)
#define TSU_STREAM_DENSE_INIT 0x44494E49u  // 'DINI'
#define TSU_STREAM_LANGEVIN 0x4C414E47u  // 'LANG'
#define TSU_STREAM_LANGEVIN_INIT 0x4C494E49u  // 'LINI'
#define TSU_STREAM_PT_SWAP 0x50545357u   // 'PTSW'

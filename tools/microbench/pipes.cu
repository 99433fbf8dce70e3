// Issue-rate microbenchmark for the integer instructions the lattice kernel is made of (B200, sm_100a):
// which pairs of pipes overlap?  nvcc -arch=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Every kernel runs ILP independent dependency chains per thread, 8 warps per SM sub-partition; the result is
// warp instructions per cycle per SM sub-partition (1.0 = the issue limit).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
  uint32_t x[ILP], y[ILP];
  unsigned long long p[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { x[i] = seed + threadIdx.x * 7 + i; y[i] = seed ^ (i * 0x9e3779b9u); p[i] = x[i]; }
  const uint32_t z = seed * 3 + 1;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) {  // LOP3, three register sources
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(z));
      } else if (MODE == 1) {  // IMAD.WIDE chain (64-bit accumulate), immediate multiplier
        asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(p[i]) : "r"(y[i]));
      } else if (MODE == 2) {  // LOP3 + IMAD.WIDE, independent
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(z));
        asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(p[i]) : "r"(y[i]));
      } else if (MODE == 3) {  // IMAD 32-bit
        asm volatile("mad.lo.u32 %0, %0, 0xD2511F53, %1;" : "+r"(x[i]) : "r"(y[i]));
      } else if (MODE == 4) {  // LOP3 + IMAD 32-bit
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(z));
        asm volatile("mad.lo.u32 %0, %0, 0xD2511F53, %1;" : "+r"(y[i]) : "r"(z));
      } else if (MODE == 5) {  // LOP3, two register sources + immediate
        asm volatile("lop3.b32 %0, %0, %1, 0x9e3779b9, 0x96;" : "+r"(x[i]) : "r"(y[i]));
      } else if (MODE == 6) {  // Philox-like round: wide multiply feeding a 3-input xor (dependent)
        unsigned long long q;
        asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(q) : "r"(x[i]));
        uint32_t lo = (uint32_t)q, hi = (uint32_t)(q >> 32);
        asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[i]) : "r"(hi), "r"(y[i]), "r"(z));
        y[i] = lo;
      } else if (MODE == 7) {  // mul.hi + mul.lo instead of mul.wide
        uint32_t lo, hi;
        asm volatile("mul.hi.u32 %0, %1, 0xD2511F53;" : "=r"(hi) : "r"(x[i]));
        asm volatile("mul.lo.u32 %0, %1, 0xD2511F53;" : "=r"(lo) : "r"(x[i]));
        asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[i]) : "r"(hi), "r"(y[i]), "r"(z));
        y[i] = lo;
      } else if (MODE == 8) {  // FFMA (fp32 pipe) + LOP3
        float f = __uint_as_float(y[i]);
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f) : "f"(1.0001f));
        y[i] = __float_as_uint(f);
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(z), "r"(z));
      } else if (MODE == 9) {  // shifts (SHF) + LOP3
        asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(y[i]) : "r"(z));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(z), "r"(z));
      } else if (MODE == 10) {  // IADD3
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
      } else if (MODE == 11) {  // LOP3 with a constant-bank operand (kernel parameter)
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(seed));
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc ^= x[i] ^ y[i] ^ (uint32_t)p[i] ^ (uint32_t)(p[i] >> 32);
  if (acc == 0x12345678u) out[0] = acc;
}

template <int MODE>
void run(const char* name, double inst_per_iter, uint32_t* d) {
  const int blocks = 148 * 8;  // 8 CTAs of 256 threads per SM = 64 warps = 16 per sub-partition
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<blocks, 256>>>(d, 1u);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<MODE><<<blocks, 256>>>(d, 1u);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double warp_inst = (double)blocks * 8 * ITERS * ILP * inst_per_iter;
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-52s %8.3f ms  %.3f warp-inst/cycle/SMSP (at %d MHz)\n", name, ms, warp_inst / cycles / (148 * 4), clk / 1000);
}

int main() {
  uint32_t* d; cudaMalloc(&d, 4);
  run<0>("LOP3 (3 regs)", 1, d);
  run<5>("LOP3 (2 regs + imm)", 1, d);
  run<11>("LOP3 (2 regs + const bank)", 1, d);
  run<10>("IADD", 1, d);
  run<1>("IMAD.WIDE (imm multiplier)", 1, d);
  run<3>("IMAD 32-bit (imm multiplier)", 1, d);
  run<2>("LOP3 + IMAD.WIDE independent (2 inst)", 2, d);
  run<4>("LOP3 + IMAD 32 independent (2 inst)", 2, d);
  run<6>("mul.wide -> xor3 (Philox half round, 2 inst)", 2, d);
  run<7>("mul.hi + mul.lo -> xor3 (3 inst)", 3, d);
  run<8>("FFMA + LOP3 (2 inst)", 2, d);
  run<9>("SHF + LOP3 (2 inst)", 2, d);
  return 0;
}

// Micro-benchmark: issue rate of scalar FFMA (constant / 3-register operands), packed FFMA2
// (fma.rn.f32x2, sm_100+), and mixes with ALU-pipe instructions (FMNMX).  B200, 148 SMs.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
  float ra = a + threadIdx.x * 1e-9f, rb = b + threadIdx.x * 1e-9f;   // register operands
  uint64_t p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = pack(x[2 * i], x[2 * i + 1]);
  const uint64_t pa = pack(ra, ra), pb = pack(rb, rb);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (MODE == 0) {          // scalar FFMA, constant-bank operands: 16 per round
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = x[i] * a + b;
      } else if (MODE == 1) {   // scalar FFMA, three register operands
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(ra), "f"(rb));
      } else if (MODE == 2) {   // packed FFMA2: 8 per round = 16 FMAs
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], pa, pb);
      } else if (MODE == 3) {   // 16 scalar FFMA + 8 FMNMX
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(ra), "f"(rb));
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(rb));
      } else if (MODE == 4) {   // 8 FFMA2 + 8 FMNMX (the same arithmetic as mode 3)
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], pa, pb);
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(rb));
      } else if (MODE == 5) {   // 16 FMNMX only (ALU pipe rate)
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(rb));
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)(p[i] & 0xffffffffu)) + __uint_as_float((unsigned)(p[i] >> 32));
  if (s == -1.2345f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double fma_per_round, double instr_per_round) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000, grid = 148 * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<MODE><<<grid, 256>>>(out, 100, 1.0000001f, 1e-9f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  bench<MODE><<<grid, 256>>>(out, iters, 1.0000001f, 1e-9f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double threads = (double)grid * 256, rounds = (double)iters * 4;
  const double warp_instr_per_s = threads / 32 * rounds * instr_per_round / (ms * 1e-3);
  printf("%-34s %7.2f ms  %6.1f TFLOP/s  %.3f warp-instr/clk/SMSP (at 1965 MHz)\n", name, ms,
         threads * rounds * fma_per_round * 2 / (ms * 1e-3) / 1e12, warp_instr_per_s / (148.0 * 4 * 1.965e9));
  cudaFree(out);
}
int main() {
  run<0>("FFMA const operands", 16, 16);
  run<1>("FFMA 3 registers", 16, 16);
  run<2>("FFMA2 (f32x2)", 16, 8);
  run<3>("16 FFMA + 8 FMNMX", 16, 24);
  run<4>("8 FFMA2 + 8 FMNMX", 16, 16);
  run<5>("FMNMX only", 0, 16);
  return 0;
}

// Throughput of FFMA2 by operand shape on sm_100a: accumulator + two fixed pairs, three distinct
// varying register pairs, scalar-broadcast operands.  8 independent chains, 8 warps per SMSP.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  uint64_t p[8], q[8], r[8];
  float sa[8], sb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    p[i] = pack(1.f + i + threadIdx.x, 2.f + i);
    q[i] = pack(1.0000001f + i * 1e-9f, 1.0000002f + threadIdx.x * 1e-9f);
    r[i] = pack(1e-9f * (i + 1), 2e-9f * (i + 1) + threadIdx.x * 1e-12f);
    sa[i] = a + i * 1e-9f + threadIdx.x * 1e-10f; sb[i] = b + i * 1e-10f;
  }
  const uint64_t pa = pack(a + threadIdx.x * 1e-9f, a), pb = pack(b, b + threadIdx.x * 1e-10f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
        if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(q[i]), "l"(r[i]));
        if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(q[(i + rr) & 7]), "l"(r[(i + 2 * rr + 1) & 7]));
        if (MODE == 3) {   // scalar-broadcast operands like the kernel's per-step scalars
          uint64_t s1 = pack(sa[i], sa[i]), s2 = pack(sb[i], sb[i]);
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(s1), "l"(s2));
        }
        if (MODE == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(r[(i + rr) & 7]));
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  if (s == -1.2345f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000, grid = 148 * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(out, 100, 1.0000001f, 1e-9f);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, iters, 1.0000001f, 1e-9f);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double wi = (double)grid * 8 * iters * 64.0 / (ms * 1e-3);
  printf("%-44s %.3f packed instr/clk/SMSP\n", name, wi / (148.0 * 4 * 1.965e9));
  cudaFree(out);
}
int main() {
  run<0>("FFMA2 acc, fixed pair, fixed pair");
  run<1>("FFMA2 acc, varying pair, varying pair");
  run<2>("FFMA2 acc, rotating pairs");
  run<3>("FFMA2 acc, scalar broadcast x2");
  run<4>("FADD2 acc, rotating pair");
  return 0;
}

// Dependent-issue latency of scalar FFMA and packed FFMA2 / FADD2 / FMUL2 on sm_100a: one warp per SM
// runs a serial chain; cycles per instruction = latency.  Also chains of ILP 2 and 4.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
template <int MODE, int ILP>
__global__ void chain(float* out, long long* cyc, int iters, float a, float b) {
  float x[4] = {1.f + threadIdx.x, 2.f, 3.f, 4.f};
  uint64_t p[4] = {pack(1.f, 2.f), pack(3.f, 4.f), pack(5.f, 6.f), pack(7.f, 8.f)};
  const float ra = a + threadIdx.x * 1e-9f, rb = b;
  const uint64_t pa = pack(ra, ra), pb = pack(rb, rb);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(ra), "f"(rb));
        if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
        if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
        if (MODE == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
        if (MODE == 4) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(rb));
      }
    }
  }
  const long long t1 = clock64();
  float s = x[0] + x[1] + x[2] + x[3];
  for (int i = 0; i < 4; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (s == -1.2345f) out[threadIdx.x] = s;
}
template <int MODE, int ILP>
void run(const char* name) {
  float* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  chain<MODE, ILP><<<1, 32>>>(out, cyc, 10, 1.0000001f, 1e-9f);
  chain<MODE, ILP><<<1, 32>>>(out, cyc, iters, 1.0000001f, 1e-9f);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-10s ILP %d: %.2f cycles per instruction (%.2f per chain step)\n", name, ILP, (double)h / (iters * 16.0 * ILP), (double)h / (iters * 16.0));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 1>("FFMA"); run<0, 2>("FFMA"); run<0, 4>("FFMA");
  run<1, 1>("FFMA2"); run<1, 2>("FFMA2"); run<1, 4>("FFMA2");
  run<2, 1>("FADD2"); run<2, 4>("FADD2");
  run<3, 1>("FMUL2"); run<3, 4>("FMUL2");
  run<4, 1>("FMNMX"); run<4, 4>("FMNMX");
  return 0;
}

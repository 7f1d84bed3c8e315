#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
namespace cde = cuda::device::experimental;
using barrier = cuda::barrier<cuda::thread_scope_block>;
__global__ void k(const __grid_constant__ CUtensorMap tmap, float* out, int x, int y) {
  __shared__ alignas(128) float win[24][48];
  #pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&win, &tmap, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(win));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < 48*24; i += blockDim.x) out[i] = ((float*)win)[i];
}
int main(int argc, char** argv) {
  const int W = 256, H = 128;
  std::vector<float> h(W*H);
  for (int i = 0; i < W*H; ++i) h[i] = (float)i;
  float *d, *o; cudaMalloc(&d, W*H*4); cudaMalloc(&o, 48*24*4);
  cudaMemcpy(d, h.data(), W*H*4, cudaMemcpyHostToDevice);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q);
  printf("entry %d %d %p\n", (int)e, (int)q, fn);
  cuuint64_t dims[2] = {W, H}; cuuint64_t strides[1] = {W*4}; cuuint32_t box[2] = {48, 24}; cuuint32_t es[2] = {1,1};
  CUtensorMap m;
  CUresult r = ((Enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  for (int i = 0; i < 16; ++i) printf("%016llx ", (unsigned long long)m.opaque[i]); printf("\n");
  k<<<1, 64>>>(m, o, 10, 7);
  e = cudaDeviceSynchronize(); printf("sync %s\n", cudaGetErrorString(e));
  std::vector<float> out(48*24); cudaMemcpy(out.data(), o, 48*24*4, cudaMemcpyDeviceToHost);
  printf("out[0]=%g expect %g ; out[49]=%g expect %g\n", out[0], (float)(7*W+10), out[49], (float)(8*W+11));
  return 0;
}

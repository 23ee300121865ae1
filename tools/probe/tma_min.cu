// tools/probe/tma_min.cu — smallest TMA round trip with the exact PTX helpers of kernels_post.cu (debug aid):
// load a [32 rows x 384 B] tile of a 3-D u32 tensor into shared memory, copy it out with a TMA store, compare.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define BSB_D __device__ __forceinline__
namespace tma {
BSB_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
BSB_D void mbar_init(uint64_t* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
BSB_D void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
BSB_D void mbar_expect_tx(uint64_t* bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
BSB_D bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
BSB_D void mbar_wait(uint64_t* bar, unsigned parity) { while (!mbar_try_wait(bar, parity)) {} }
BSB_D void load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
BSB_D void store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" :: "l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
BSB_D void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
BSB_D void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
BSB_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}
struct Maps { CUtensorMap in, out, small; };
__global__ void k(const __grid_constant__ Maps tm, int mode) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
  if (threadIdx.x == 0) { tma::mbar_init(bar, 1); tma::fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (mode == 0) { tma::mbar_expect_tx(bar, 12288); tma::load_3d(smem, &tm.in, blockIdx.x * 96, blockIdx.y * 32, blockIdx.z, bar); }
    else { tma::mbar_expect_tx(bar, 1920); tma::load_3d(smem, &tm.small, 16 * blockIdx.x, 1, blockIdx.z, bar); }
  }
  tma::mbar_wait(bar, 0);
  if (mode == 0) {
    reinterpret_cast<unsigned*>(smem)[threadIdx.x] ^= 0u;       // generic-proxy touch
    tma::fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) { tma::store_3d(&tm.out, smem, blockIdx.x * 96, blockIdx.y * 32, blockIdx.z); tma::store_commit(); tma::store_wait_read(); }
  } else if (threadIdx.x == 0 && blockIdx.x == 0) printf("small patch first bytes: %d %d %d\n", smem[0], smem[1], smem[80]);
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("entry point: err %d q %d p %p\n", (int)e, (int)q, p);
  Enc enc = (Enc)p;
  const int W = 640, H = 480, B = 2;
  uint8_t *din, *dout, *dsmall;
  cudaMalloc(&din, (size_t)W * 3 * H * B); cudaMalloc(&dout, (size_t)W * 3 * H * B); cudaMalloc(&dsmall, 256 * 256 * B);
  std::vector<uint8_t> h((size_t)W * 3 * H * B); for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 9));
  cudaMemcpy(din, h.data(), h.size(), cudaMemcpyHostToDevice); cudaMemset(dout, 0, h.size());
  std::vector<uint8_t> hs(256 * 256 * B); for (size_t i = 0; i < hs.size(); ++i) hs[i] = (uint8_t)(i & 255);
  cudaMemcpy(dsmall, hs.data(), hs.size(), cudaMemcpyHostToDevice);
  Maps tm;
  cuuint64_t gd[3] = {W * 3 / 4, H, B}, gs[2] = {(cuuint64_t)W * 3, (cuuint64_t)W * 3 * H}; cuuint32_t bx[3] = {96, 32, 1}, es[3] = {1, 1, 1};
  CUresult r1 = enc(&tm.in, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, din, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&tm.out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, dout, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cuuint64_t gd2[3] = {256, 256, B}, gs2[2] = {256, 65536}; cuuint32_t bx2[3] = {80, 24, 1};
  CUresult r3 = enc(&tm.small, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, dsmall, gd2, gs2, bx2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d %d %d\n", (int)r1, (int)r2, (int)r3);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<<<dim3(5, 15, B), 256, 65536>>>(tm, 0);
  e = cudaDeviceSynchronize(); printf("big tile kernel: %s\n", cudaGetErrorString(e));
  std::vector<uint8_t> o(h.size()); cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost);
  size_t bad = 0; for (size_t i = 0; i < o.size(); ++i) bad += o[i] != h[i];
  printf("mismatches: %zu of %zu\n", bad, o.size());
  k<<<dim3(2, 1, B), 256, 65536>>>(tm, 1);
  e = cudaDeviceSynchronize(); printf("small patch kernel: %s (expect bytes %d %d %d)\n", cudaGetErrorString(e), hs[256], hs[256 + 1], hs[512]);
  return 0;
}

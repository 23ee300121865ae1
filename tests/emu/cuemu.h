// tests/emu/cuemu.h — a tiny CUDA *kernel-logic emulator* for the GPU-less build box.
//
// TEST INFRASTRUCTURE ONLY.  The product library (backscrub_b200/libbackscrub_b200.so)
// is compiled by nvcc for sm_100a and has no CPU path.  This header lets the very same
// .cu sources be compiled by g++ with -DBSB_EMU into tests/emu/libbsb_emu.so, where every
// __global__ kernel runs block by block on the host with CUDA thread semantics
// (threadIdx/blockIdx, __shared__, __syncthreads, warp shuffles via fibers).  It exists so
// that `pytest -m "not gpu"` can check indexing / fusion / planner logic of the CUDA
// kernels against the oracle before GPU minutes are spent.  It is never loaded by the
// backscrub_b200 package, bench.py or smoke(), and nothing it computes is ever reported
// as a result of the product.
#pragma once
#ifndef BSB_EMU
#error "cuemu.h is only for the -DBSB_EMU test build"
#endif

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

// ---- qualifiers -----------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static
#define __restrict__ __restrict
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types -----------------------------------------------------------
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
struct uint3 { unsigned x, y, z; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct uchar3 { unsigned char x, y, z; };
struct uchar4 { unsigned char x, y, z, w; };
struct ushort2 { unsigned short x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }
static inline uchar4 make_uchar4(unsigned char a, unsigned char b, unsigned char c, unsigned char d) { return uchar4{a, b, c, d}; }

// ---- runtime state ----------------------------------------------------------
namespace cuemu {
extern thread_local uint3 t_threadIdx, t_blockIdx;
extern thread_local dim3 t_blockDim, t_gridDim;
unsigned char* dyn_smem();
void syncthreads();
int syncthreads_and(int pred);
void syncwarp();
unsigned shfl_exchange(unsigned value, int src_lane_or_delta, int mode, int width);  // mode 0 idx,1 down,2 up,3 xor
unsigned ballot(int pred);
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body, const void* key);
long launches();
}  // namespace cuemu
#define threadIdx (cuemu::t_threadIdx)
#define blockIdx (cuemu::t_blockIdx)
#define blockDim (cuemu::t_blockDim)
#define gridDim (cuemu::t_gridDim)
static const int warpSize = 32;

// ---- device intrinsics --------------------------------------------------------
static inline void __syncthreads() { cuemu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cuemu::syncwarp(); }
static inline int __syncthreads_and(int pred) { return cuemu::syncthreads_and(pred); }
static inline int __syncthreads_or(int pred) { return !cuemu::syncthreads_and(!pred); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int lane, int width = 32) {
  unsigned u; static_assert(sizeof(T) == 4, "32-bit shuffles only"); memcpy(&u, &v, 4);
  u = cuemu::shfl_exchange(u, lane, 0, width); T r; memcpy(&r, &u, 4); return r; }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  unsigned u; memcpy(&u, &v, 4); u = cuemu::shfl_exchange(u, (int)d, 1, width); T r; memcpy(&r, &u, 4); return r; }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  unsigned u; memcpy(&u, &v, 4); u = cuemu::shfl_exchange(u, (int)d, 2, width); T r; memcpy(&r, &u, 4); return r; }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  unsigned u; memcpy(&u, &v, 4); u = cuemu::shfl_exchange(u, m, 3, width); T r; memcpy(&r, &u, 4); return r; }
static inline unsigned __ballot_sync(unsigned, int pred) { return cuemu::ballot(pred); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline int __float2int_rn(float a) { return (int)lrintf(a); }
static inline int __float2int_rd(float a) { return (int)floorf(a); }
static inline float __int2float_rn(int a) { return (float)a; }
static inline float __uint2float_rn(unsigned a) { return (float)a; }
static inline int __float_as_int(float a) { int r; memcpy(&r, &a, 4); return r; }
static inline float __int_as_float(int a) { float r; memcpy(&r, &a, 4); return r; }
static inline unsigned __float_as_uint(float a) { unsigned r; memcpy(&r, &a, 4); return r; }
static inline float __uint_as_float(unsigned a) { float r; memcpy(&r, &a, 4); return r; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline int __popc(unsigned a) { return __builtin_popcount(a); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  uint64_t v = ((uint64_t)b << 32) | a; unsigned r = 0;
  for (int i = 0; i < 4; ++i) { unsigned sel = (s >> (4 * i)) & 0xf; unsigned byte = (unsigned)(v >> (8 * (sel & 7))) & 0xff;
    if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00; r |= byte << (8 * i); }
  return r; }
static inline int __dp4a(unsigned a, unsigned b, int c) {  // unsigned x unsigned bytes
  for (int i = 0; i < 4; ++i) c += (int)((a >> (8 * i)) & 0xff) * (int)((b >> (8 * i)) & 0xff); return c; }
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
  for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff); return c; }
static inline unsigned __dp2a_lo(unsigned a, unsigned b, unsigned c) { return c + (a & 0xffffu) * (b & 0xffu) + (a >> 16) * ((b >> 8) & 0xffu); }
static inline unsigned __dp2a_hi(unsigned a, unsigned b, unsigned c) { return c + (a & 0xffffu) * ((b >> 16) & 0xffu) + (a >> 16) * ((b >> 24) & 0xffu); }
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline void __threadfence() {}
static inline int __vimin_s32_relu(int a, int b) { int m = a < b ? a : b; return m < 0 ? 0 : m; }
static inline int __viaddmax_s32(int a, int b, int c) { int s = a + b; return s > c ? s : c; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline float fminf_(float a, float b) { return fminf(a, b); }

// ---- host runtime API subset (device memory == host memory) -------------------
typedef int cudaError_t;
typedef struct cuemu_stream* cudaStream_t;
typedef struct cuemu_event* cudaEvent_t;
typedef struct cuemu_graph* cudaGraph_t;
typedef struct cuemu_graphexec* cudaGraphExec_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaStreamCaptureModeThreadLocal = 1, cudaHostAllocDefault = 0 };
static inline const char* cudaGetErrorString(cudaError_t) { return "cuemu"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
  for (size_t y = 0; y < h; ++y) { memmove((char*)d + y * dp, (const char*)s + y * sp, w); }
  return 0; }
static inline cudaError_t cudaMemcpy2D(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind k) {
  return cudaMemcpy2DAsync(d, dp, s, sp, w, h, k, nullptr); }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }

#define BSB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  cuemu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); }, (const void*)#kernel)
#define BSB_DYN_SMEM(name) unsigned char* name = cuemu::dyn_smem()

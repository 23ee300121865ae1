// backscrub_b200/csrc/kernels_tc.cu — tensor-core pointwise (1x1) convolution for sm_100a.
//
// The 1x1 convs are the only dense GEMMs on the path (out[M][N] = A[M][K] * W[N][K]^T, M = frames*H*W).
// This kernel runs them on the 5th-generation tensor cores: tcgen05.mma kind::tf32, accumulators in
// TMEM, operands in 128-byte-swizzled shared memory, completion through mbarriers.  Single-pass TF32
// breaks the mask-parity bar, so every product is the 3xTF32 split
//     a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo,   x_hi = tf32(x) (hardware truncation), x_lo = x - x_hi,
// three MMAs per K step into the same accumulator (the dropped a_lo*w_lo term is ~2^-22 relative).
// It is NOT bit-identical to the FFMA path / the oracle (different summation order); it is selected only
// with BSB_FLAG_TENSOR_CORES and is validated on output tolerance + decision agreement.
//
// Shared-memory operand layout (K-major, SWIZZLE_128B): one K chunk = 32 fp32 = one 128-byte row per
// M/N index; 8 rows form a 1024-byte swizzle atom in which the 16-byte column c of row r is stored at
// column c ^ (r & 7).  Descriptor: start >> 4, LBO = 1, SBO = 1024 >> 4, version 1, layout 2.
#include "kernels.h"

namespace bsb {

void count_launch();

#ifndef BSB_EMU

struct EpiDevTc { const float* bias; const float* residual; int ld_res; int act1, act2, act3; };

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);          // start address, bits [0,14)
  d |= (uint64_t)1u << 16;                           // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024u >> 4) << 32;                 // stride byte offset between 8-row groups, bits [32,46)
  d |= (uint64_t)1u << 46;                           // descriptor version 1 (Blackwell), bits [46,48)
  d |= (uint64_t)2u << 61;                           // layout type SWIZZLE_128B, bits [61,64)
  return d;
}

__device__ __forceinline__ uint32_t swz(int row, int c16) {       // byte offset of 16-byte column c16 of row `row`
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((c16 ^ (row & 7)) << 4));
}

}  // namespace tc

struct TcArgs {
  const float* A; const float* w_hi; const float* w_lo; float* out;
  int M, K, N, ld_a, ld_out, kpad;      // w_*: [npad][kpad], zero padded; kpad % 32 == 0, npad % BN == 0
  EpiDevTc e;
};

__device__ __forceinline__ float tc_epilogue(float total, int ch, size_t pix, const EpiDevTc& e) {
  float v = total + (e.bias ? __ldg(e.bias + ch) : 0.f);
  v = bsb_act(v, e.act1);
  v = bsb_act(v, e.act2);
  if (e.residual) v = bsb_act(v + __ldg(e.residual + pix * (size_t)e.ld_res + ch), e.act3);
  return v;
}

// One CTA = one 128 x BN output tile (BN <= 128).  256 threads.  Three shared-memory stages fed by
// cp.async (16-byte copies land directly at their swizzled position, two K chunks in flight ahead of the
// tensor core); per chunk the threads split the A tile in place into hi / lo, thread 0 issues the twelve
// tcgen05.mma (3 products x 4 K steps) and commits to the stage's mbarrier.  All 8 warps drain TMEM
// (warp w reads lane quadrant w % 4, column half w / 4) with 16-byte stores.
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g, bool valid) {
  const int sz = valid ? 16 : 0;                      // src-size 0 => 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(saddr), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int BN>
__global__ void __launch_bounds__(256, 1) k_pointwise_tc(TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  constexpr int NS = 3;
  constexpr int A_BYTES = 128 * 128;            // 128 rows x 32 fp32
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : 128);
  uint8_t* base = smem_dyn;                      // 1024-byte aligned by declaration (SWIZZLE_128B atoms need it)
  __shared__ __align__(8) uint64_t bar_free[NS];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * BN;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) tc::mbar_init(&bar_free[i], 1);
    tc::mbar_init(&bar_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int nchunks = a.kpad / 32;
  const uint32_t sbase = tc::smem_u32(base);

  auto issue_loads = [&](int c) {           // global -> shared for K chunk c (A raw, W hi, W lo)
    const uint32_t st = sbase + (uint32_t)((c % NS) * STAGE);
    const int k0 = c * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = tid + 256 * i, row = q >> 3, c16 = q & 7;
      const int gm = m0 + row, gk = k0 + c16 * 4;
      const bool ok = gm < a.M && gk < a.K;
      cp_async16(st + tc::swz(row, c16), ok ? (const void*)(a.A + (size_t)gm * a.ld_a + gk) : (const void*)a.A, ok);
    }
    for (int q = tid; q < BN * 8; q += 256) {
      const int row = q >> 3, c16 = q & 7;
      const size_t g = (size_t)(n0 + row) * a.kpad + k0 + c16 * 4;
      const uint32_t off = tc::swz(row, c16);
      cp_async16(st + 2 * A_BYTES + off, a.w_hi + g, true);
      cp_async16(st + 2 * A_BYTES + B_BYTES + off, a.w_lo + g, true);
    }
  };

  // prologue: two chunks in flight
  issue_loads(0); cp_async_commit();
  if (nchunks > 1) issue_loads(1);
  cp_async_commit();

  for (int c = 0; c < nchunks; ++c) {
    const int s = c % NS;
    // refill the stage that chunk c+2 will use (last read by the MMAs of chunk c-1)
    if (c + 2 < nchunks) {
      if (c >= 1) { tc::mbar_wait(&bar_free[(c + 2) % NS], (uint32_t)(((c - 1) / NS) & 1)); tc::tc_fence_after(); }
      issue_loads(c + 2);
    }
    cp_async_commit();
    cp_async_wait<2>();                      // this thread's copies of chunk c have landed (it splits exactly those)
    uint8_t* sA_hi = base + s * STAGE;
    uint8_t* sA_lo = sA_hi + A_BYTES;
    // split the A tile in place: hi = tf32-representable part, lo = a - hi (exact)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = tid + 256 * i;
      const uint32_t off = tc::swz(q >> 3, q & 7);
      const float4 v = *reinterpret_cast<const float4*>(sA_hi + off);
      float4 hi, lo;
      hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo.x = v.x - hi.x;
      hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo.y = v.y - hi.y;
      hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo.z = v.z - hi.z;
      hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo.w = v.w - hi.w;
      *reinterpret_cast<float4*>(sA_hi + off) = hi;
      *reinterpret_cast<float4*>(sA_lo + off) = lo;
    }
    tc::fence_proxy_async();                 // generic-proxy smem writes (cp.async + split) -> async proxy (tensor core)
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const uint32_t st = sbase + (uint32_t)(s * STAGE);
      const uint64_t dA_hi = tc::make_desc(st), dA_lo = tc::make_desc(st + A_BYTES);
      const uint64_t dB_hi = tc::make_desc(st + 2 * A_BYTES), dB_lo = tc::make_desc(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {                      // UMMA_K = 8 tf32 = 32 bytes -> +2 in the (>>4) address field
        const uint64_t adv = (uint64_t)(ks * 2);
        tc::mma_tf32(tmem_d, dA_lo + adv, dB_hi + adv, idesc, (c > 0 || ks > 0) ? 1u : 0u);   // small terms first
        tc::mma_tf32(tmem_d, dA_hi + adv, dB_lo + adv, idesc, 1u);
        tc::mma_tf32(tmem_d, dA_hi + adv, dB_hi + adv, idesc, 1u);
      }
      tc::mma_commit(&bar_free[s]);          // arrives when the MMAs above have finished reading this stage
      if (c == nchunks - 1) tc::mma_commit(&bar_done);
    }
  }
  // ---- epilogue: TMEM -> registers -> bias / activation / residual -> global (16-byte stores) ----
  tc::mbar_wait(&bar_done, 0);
  tc::tc_fence_after();
  {
    const int quad = warp & 3, half = warp >> 2;
    const int gm = m0 + quad * 32 + lane;
    constexpr int NCH = (BN + 31) / 32;                       // 32-column chunks in the tile
    const bool vec = (a.ld_out & 3) == 0;
    for (int ci = half; ci < NCH; ci += 2) {
      const int cc = ci * 32;
      float v[32];
      tc::tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)cc, v);
      if (gm < a.M) {
        float* op = a.out + (size_t)gm * a.ld_out + n0 + cc;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int ch = n0 + cc + j;
          if (cc + j + 3 < BN && ch + 3 < a.N && vec) {
            *reinterpret_cast<float4*>(op + j) = make_float4(tc_epilogue(v[j], ch, (size_t)gm, a.e), tc_epilogue(v[j + 1], ch + 1, (size_t)gm, a.e),
                                                             tc_epilogue(v[j + 2], ch + 2, (size_t)gm, a.e), tc_epilogue(v[j + 3], ch + 3, (size_t)gm, a.e));
          } else {
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (cc + j + t < BN && ch + t < a.N) op[j + t] = tc_epilogue(v[j + t], ch + t, (size_t)gm, a.e);
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_d, TMEM_COLS); }
}

template <int BN>
static bool launch_tc_bn(cudaStream_t s, const TcArgs& a, int npad) {
  const size_t smem = 3 * (2 * 128 * 128 + 2 * BN * 128);
  if (!ensure_dyn_smem(reinterpret_cast<const void*>(k_pointwise_tc<BN>), smem)) return false;
  dim3 grid((unsigned)ceil_div(a.M, 128), (unsigned)(npad / BN));
  k_pointwise_tc<BN><<<grid, 256, smem, s>>>(a);
  return true;
}

int pointwise_tc_tile_n(int N) {       // tile width the launcher will use for this N (weights are padded to a multiple of it)
  if (N <= 16) return 16;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 96) return 96;
  if (N == 160 || N == 240 || N == 480) return 80;
  return 128;
}

bool launch_pointwise_tc(cudaStream_t s, int M, int K, int N, const float* A, int ld_a, const float* w_hi, const float* w_lo,
                         int kpad, int npad, float* out, int ld_out, const Epilogue& e) {
  if (K % 4 != 0 || ld_a % 4 != 0 || kpad % 32 != 0) return false;
  TcArgs a{A, w_hi, w_lo, out, M, K, N, ld_a, ld_out, kpad, EpiDevTc{e.bias, e.residual, e.ld_res, e.act1, e.act2, e.act3}};
  const int bn = pointwise_tc_tile_n(N);
  if (npad % bn != 0) return false;
  bool ok = false;
  switch (bn) {
    case 16: ok = launch_tc_bn<16>(s, a, npad); break;
    case 32: ok = launch_tc_bn<32>(s, a, npad); break;
    case 64: ok = launch_tc_bn<64>(s, a, npad); break;
    case 80: ok = launch_tc_bn<80>(s, a, npad); break;
    case 96: ok = launch_tc_bn<96>(s, a, npad); break;
    case 128: ok = launch_tc_bn<128>(s, a, npad); break;
    default: return false;
  }
  if (!ok) return false;          // the caller falls back to the exact FFMA kernel
  count_launch();
  return true;
}

#else   // BSB_EMU: tcgen05 cannot be emulated; the emulator build never selects this path

int pointwise_tc_tile_n(int) { return 0; }
bool launch_pointwise_tc(cudaStream_t, int, int, int, const float*, int, const float*, const float*, int, int, float*, int, const Epilogue&) { return false; }

#endif

}  // namespace bsb

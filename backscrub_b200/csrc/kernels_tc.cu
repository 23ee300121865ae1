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
#include <algorithm>

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


// ---------------------------------------------------------------------------------------------------------
// Warp-specialised, persistent, TMA-fed version (the default tensor-core kernel).
//
//   warp 0      TMA producer: per K chunk of 32 floats one bulk tensor copy each for the raw A tile (128 rows, fp32 as
//               stored by the producing layer), W_hi and W_lo (BN rows) into 128-byte-swizzled stages (SASS: UTMALDG)
//   warp 1      MMA issuer: per chunk 12 tcgen05.mma kind::tf32 (a_lo*w_hi, a*w_lo, a*w_hi for 4 K steps; the tensor
//               core truncates the raw fp32 operand to tf32 itself, so the "hi" part of A needs no pass of its own),
//               tcgen05.commit -> the stage's `empty` barrier, and after a tile's last chunk -> `tmem_full`
//   warp 2      TMEM allocation (2 accumulator buffers of BN columns: the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 4-7   epilogue: tcgen05.ld (one 32-lane quadrant each) -> bias / activations / residual -> 16-byte stores
//   warps 8-11  splitter: a_lo = a - tf32_trunc(a) for their row of the raw tile (exact), written to the stage's A_lo tile
// CTAs are persistent (grid = #SMs) and walk the (M tile, N tile) list with a static stride.
// ---------------------------------------------------------------------------------------------------------
}  // namespace bsb
#include <cuda.h>
namespace bsb {

namespace tc {
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :: "r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
}  // namespace tc

struct TcMaps { CUtensorMap a, whi, wlo; };
struct Tc2Args { float* out; int M, K, N, ld_out, tiles_m, tiles_n, nchunks, mask_hi; EpiDevTc e; };

template <int BN>
struct Tc2Cfg {
  static constexpr int A_BYTES = 128 * 128;                       // 128 rows x 32 fp32
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;        // A raw | A lo | W hi | W lo
  static constexpr int NS = (STAGE * 4 <= 220 * 1024) ? 4 : 3;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM = NS * STAGE + BAR_BYTES + 1024 + 1024;   // + bias tile + slack to align the stages to 1024 bytes
  static constexpr uint32_t TMEM_COLS = BN <= 16 ? 32 : (BN <= 32 ? 64 : (BN <= 64 ? 128 : 256));   // two accumulators
};

template <int BN>
__global__ void __launch_bounds__(384, 1) k_pointwise_tc2(const __grid_constant__ TcMaps tm, const Tc2Args a) {
  using C = Tc2Cfg<BN>;
  extern __shared__ uint8_t smem_tc2[];
  const uint32_t raw_base = tc::smem_u32(smem_tc2);
  const uint32_t sbase = (raw_base + 1023u) & ~1023u;             // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* gen_base = smem_tc2 + (sbase - raw_base);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen_base + C::NS * C::STAGE);
  uint64_t* full_raw = bars;                    // [NS] TMA bytes landed
  uint64_t* full_lo = bars + C::NS;             // [NS] 128 splitter threads done
  uint64_t* empty = bars + 2 * C::NS;           // [NS] the MMAs that read the stage have completed
  uint64_t* tmem_full = bars + 3 * C::NS;       // [2]
  uint64_t* tmem_empty = bars + 3 * C::NS + 2;  // [2] 128 epilogue threads done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * C::NS + 4);
  float* sbias = reinterpret_cast<float*>(gen_base + C::NS * C::STAGE + C::BAR_BYTES);   // [BN] bias of the current tile (epilogue warps)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < C::NS; ++i) { tc::mbar_init(&full_raw[i], 1); tc::mbar_init(&full_lo[i], 128); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tmem_full[i], 1); tc::mbar_init(&tmem_empty[i], 128); }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int ntiles = a.tiles_m * a.tiles_n;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int m0 = (tile / a.tiles_n) * 128, n0 = (tile % a.tiles_n) * BN;
        for (int c = 0; c < a.nchunks; ++c, ++it) {
          const int s = it % C::NS;
          if (it >= C::NS) tc::mbar_wait(&empty[s], (uint32_t)((it / C::NS - 1) & 1));
          const uint32_t st = sbase + (uint32_t)(s * C::STAGE);
          tc::mbar_expect_tx(&full_raw[s], (uint32_t)(C::A_BYTES + 2 * C::B_BYTES));
          tc::tma_load_2d(st, &tm.a, c * 32, m0, &full_raw[s]);
          tc::tma_load_2d(st + 2 * C::A_BYTES, &tm.whi, c * 32, n0, &full_raw[s]);
          tc::tma_load_2d(st + 2 * C::A_BYTES + C::B_BYTES, &tm.wlo, c * 32, n0, &full_raw[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int acc = tcount & 1;
      if (tcount >= 2) tc::mbar_wait(&tmem_empty[acc], (uint32_t)((tcount / 2 - 1) & 1));      // the epilogue drained this accumulator
      tc::tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int c = 0; c < a.nchunks; ++c, ++it) {
        const int s = it % C::NS;
        const uint32_t ph = (uint32_t)((it / C::NS) & 1);
        tc::mbar_wait(&full_raw[s], ph);
        tc::mbar_wait(&full_lo[s], ph);
        tc::tc_fence_after();
        if (lane == 0) {
          const uint32_t st = sbase + (uint32_t)(s * C::STAGE);
          const uint64_t dA = tc::make_desc(st), dA_lo = tc::make_desc(st + C::A_BYTES);
          const uint64_t dB_hi = tc::make_desc(st + 2 * C::A_BYTES), dB_lo = tc::make_desc(st + 2 * C::A_BYTES + C::B_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {                  // UMMA_K = 8 tf32 = 32 bytes -> +2 in the (>>4) address field
            const uint64_t adv = (uint64_t)(ks * 2);
            tc::mma_tf32(tmem_d, dA_lo + adv, dB_hi + adv, idesc, (c > 0 || ks > 0) ? 1u : 0u);   // small terms first
            tc::mma_tf32(tmem_d, dA + adv, dB_lo + adv, idesc, 1u);
            tc::mma_tf32(tmem_d, dA + adv, dB_hi + adv, idesc, 1u);
          }
          tc::mma_commit(&empty[s]);                        // arrives when the MMAs above have finished reading the stage
          if (c == a.nchunks - 1) tc::mma_commit(&tmem_full[acc]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== epilogue =====
    const int quad = warp & 3;
    int tcount = 0;
    const bool vec = (a.ld_out & 3) == 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int acc = tcount & 1;
      const int m0 = (tile / a.tiles_n) * 128, n0 = (tile % a.tiles_n) * BN;
      tc::mbar_wait(&tmem_full[acc], (uint32_t)((tcount / 2) & 1));
      tc::tc_fence_after();
      const int gm = m0 + quad * 32 + lane;
      constexpr int NCH = (BN + 31) / 32;
      // bias of this tile's columns -> shared memory once (the 128 epilogue threads only: named barrier 1)
      for (int i = tid - 128; i < BN; i += 128) sbias[i] = (a.e.bias && n0 + i < a.N) ? __ldg(a.e.bias + n0 + i) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // the activations of the GEMM-heavy graphs (none / relu / relu6 after the bias, then an optional residual add) become
      // two clamps with +-inf bounds; anything else takes the generic per-element path
      const bool simple = a.e.act2 == ACT_NONE && a.e.act3 == ACT_NONE && (a.e.act1 == ACT_NONE || a.e.act1 == ACT_RELU || a.e.act1 == ACT_RELU6);
      const float lo_b = a.e.act1 == ACT_NONE ? -INFINITY : 0.f, hi_b = a.e.act1 == ACT_RELU6 ? 6.f : INFINITY;
#pragma unroll 1
      for (int ci = 0; ci < NCH; ++ci) {
        const int cc = ci * 32;
        float v[32];
        tc::tmem_ld32(tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(quad * 32) << 16) + (uint32_t)cc, v);
        if (gm < a.M) {
          float* op = a.out + (size_t)gm * a.ld_out + n0 + cc;
          const float* rp = a.e.residual ? a.e.residual + (size_t)gm * a.e.ld_res + n0 + cc : nullptr;
          const bool rvec = rp && (a.e.ld_res & 3) == 0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int ch = n0 + cc + j;
            if (cc + j >= BN || ch >= a.N) break;
            if (simple && cc + j + 3 < BN && ch + 3 < a.N && vec && (!rp || rvec)) {
              const float4 b4 = *reinterpret_cast<const float4*>(sbias + cc + j);
              float4 r = make_float4(fminf(fmaxf(v[j] + b4.x, lo_b), hi_b), fminf(fmaxf(v[j + 1] + b4.y, lo_b), hi_b),
                                     fminf(fmaxf(v[j + 2] + b4.z, lo_b), hi_b), fminf(fmaxf(v[j + 3] + b4.w, lo_b), hi_b));
              if (rp) { const float4 q = __ldg(reinterpret_cast<const float4*>(rp + j)); r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w; }
              *reinterpret_cast<float4*>(op + j) = r;
            } else {
#pragma unroll
              for (int t = 0; t < 4; ++t)
                if (cc + j + t < BN && ch + t < a.N) op[j + t] = tc_epilogue(v[j + t], ch + t, (size_t)gm, a.e);
            }
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");            // everyone is done with sbias before the next tile refills it
      tc::tc_fence_before();
      tc::mbar_arrive(&tmem_empty[acc]);
    }
  } else if (warp >= 8) {
    // ===== splitter: A_lo = a - tf32_trunc(a), one row (128 bytes) per thread =====
    const int row = tid - 256;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int c = 0; c < a.nchunks; ++c, ++it) {
        const int s = it % C::NS;
        tc::mbar_wait(&full_raw[s], (uint32_t)((it / C::NS) & 1));
        uint8_t* sA = gen_base + s * C::STAGE;
        uint8_t* sLo = sA + C::A_BYTES;
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16) {
          const uint32_t off = tc::swz(row, c16);
          const float4 v = *reinterpret_cast<const float4*>(sA + off);
          float4 lo;
          lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          *reinterpret_cast<float4*>(sLo + off) = lo;
          if (a.mask_hi) {        // measurement switch: do not rely on the tensor core truncating the raw operand
            float4 hi;
            hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
            hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
            *reinterpret_cast<float4*>(sA + off) = hi;
          }
        }
        tc::fence_proxy_async();                 // generic-proxy writes of A_lo -> visible to the tensor core (async proxy)
        tc::mbar_arrive(&full_lo[s]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, C::TMEM_COLS); }
}

typedef CUresult (*TcEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TcEncodeFn tc_encode_fn() {
  static TcEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); p = nullptr; }
    return reinterpret_cast<TcEncodeFn>(p);
  }();
  return fn;
}
// fp32 matrix [rows][cols], row pitch in floats; box = 32 floats x box_rows, 128-byte swizzle
static bool tc_make_map(CUtensorMap* m, const float* base, size_t rows, size_t cols, size_t pitch_floats, unsigned box_rows) {
  TcEncodeFn enc = tc_encode_fn();
  if (!enc || (reinterpret_cast<uintptr_t>(base) & 15) || (pitch_floats & 3)) return false;
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstr[1] = {pitch_floats * 4};
  const cuuint32_t box[2] = {32, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN>
static bool launch_tc2_bn(cudaStream_t s, int M, int K, int N, const float* A, int ld_a, const float* w_hi, const float* w_lo, int kpad, int npad,
                          float* out, int ld_out, const EpiDevTc& e) {
  TcMaps tm;
  if (!tc_make_map(&tm.a, A, (size_t)M, (size_t)K, (size_t)ld_a, 128)) return false;
  if (!tc_make_map(&tm.whi, w_hi, (size_t)npad, (size_t)kpad, (size_t)kpad, BN)) return false;
  if (!tc_make_map(&tm.wlo, w_lo, (size_t)npad, (size_t)kpad, (size_t)kpad, BN)) return false;
  Tc2Args a{out, M, K, N, ld_out, ceil_div(M, 128), npad / BN, kpad / 32, tuning().tc_mask_hi, e};
  if (!ensure_dyn_smem(reinterpret_cast<const void*>(k_pointwise_tc2<BN>), Tc2Cfg<BN>::SMEM)) return false;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(sms, a.tiles_m * a.tiles_n);
  k_pointwise_tc2<BN><<<grid, 384, Tc2Cfg<BN>::SMEM, s>>>(tm, a);
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
  if (tuning().tc_variant == 2) {
    const EpiDevTc ed{e.bias, e.residual, e.ld_res, e.act1, e.act2, e.act3};
    switch (bn) {
      case 16: ok = launch_tc2_bn<16>(s, M, K, N, A, ld_a, w_hi, w_lo, kpad, npad, out, ld_out, ed); break;
      case 32: ok = launch_tc2_bn<32>(s, M, K, N, A, ld_a, w_hi, w_lo, kpad, npad, out, ld_out, ed); break;
      case 64: ok = launch_tc2_bn<64>(s, M, K, N, A, ld_a, w_hi, w_lo, kpad, npad, out, ld_out, ed); break;
      case 80: ok = launch_tc2_bn<80>(s, M, K, N, A, ld_a, w_hi, w_lo, kpad, npad, out, ld_out, ed); break;
      case 96: ok = launch_tc2_bn<96>(s, M, K, N, A, ld_a, w_hi, w_lo, kpad, npad, out, ld_out, ed); break;
      case 128: ok = launch_tc2_bn<128>(s, M, K, N, A, ld_a, w_hi, w_lo, kpad, npad, out, ld_out, ed); break;
      default: return false;
    }
    if (!ok) return false;
    count_launch();
    return true;
  }
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

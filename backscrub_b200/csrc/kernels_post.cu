// backscrub_b200/csrc/kernels_post.cu — TMA-staged variant of the fused post stage (sm_100a).
//
// Same stage as k_post_fast (kernels_img.cu): mask upsample (cv::resize 8UC1) + cv::blur 5x5 inside the ROI
// (lib/libbackscrub.cc:366-371), alpha_blend (app/deepseg.cc:108-134), convert_rgb_to_yuyv (:87-106), mask store —
// reorganised around the copy engine:
//   * every global access of the tile is a TMA bulk tensor copy (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG)
//     between global memory and shared memory, issued by one thread: no per-thread address arithmetic, no
//     sector bookkeeping, image edges clipped by the tensor map;
//   * the source patch of the small mask arrives first; if it is uniformly 255 (background) or 0 (person) — exact
//     for any interpolation weights, see k_post_fast — the tile needs NO per-pixel arithmetic on the copy path:
//       background tile:  out <- background tile, YUYV <- cached YUYV of the background, mask <- 255
//                         (the camera frame is not even read),
//       person tile:      out <- frame tile (camera YUYV converted per thread when the frames are in wire format),
//       mixed tile:       the k_post_fast arithmetic on shared-memory operands;
//   * results are staged in shared memory and leave with TMA stores.
// Arithmetic is shared with k_post_fast through post_math.h, so both kernels produce the oracle's bits.
#include "kernels.h"
#include "post_math.h"

#ifndef BSB_EMU
#include <cuda.h>   // CUtensorMap types only; cuTensorMapEncodeTiled is resolved at run time (no libcuda link)
#endif

namespace bsb {

void count_launch();

#ifndef BSB_EMU

namespace tma {
BSB_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
BSB_D void mbar_init(uint64_t* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
BSB_D void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
BSB_D void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
BSB_D bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
BSB_D void mbar_wait(uint64_t* bar, unsigned parity) { while (!mbar_try_wait(bar, parity)) {} }
BSB_D void load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
BSB_D void store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               :: "l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
BSB_D void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
BSB_D void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
BSB_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
BSB_D void prefetch_map(const CUtensorMap* map) { asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory"); }
}  // namespace tma

struct PostMaps { CUtensorMap frame, bg, bgy, out, yuyv, mask, ofinal; };
struct PostTmaCfg { int has_out, has_yuyv, has_mask, has_bgy; };

// Tile geometry.  A CTA owns one TW x 32 tile (TW = 64 or 128), one thread per 16 pixels of a row.  Measured on a B200:
//   * two 128-wide tiles per CTA (95 KB, two CTAs per SM) is SLOWER than one (88.6 vs 65.6 us per 32-frame 720p launch,
//     run r2i): the halved warp count costs more than the amortised load latency saves;
//   * the stage is a dependent chain per CTA (patch -> classify -> tiles -> arithmetic -> stores), so what hides latency
//     is the number of independent chains per SM: a 64-wide tile needs ~27 KB and 128 threads, eight CTAs per SM instead
//     of four with the same number of resident warps (tuning switch post_tile, see DESIGN.md for the measured numbers).
// The patch of the small mask a (TW + 4) x 36 halo tile can touch has <= PT_RMAX rows and <= PCOLS columns (up-scales
// >= ~1.75x).  A TMA box must START on a 16-byte boundary of global memory (an unaligned innermost coordinate raises
// "illegal instruction"), so the box begins at the patch's first column rounded down to 16 and is 16 columns wider.
constexpr int PT_RMAX = 24;
template <int TW> struct PtL {
  static constexpr int NT = TW * PF_H / PF_PX;                   // threads
  static constexpr int UW = TW + 4, US = TW + 8;                 // halo tile width, its row stride (u16 elements)
  static constexpr int PCOLS = TW == 128 ? 80 : 48, PW = PCOLS + 16;
  static constexpr int F_BYTES = TW * PF_H * 3, Y_BYTES = TW * PF_H * 2, M_BYTES = TW * PF_H;
  static constexpr int OFF_F = 0, OFF_B = F_BYTES, OFF_Y = 2 * F_BYTES, OFF_P = 2 * F_BYTES + Y_BYTES;
  static constexpr int OFF_HS = OFF_P + PT_RMAX * PW;            // Hs [PT_RMAX][US] u16, later Vs [PF_H][US] u16
  static constexpr int HV_BYTES = (PT_RMAX > PF_H ? PT_RMAX : PF_H) * US * 2;
  static constexpr int OFF_ROWS = OFF_HS + PT_RMAX * US * 2;     // [PF_UH] uint4 in the part of Vs that Hs does not use
  static constexpr int OFF_US = OFF_HS + HV_BYTES;               // Us [PF_UH][US] u16, later the mask staging tile
  static constexpr int OFF_BAR = OFF_US + PF_UH * US * 2;
  static constexpr int OFF_GEO = OFF_BAR + 32;                   // 12 ints of tile geometry / classification
  static constexpr int SMEM = OFF_GEO + 48;
  static constexpr int CTAS = TW == 128 ? 4 : 8;
  static_assert(PW % 16 == 0 && OFF_P % 128 == 0 && OFF_HS % 16 == 0 && OFF_US % 128 == 0 && OFF_ROWS % 16 == 0 && OFF_BAR % 8 == 0 &&
                (PF_H - PT_RMAX) * US * 2 >= PF_UH * 16 && PF_UH * US * 2 >= M_BYTES && US % 8 == 0 && NT % 64 == 0 && NT >= PF_UH &&
                (SMEM + 1024) * CTAS <= 228 * 1024, "smem layout");
};

template <bool IN_YUYV, int TW>
__global__ void __launch_bounds__(PtL<TW>::NT, PtL<TW>::CTAS) k_post_tma(const __grid_constant__ PostMaps tm, const PostArgs a, const PostTmaCfg cfg) {
  using L = PtL<TW>;
  constexpr int NT = L::NT, UW = L::UW, US = L::US, PW = L::PW;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned short* Hs = reinterpret_cast<unsigned short*>(smem + L::OFF_HS);
  unsigned short* Us = reinterpret_cast<unsigned short*>(smem + L::OFF_US);
  unsigned short* Vs = Hs;
  uint4* rows = reinterpret_cast<uint4*>(smem + L::OFF_ROWS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);      // patch | background (+ its YUYV) | camera tile
  uint8_t* sF = smem + L::OFF_F;          // frame tile [32][3 TW] (BGR) or [32][2 TW] (YUYV); later the blended tile
  uint8_t* sB = smem + L::OFF_B;          // background tile [32][3 TW]
  uint8_t* sY = smem + L::OFF_Y;          // YUYV tile [32][2 TW]: cached background YUYV in, result out
  uint8_t* sM = smem + L::OFF_US;         // mask tile [32][TW] (reuses Us, dead by then)
  uint8_t* sP = smem + L::OFF_P;          // source patch of the small mask [PT_RMAX][PW]
  uint64_t* barP = bars; uint64_t* barB = bars + 1; uint64_t* barF = bars + 2;

  const int b = blockIdx.z;
  const int ty0 = blockIdx.y * PF_H, tx0 = blockIdx.x * TW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* geo = reinterpret_cast<int*>(smem + L::OFF_GEO);                 // tile geometry, published by thread 0

  // ---- thread 0: geometry of the tile and ALL the loads, before anybody else does anything.  The geometry is uniform
  //      over the CTA: computing it once and publishing it through shared memory takes ~25 instructions off every other
  //      warp (the stage is half issue-bound, profiles/r2_ncu_k_post_tma.txt) ----
  if (tid == 0) {
    tma::mbar_init(barP, 1); tma::mbar_init(barB, 1); tma::mbar_init(barF, 1);
    tma::fence_barrier_init();
    const bool hit = tx0 < a.roi_x + a.roi_w && tx0 + TW > a.roi_x && ty0 < a.roi_y + a.roi_h && ty0 + PF_H > a.roi_y;
    // the patch of the small mask this tile touches comes from two per-context tables (engine.cu): two independent
    // loads whose addresses need nothing but the block index, then the patch request — the head of the CTA's critical
    // path — goes out first ...
    int gy_lo = 0, gx_lo = 0, rmin = 0, nrows = 0, cmin = 0, ncols = 0, coff = 0;
    if (hit) {
      const int2 gr = __ldg(a.geo_rows + blockIdx.y), gc = __ldg((TW == 64 ? a.geo_cols64 : a.geo_cols128) + blockIdx.x);
      gy_lo = ty0 - a.roi_y - 2; gx_lo = tx0 - a.roi_x - 2;
      rmin = gr.x; nrows = gr.y; cmin = gc.x; ncols = gc.y;
      coff = (a.out_x + cmin) & 15;                     // the patch's first column inside the 16-byte aligned box
      tma::mbar_expect_tx(barP, PT_RMAX * PW);
      tma::load_3d(sP, &tm.ofinal, a.out_x + cmin - coff, a.out_y + rmin, b, barP);
    }
    // ... then the big tiles: the background tile (L2-resident for a still image) and its cached YUYV (a background
    // tile needs nothing else, a mixed tile has its second operand early) ...
    int bgi = 0;
    if (a.bg_cursor) bgi = (int)(((unsigned)__ldg(a.bg_cursor) + (unsigned)b * (unsigned)a.bg_advance) % (unsigned)a.bg_count);
    else if (a.bg_stride) bgi = b;
    tma::mbar_expect_tx(barB, (unsigned)L::F_BYTES + (cfg.has_bgy ? (unsigned)L::Y_BYTES : 0u));
    tma::load_3d(sB, &tm.bg, blockIdx.x * (TW * 3 / 4), ty0, bgi, barB);
    if (cfg.has_bgy) tma::load_3d(sY, &tm.bgy, blockIdx.x * (TW / 2), ty0, bgi, barB);
    // ... and the camera tile when the tile can contain a person at all: waiting for the classification first would put a
    // second memory round trip on the critical path of every person / mixed tile
    if (hit) {
      tma::mbar_expect_tx(barF, IN_YUYV ? (unsigned)L::Y_BYTES : (unsigned)L::F_BYTES);
      tma::load_3d(sF, &tm.frame, blockIdx.x * (IN_YUYV ? TW / 2 : TW * 3 / 4), ty0, b, barF);
    }
    const bool inside = tx0 >= a.roi_x && min(tx0 + TW, a.W) <= a.roi_x + a.roi_w && ty0 >= a.roi_y && min(ty0 + PF_H, a.H) <= a.roi_y + a.roi_h;
    *reinterpret_cast<int4*>(geo) = make_int4(hit ? 1 : 0, gy_lo, gx_lo, rmin);
    *reinterpret_cast<int4*>(geo + 4) = make_int4(nrows, cmin, ncols, coff);
    geo[8] = inside ? 1 : 0;
  }
  __syncthreads();
  const int4 g0 = *reinterpret_cast<const int4*>(geo), g1 = *reinterpret_cast<const int4*>(geo + 4);
  const bool hits_roi = g0.x != 0;
  const int gy_lo = g0.y, gx_lo = g0.z, rmin = g0.w, nrows = g1.x, cmin = g1.y, ncols = g1.z, coff = g1.w;

  const int lx = (tid % (TW / PF_PX)) * PF_PX, ly = tid / (TW / PF_PX);
  const int y = ty0 + ly;

  int tile_const = -1, kind = 0;                         // background / person / mixed
  if (hits_roi) {
    if (warp == 0) {
      // warp 0 alone waits for the patch and classifies it: all-255 / all-0 test, one 32-bit word per lane, over the box
      // columns [0, coff + ncols) rounded up to words — a few columns more than the patch.  The classification only
      // selects the code path (the mixed path is always correct), so a conservative test costs nothing in exactness and
      // needs no per-byte masks.
      tma::mbar_wait(barP, 0);
      constexpr int WP2 = PW / 4 <= 16 ? 16 : 32;                      // words per patch row, rounded up to a power of two
      const int wpr = (coff + ncols + 3) >> 2;                         // words per row to look at (<= PW / 4)
      const int wc = lane % WP2;
      bool hi = true, lo = true;
      if (wc < wpr) {
        for (int r = lane / WP2; r < nrows; r += 32 / WP2) {
          const unsigned v = *reinterpret_cast<const unsigned*>(sP + r * PW + wc * 4);
          hi = hi && (v == 0xffffffffu);
          lo = lo && (v == 0u);
        }
      }
      const bool all_hi = __all_sync(0xffffffffu, hi), all_lo = __all_sync(0xffffffffu, lo);
      if (lane == 0) {
        const int tc = all_hi ? 255 : (all_lo ? 0 : -1);
        *reinterpret_cast<int2*>(geo + 10) = make_int2(tc, tc == 255 ? 0 : ((tc == 0 && geo[8]) ? 1 : 2));
      }
    } else if (tid - 32 < PF_UH) {                    // row parameters of the vertical resize pass (used by mixed tiles)
      const int uy = tid - 32;
      int gy = gy_lo + uy;
      gy = gy < 0 ? -gy : gy; gy = gy >= a.roi_h ? 2 * a.roi_h - 2 - gy : gy;     // reflect-101 (single fold)
      gy = min(max(gy, 0), a.roi_h - 1);
      rows[uy] = make_uint4((unsigned)__ldg(a.tab.yofs0 + gy), (unsigned)__ldg(a.tab.yofs1 + gy),
                            (unsigned)(int)__ldg(a.tab.yw + 2 * gy) << 16, (unsigned)(int)__ldg(a.tab.yw + 2 * gy + 1) << 16);
    }
    __syncthreads();
    const int2 tk = *reinterpret_cast<const int2*>(geo + 10);
    tile_const = tk.x; kind = tk.y;
  }

  if (kind == 2 && tile_const < 0) {
    // ---- A1: horizontal pass of cv::resize on the patch rows (see k_post_fast) ----
    {
      constexpr int HALF = NT / 2;
      const int phase = tid / HALF;
      for (int ux = tid % HALF; ux < UW; ux += HALF) {
        int gx = gx_lo + ux;
        gx = gx < 0 ? -gx : gx; gx = gx >= a.roi_w ? 2 * a.roi_w - 2 - gx : gx;
        gx = min(max(gx, 0), a.roi_w - 1);
        const uint2 xc = __ldg(a.tab.xcol + gx);
        const int sx = (int)(xc.x & 0xffffu) - cmin + coff, sx1 = (int)(xc.x >> 16) - cmin + coff;
        const int a0 = (int)(short)(xc.y & 0xffffu), a1 = (int)(short)(xc.y >> 16);
        const uint8_t* pr = sP + phase * PW;
        unsigned short* hp = Hs + phase * US + ux;
        for (int r = phase; r < nrows; r += 2) {
          *hp = (unsigned short)(((int)pr[sx] * a0 + (int)pr[sx1] * a1) >> 4);
          pr += 2 * PW; hp += 2 * US;
        }
      }
    }
    __syncthreads();
    // ---- A2: vertical pass -> upsampled tile ----
    for (int uy = warp; uy < PF_UH; uy += NT / 32) {
      const uint4 rp = rows[uy];
      const unsigned short* h0 = Hs + ((int)rp.x - rmin) * US + lane;
      const unsigned short* h1 = Hs + ((int)rp.y - rmin) * US + lane;
      unsigned short* up = Us + uy * US + lane;
#pragma unroll
      for (int k = 0; k < TW / 32; ++k)
        up[32 * k] = (unsigned short)((__umulhi(rp.z, (unsigned)h0[32 * k]) + __umulhi(rp.w, (unsigned)h1[32 * k]) + 2u) >> 2);
      if (lane < UW - TW)
        up[TW] = (unsigned short)((__umulhi(rp.z, (unsigned)h0[TW]) + __umulhi(rp.w, (unsigned)h1[TW]) + 2u) >> 2);
    }
    __syncthreads();
    // ---- B: vertical 5-sums, two columns per word, sliding window over 8 rows (Vs overwrites Hs and the row table) ----
    for (int it = tid; it < (UW / 2) * 4; it += NT) {
      const int pair = it % (UW / 2), seg = it / (UW / 2);
      const unsigned* up = reinterpret_cast<const unsigned*>(Us + (seg * 8) * US) + pair;
      unsigned* vp = reinterpret_cast<unsigned*>(Vs + (seg * 8) * US) + pair;
      unsigned u[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) u[k] = up[k * (US / 2)];
      unsigned v = u[0] + u[1] + u[2] + u[3] + u[4];
      vp[0] = v;
#pragma unroll
      for (int k = 1; k < 8; ++k) { v = v + u[k + 4] - u[k - 1]; vp[k * (US / 2)] = v; }
    }
    __syncthreads();
  }

  const int x0 = tx0 + lx;
  uint4* mdst = reinterpret_cast<uint4*>(sM + ly * TW + lx);
  uint4* ydst = reinterpret_cast<uint4*>(sY + ly * (TW * 2) + lx * 2);
  const uint8_t* out_src = sB;

  // a background tile with a cached YUYV copy is moved by the copy engine alone: only thread 0 (which issues the stores)
  // waits for the background tiles, and for the (speculatively loaded, unused) camera tile after the stores are on their
  // way — shared memory must be quiet when the CTA exits
  const bool engine_only = kind == 0 && !(cfg.has_yuyv && !cfg.has_bgy);
  if (!engine_only) tma::mbar_wait(barB, 0);
  if (hits_roi && kind != 0) tma::mbar_wait(barF, 0);
  if (kind == 0) {
    // ---- background tile: out = background tile, YUYV = cached YUYV tile, mask = 255.  No per-pixel arithmetic ----
    if (cfg.has_mask) *mdst = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (cfg.has_yuyv && !cfg.has_bgy) {
      const uint4* gq = reinterpret_cast<const uint4*>(sB + ly * (TW * 3) + lx * 3);
      const uint4 g0 = gq[0], g1 = gq[1], g2 = gq[2];
      const unsigned gg[12] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x, g2.y, g2.z, g2.w};
      unsigned m[PF_PX], o[12], yy[8];
#pragma unroll
      for (int i = 0; i < PF_PX; ++i) m[i] = 255u;
      post_blend16<false, true>(gg, gg, m, o, yy);
      ydst[0] = make_uint4(yy[0], yy[1], yy[2], yy[3]); ydst[1] = make_uint4(yy[4], yy[5], yy[6], yy[7]);
    }
  } else {
    unsigned m[PF_PX];
    if (kind == 1) {
#pragma unroll
      for (int i = 0; i < PF_PX; ++i) m[i] = 0u;
    } else {
      const bool row_in = y >= a.roi_y && y < a.roi_y + a.roi_h;
      post_mask16(a, row_in, tile_const, x0, Vs + ly * US + lx, m);
    }
    unsigned f[12], gg[12];
    if (IN_YUYV) {
      const uint4* fq = reinterpret_cast<const uint4*>(sF + ly * (TW * 2) + lx * 2);
      const uint4 w0 = fq[0], w1 = fq[1];
      const unsigned wa[4] = {w0.x, w0.y, w0.z, w0.w}, wb[4] = {w1.x, w1.y, w1.z, w1.w};
      yuyv8_to_bgr24(wa, f); yuyv8_to_bgr24(wb, f + 6);
    } else {
      const uint4* fq = reinterpret_cast<const uint4*>(sF + ly * (TW * 3) + lx * 3);
      const uint4 f0 = fq[0], f1 = fq[1], f2 = fq[2];
      f[0] = f0.x; f[1] = f0.y; f[2] = f0.z; f[3] = f0.w; f[4] = f1.x; f[5] = f1.y; f[6] = f1.z; f[7] = f1.w;
      f[8] = f2.x; f[9] = f2.y; f[10] = f2.z; f[11] = f2.w;
    }
    if (kind == 2) {
      const uint4* gq = reinterpret_cast<const uint4*>(sB + ly * (TW * 3) + lx * 3);
      const uint4 g0 = gq[0], g1 = gq[1], g2 = gq[2];
      gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w; gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
      gg[8] = g2.x; gg[9] = g2.y; gg[10] = g2.z; gg[11] = g2.w;
    } else {
#pragma unroll
      for (int i = 0; i < 12; ++i) gg[i] = f[i];
    }
    unsigned o[12], yy[8];
    post_blend16<true, true>(f, gg, m, o, yy);
    // the YUYV camera tile is fully consumed before the (wider) BGR tile overwrites it; the mask staging tile reuses Us,
    // whose last readers (phase B) are behind a barrier already
    if (IN_YUYV) __syncthreads();
    uint4* odst = reinterpret_cast<uint4*>(sF + ly * (TW * 3) + lx * 3);
    odst[0] = make_uint4(o[0], o[1], o[2], o[3]); odst[1] = make_uint4(o[4], o[5], o[6], o[7]); odst[2] = make_uint4(o[8], o[9], o[10], o[11]);
    ydst[0] = make_uint4(yy[0], yy[1], yy[2], yy[3]); ydst[1] = make_uint4(yy[4], yy[5], yy[6], yy[7]);
    *mdst = make_uint4(m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24), m[4] | (m[5] << 8) | (m[6] << 16) | (m[7] << 24),
                       m[8] | (m[9] << 8) | (m[10] << 16) | (m[11] << 24), m[12] | (m[13] << 8) | (m[14] << 16) | (m[15] << 24));
    out_src = sF;
  }
  tma::fence_proxy_async();            // generic-proxy writes of the staging tiles -> visible to the TMA engine
  __syncthreads();
  if (tid == 0) {
    if (engine_only) tma::mbar_wait(barB, 0);
    if (cfg.has_out) tma::store_3d(&tm.out, out_src, blockIdx.x * (TW * 3 / 4), ty0, b);
    if (cfg.has_yuyv) tma::store_3d(&tm.yuyv, sY, blockIdx.x * (TW / 2), ty0, b);
    if (cfg.has_mask) tma::store_3d(&tm.mask, sM, tx0, ty0, b);
    tma::store_commit();
    if (hits_roi && kind == 0) tma::mbar_wait(barF, 0);
    tma::store_wait_read();            // shared memory may be handed to the next CTA only after the engine has read it
  }
}

// ---- host side: tensor maps --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); p = nullptr; }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// [n2][n1][row_bytes] bytes, row pitch / plane stride in bytes; box = box_bytes x box_rows x 1.  Elements are 32-bit
// words when `words` (row_bytes % 4 == 0), else bytes.
static bool make_map(CUtensorMap* m, const void* base, bool words, size_t row_bytes, size_t n1, size_t n2, size_t pitch, size_t stride,
                     unsigned box_bytes, unsigned box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  const unsigned es = words ? 4u : 1u;
  const cuuint64_t gdim[3] = {row_bytes / es, n1, n2 ? n2 : 1};
  const cuuint64_t gstr[2] = {pitch, stride ? stride : pitch * n1};
  const cuuint32_t box[3] = {box_bytes / es, box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, words ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool build_post_maps(const PostArgs& a, PostMaps* tm, int TW) {
  const unsigned bw3 = (unsigned)TW * 3, bw2 = (unsigned)TW * 2, pw = TW == 128 ? (unsigned)PtL<128>::PW : (unsigned)PtL<64>::PW;
  const size_t W = (size_t)a.W, H = (size_t)a.H, B = (size_t)a.B;
  const size_t nbg = a.bg_cursor ? (size_t)a.bg_count : (a.bg_stride ? B : 1);
  bool ok = true;
  if (a.yuyv_in) ok = ok && make_map(&tm->frame, a.yuyv_in, true, W * 2, H, B, W * 2, a.yuyv_in_stride, bw2, PF_H);
  else ok = ok && make_map(&tm->frame, a.frames, true, W * 3, H, B, a.frame_pitch, a.frame_stride, bw3, PF_H);
  ok = ok && make_map(&tm->bg, a.bg, true, W * 3, H, nbg, a.bg_pitch, a.bg_stride ? a.bg_stride : a.bg_pitch * H, bw3, PF_H);
  // unused maps still have to be valid descriptors: alias them to a live one
  if (a.bg_yuyv) ok = ok && make_map(&tm->bgy, a.bg_yuyv, true, W * 2, H, nbg, W * 2, W * 2 * H, bw2, PF_H); else tm->bgy = tm->bg;
  if (a.out) ok = ok && make_map(&tm->out, a.out, true, W * 3, H, B, a.out_pitch, a.out_stride, bw3, PF_H); else tm->out = tm->bg;
  if (a.yuyv) ok = ok && make_map(&tm->yuyv, a.yuyv, true, W * 2, H, B, W * 2, a.yuyv_stride, bw2, PF_H); else tm->yuyv = tm->bg;
  if (a.mask) ok = ok && make_map(&tm->mask, a.mask, false, W, H, B, W, a.mask_stride, (unsigned)TW, PF_H); else tm->mask = tm->bg;
  ok = ok && make_map(&tm->ofinal, a.ofinal, false, (size_t)a.ow, (size_t)a.oh, B, (size_t)a.opitch, (size_t)a.opitch * a.oh, pw, PT_RMAX);
  return ok;
}

// measured (profiles/r2_post_tile_ab.txt): 64-wide tiles (eight CTAs per SM) win at 720p, 60.9 vs 65.6 us per 32-frame launch;
// at 3840x2160, where > 95 % of the tiles are pure copies, 128-wide tiles halve the per-CTA fixed cost: 136.7 vs 140.2 us
// per 8-frame launch (0.81 vs 0.79 of the measured HBM peak)
static int post_tile_width(const PostArgs& a) {
  const int t = tuning().post_tile;
  if (t == 64 || t == 128) return t;
  return a.W >= 2560 ? 128 : 64;
}

static bool post_tma_shape_ok(const PostArgs& a) {
  if (!tuning().post_tma || !encode_fn()) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (a.B < 1 || a.W % 16 != 0 || a.area2x2 || a.tab.xcol == nullptr || a.geo_rows == nullptr) return false;
  if (a.yuyv_in) { if (!al16(a.yuyv_in) || a.yuyv_in_stride % 16) return false; }
  else if (!al16(a.frames) || a.frame_pitch % 16 || a.frame_stride % 16) return false;
  if (!al16(a.bg) || a.bg_pitch % 16 || a.bg_stride % 16 || (a.bg_yuyv && !al16(a.bg_yuyv))) return false;
  if (a.out && (!al16(a.out) || a.out_pitch % 16 || a.out_stride % 16)) return false;
  if (a.yuyv && (!al16(a.yuyv) || a.yuyv_stride % 16)) return false;
  if (a.mask && (!al16(a.mask) || a.mask_stride % 16)) return false;
  if (!(a.out || a.yuyv || a.mask) || !al16(a.ofinal) || a.opitch % 16) return false;
  if (a.ow > 32000 || a.oh > 32000 || a.roi_w < 8 || a.roi_h < 8) return false;
  const double scale_y = (double)a.out_h / (double)a.roi_h, scale_x = (double)a.out_w / (double)a.roi_w;
  const int tw = post_tile_width(a);
  if ((int)(PF_UH * scale_y) + 3 > PT_RMAX || (int)((tw + 4) * scale_x) + 4 > (tw == 128 ? PtL<128>::PCOLS : PtL<64>::PCOLS)) return false;
  return true;
}

bool post_tma_eligible(const PostArgs& a) {
  PostMaps tm;
  return post_tma_shape_ok(a) && build_post_maps(a, &tm, post_tile_width(a));
}

template <bool IN_YUYV, int TW>
static bool launch_post_tma_t(cudaStream_t s, const PostMaps& tm, const PostArgs& a, const PostTmaCfg& cfg) {
  using L = PtL<TW>;
  if (!ensure_dyn_smem(reinterpret_cast<const void*>(k_post_tma<IN_YUYV, TW>), L::SMEM)) return false;
  const dim3 grid((unsigned)ceil_div(a.W, TW), (unsigned)ceil_div(a.H, PF_H), (unsigned)a.B);
  k_post_tma<IN_YUYV, TW><<<grid, L::NT, L::SMEM, s>>>(tm, a, cfg);
  return true;
}

bool launch_post_tma(cudaStream_t s, const PostArgs& a) {
  PostMaps tm;
  const int tw = post_tile_width(a);
  if (!post_tma_shape_ok(a) || !build_post_maps(a, &tm, tw)) return false;
  const PostTmaCfg cfg{a.out ? 1 : 0, a.yuyv ? 1 : 0, a.mask ? 1 : 0, a.bg_yuyv ? 1 : 0};
  bool ok;
  if (a.yuyv_in) ok = tw == 128 ? launch_post_tma_t<true, 128>(s, tm, a, cfg) : launch_post_tma_t<true, 64>(s, tm, a, cfg);
  else ok = tw == 128 ? launch_post_tma_t<false, 128>(s, tm, a, cfg) : launch_post_tma_t<false, 64>(s, tm, a, cfg);
  if (ok) count_launch();
  return ok;
}

#else   // BSB_EMU: TMA cannot be emulated; the emulator build always takes k_post_fast / k_post

bool post_tma_eligible(const PostArgs&) { return false; }
bool launch_post_tma(cudaStream_t, const PostArgs&) { return false; }

#endif

}  // namespace bsb

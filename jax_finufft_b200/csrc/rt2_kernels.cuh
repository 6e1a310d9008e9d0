// rt2_kernels.cuh -- register-TILE spreading / interpolation, 2-D, float, ns <= 8.
// Replaces spread_2d_subprob and interp_2d_nupts_driven / interp_2d_subprob
// (V/src/cuda/2d/spreadinterp2d.cuh:116-200, 203-330) on the 2-D c64 paths (BASELINE configs C2
// and C4) when the point set is dense enough to fill the bins.
//
// Same idea as the 3-D sliding-window kernels (swr_kernels.cuh) with the ring dropped, because a
// 2-D window is small enough to live in registers whole:
//   * points are binned by the ANCHOR cell of their ns-wide stencil (sort.cu, swr geometry), bins
//     of BX x BY = (17 - ns)^2 anchor cells (10 x 10 at ns = 7); no sub-key, any order in a bin;
//   * ONE WARP owns a subproblem (<= maxsub points of one bin) and the 16 x 16 fine-grid cells it
//     can touch: lane (r, q) holds x cells {2q, 2q+1} of rows {r, r+4, r+8, r+12} -- 8 complex
//     accumulators per lane, all indices literals;
//   * per batch of 32 points lane t evaluates the two kernel vectors of point t (FFMA2 Horner)
//     and parks them in warp-private shared memory as the inner loop consumes them: strength *
//     x-weight pairs per window column (zero outside the stencil), y weights in
//     [row mod 4][row / 4] order (zero outside);
//   * per point the warp issues 2 LDS.128 + 8 FFMA2 (spread) -- cells outside the point's stencil
//     multiply by zero -- with the same rolling reloads as the 3-D kernel.  The y weights are NOT
//     stored as (w, w) pairs: an LDS.128 costs four shared-memory wavefronts per warp whatever
//     it broadcasts, the first version (3 LDS.128 per point) ran the shared-memory pipe at 90 %
//     (profiles/r01e_rt2_*), and the pairs FFMA2 needs are made with register moves instead;
//   * the tile is added to the fine grid ONCE per subproblem with red.global.add.v2.f32.
// The tile kernels they replace (tile_kernels.cuh) do an LDS.128 + STS.128 read-modify-write of
// shared memory per touched row segment; measured 1.28 ms per 1e7-point transform at C4 (2-D
// type 1, ns = 7), 35x off the HBM bound of (4d+12) M + 8 nf bytes.
//
// Interpolation: the tile is LOADED into registers once per subproblem; per point 8 FFMA2 fold
// the x weights in, 4 more the y weights; partial results of 8 points are transposed through a
// padded shared-memory tile exactly as in the 3-D kernel.
#pragma once
#include "swr_kernels.cuh"

namespace b2n {

template <int NS> struct Rt2Cfg {
  static constexpr int CX = 2, S = 4;
  static constexpr int WX = 8 * CX, WY = 4 * S;       // 16 x 16 window
  static constexpr int H = NS / 2;
  static constexpr int BX = WX - NS + 1, BY = WY - NS + 1;
  static constexpr int PB = 32;
  static constexpr int NP = (NS + 1) / 2;
  static constexpr int KXO = 0;                        // WX pairs: spread (c.re kx, c.im kx); interp (kx, kx)
  static constexpr int KYO = 2 * WX;                   // ky[r = row & 3][s = row >> 2]
  static constexpr int ROW0 = KYO + WY;                // 48 floats
  static constexpr int ROW = ROW0 + 4;                 // stride / 4 odd
  static constexpr int WARPS = 1;  // one warp per CTA (see SwrCfg)
  static constexpr size_t spread_smem() { return (size_t)WARPS * PB * ROW * sizeof(float); }
  static constexpr size_t interp_warp_floats = PB * ROW + 2 * 16 * 33;
  static constexpr size_t interp_smem() { return (size_t)WARPS * interp_warp_floats * sizeof(float); }
  static_assert(BX >= 1 && BY >= 1, "window too small");
};

// lane t parks the weights of one point: pr4 = record, cv = strength (spread) or (1, 1) (interp)
template <int NS>
__device__ __forceinline__ void rt2_weights(const HornerTable<float> &tab, const float4 pr4, float2 cv,
                                            int xa, int ya, float *row) {
  using C = Rt2Cfg<NS>;
  constexpr int NP = C::NP;
  const float px = pr4.x, py = pr4.y;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < C::ROW0 / 4; i++) reinterpret_cast<float4 *>(row)[i] = z4;
  const int isx = window_start(px, NS), isy = window_start(py, NS);
  float kx[2 * NP], ky[2 * NP];
  if (!tab.direct) {
    const float zx = fmaf(2.f, float(isx) - px, float(NS - 1));
    const float zy = fmaf(2.f, float(isy) - py, float(NS - 1));
    const float2 zx2 = make_float2(zx, zx), zy2 = make_float2(zy, zy);
    float2 ax[NP], ay[NP];
#pragma unroll
    for (int j = 0; j < NP; j++) ax[j] = ay[j] = make_float2(tab.c[0][2 * j], tab.c[0][2 * j + 1]);
    for (int k = 1; k < tab.ncoef; k++) {
#pragma unroll
      for (int j = 0; j < NP; j++) {
        const float2 cj = make_float2(tab.c[k][2 * j], tab.c[k][2 * j + 1]);
        ax[j] = fma2(ax[j], zx2, cj);
        ay[j] = fma2(ay[j], zy2, cj);
      }
    }
#pragma unroll
    for (int j = 0; j < NP; j++) {
      kx[2 * j] = ax[j].x; kx[2 * j + 1] = ax[j].y;
      ky[2 * j] = ay[j].x; ky[2 * j + 1] = ay[j].y;
    }
  } else {
    float tx[NS], ty[NS];
    eval_kernel<float, NS>(tx, float(isx) - px, tab);
    eval_kernel<float, NS>(ty, float(isy) - py, tab);
#pragma unroll
    for (int j = 0; j < NS; j++) { kx[j] = tx[j]; ky[j] = ty[j]; }
  }
  {
    int xl = isx - xa;
    xl = xl < 0 ? 0 : (xl > C::WX - NS ? C::WX - NS : xl);
    float2 *dst = reinterpret_cast<float2 *>(row + C::KXO) + xl;
#pragma unroll
    for (int j = 0; j < NS; j++) dst[j] = mul2(cv, make_float2(kx[j], kx[j]));
  }
  {
    int yl = isy - ya;
    yl = yl < 0 ? 0 : (yl > C::WY - NS ? C::WY - NS : yl);
    float *dst = row + C::KYO;
#pragma unroll
    for (int j = 0; j < NS; j++) {
      const int iy = yl + j;
      dst[C::S * (iy & 3) + (iy >> 2)] = ky[j];
    }
  }
}

// this lane's share of one point's row: x pairs of its two cells, ky of its four rows
struct Rt2Row {
  float4 cx;  // pairs of cells 2q, 2q+1
  float4 ky;  // rows r, r+4, r+8, r+12
  __device__ __forceinline__ void load_x(const float *myx, int ro) { cx = *reinterpret_cast<const float4 *>(myx + ro); }
  __device__ __forceinline__ void load_y(const float *myy, int ro) { ky = *reinterpret_cast<const float4 *>(myy + ro); }
  __device__ __forceinline__ float2 cxp(int c) const { return c ? make_float2(cx.z, cx.w) : make_float2(cx.x, cx.y); }
  __device__ __forceinline__ float2 kyp(int s) const {
    const float k = s == 0 ? ky.x : (s == 1 ? ky.y : (s == 2 ? ky.z : ky.w));
    return make_float2(k, k);
  }
};

// ==================================================================================== SPREAD
template <int NS>
__global__ void __launch_bounds__(32 * Rt2Cfg<NS>::WARPS)
    k_rt2_spread(const SwrArgs a, const __grid_constant__ HornerTable<float> tab) {
  using C = Rt2Cfg<NS>;
  constexpr int S = C::S, CX = C::CX;
  extern __shared__ __align__(16) float swr_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int first, cnt, x0, y0;
  if (!swr_decode(a, blockIdx.x * C::WARPS + w, first, cnt, x0, y0)) return;
  float *rows = swr_smem + w * (C::PB * C::ROW);
  const float2 *cin = a.cin + (int64_t)blockIdx.y * a.M;
  float2 *fw = a.fw + (int64_t)blockIdx.y * a.nftot;

  const int r = lane >> 3, q = lane & 7;
  const int xa = x0 - C::H, ya = y0 - C::H;
  const int nf0 = a.nf[0], nf1 = a.nf[1];

  float2 acc[S][CX];
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int c = 0; c < CX; c++) acc[s][c] = make_float2(0.f, 0.f);

  Rt2Row pr;
  const float *myx = rows + C::KXO + 2 * CX * q;
  const float *myy = rows + C::KYO + S * r;
  const PtRec<float> *recp = a.rec + first + lane;
  const float4 zrec = make_float4(0.f, 0.f, 0.f, 0.f);
  auto ldc = [&](const float4 &rc, float2 &sc) {
    const int o = __float_as_int(rc.w);
    if (a.scale) sc = __ldg(a.scale + o);
    return ld_stream2(cin + o);
  };
  const float2 zero2 = make_float2(0.f, 0.f);
  float4 recA = lane < cnt ? ld_stream4(recp) : zrec;
  float4 recB = lane + C::PB < cnt ? ld_stream4(recp + C::PB) : zrec;
  float2 sA = make_float2(1.f, 0.f), sB = sA;
  float2 cA = lane < cnt ? ldc(recA, sA) : zero2;
  for (int b0 = 0; b0 < cnt; b0 += C::PB) {
    const int nb = min(C::PB, cnt - b0);
    const float4 recC = b0 + 2 * C::PB + lane < cnt ? ld_stream4(recp + b0 + 2 * C::PB) : zrec;
    const float2 cB = b0 + C::PB + lane < cnt ? ldc(recB, sB) : zero2;
    __syncwarp();
    if (lane < nb) {
      float2 cv = cA;
      if (a.scale) cv = make_float2(cA.x * sA.x - cA.y * sA.y, cA.x * sA.y + cA.y * sA.x);
      rt2_weights<NS>(tab, recA, cv, xa, ya, rows + lane * C::ROW);
    }
    __syncwarp();
    recA = recB;
    recB = recC;
    cA = cB;
    sA = sB;
    pr.load_x(myx, 0);
    pr.load_y(myy, 0);
    int ro = 0;
#pragma unroll 2
    for (int t = 0; t < nb; t++) {
      const int ron = t + 1 < nb ? ro + C::ROW : ro;
      const float2 c0 = pr.cxp(0), c1 = pr.cxp(1);
      float2 k[S];
#pragma unroll
      for (int s = 0; s < S; s++) k[s] = pr.kyp(s);
      pr.load_x(myx, ron);
      pr.load_y(myy, ron);
#pragma unroll
      for (int s = 0; s < S; s++) {
        acc[s][0] = fma2(c0, k[s], acc[s][0]);
        acc[s][1] = fma2(c1, k[s], acc[s][1]);
      }
      ro = ron;
    }
  }
  // the tile goes to the fine grid once
#pragma unroll
  for (int s = 0; s < S; s++) {
    const int gy = wrap_once(ya + 4 * s + r, nf1);
#pragma unroll
    for (int c = 0; c < CX; c++) {
      const int gx = wrap_once(xa + CX * q + c, nf0);
      if (acc[s][c].x != 0.f || acc[s][c].y != 0.f) red_add(fw + (int64_t)gy * nf0 + gx, acc[s][c]);
    }
  }
}

// ==================================================================================== SPREAD, stacked
// NT transforms that share the points (jax-finufft's vmap stacking, n_transf > 1) in ONE pass of
// the warp: the kernel vectors are evaluated and loaded once per point, only the strength differs.
// Row layout: kx[16] | ky[16] ([r][s]) | NT strengths.  Per point: LDS.64 (kx of this lane's two
// cells) + LDS.128 (ky of its four rows) + NT/2 LDS.128 (strengths, broadcast), then per transform
// 2 FMUL2 + 8 FFMA2.  32 complex accumulators per lane at NT = 4.
template <int NS, int NT> struct Rt2NtCfg {
  using B = Rt2Cfg<NS>;
  static constexpr int KXO = 0, KYO = B::WX, CSO = B::WX + B::WY;
  static constexpr int ROW0 = CSO + 2 * NT;                      // 40 floats at NT = 4
  static constexpr int ROW = (ROW0 / 4) % 2 == 1 ? ROW0 : ROW0 + 4;
  static constexpr size_t smem() { return (size_t)B::WARPS * B::PB * ROW * sizeof(float); }
};

template <int NS, int NT>
__global__ void __launch_bounds__(32 * Rt2Cfg<NS>::WARPS)
    k_rt2_spread_nt(const SwrArgs a, const __grid_constant__ HornerTable<float> tab, int ntr) {
  using C = Rt2Cfg<NS>;
  using R = Rt2NtCfg<NS, NT>;
  constexpr int S = C::S, CX = C::CX, NP = C::NP;
  static_assert(NT % 2 == 0, "strengths are loaded two at a time");
  extern __shared__ __align__(16) float swr_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int first, cnt, x0, y0;
  if (!swr_decode(a, blockIdx.x * C::WARPS + w, first, cnt, x0, y0)) return;
  float *rows = swr_smem + w * (C::PB * R::ROW);
  const int t0 = blockIdx.y * NT;                       // first transform of this pass
  const int nt = min(NT, ntr - t0);
  const float2 *cin = a.cin + (int64_t)t0 * a.M;
  float2 *fw = a.fw + (int64_t)t0 * a.nftot;

  const int r = lane >> 3, q = lane & 7;
  const int xa = x0 - C::H, ya = y0 - C::H;
  const int nf0 = a.nf[0], nf1 = a.nf[1];

  float2 acc[NT][S][CX];
#pragma unroll
  for (int t = 0; t < NT; t++)
#pragma unroll
    for (int s = 0; s < S; s++)
#pragma unroll
      for (int c = 0; c < CX; c++) acc[t][s][c] = make_float2(0.f, 0.f);

  const float *myx = rows + R::KXO + CX * q;
  const float *myy = rows + R::KYO + S * r;
  const PtRec<float> *recp = a.rec + first + lane;
  const float4 zrec = make_float4(0.f, 0.f, 0.f, 0.f);
  const float2 zero2 = make_float2(0.f, 0.f);
  // strengths of the point of record rc for the nt transforms of this pass
  auto ldc = [&](const float4 &rc, float2 (&cv)[NT]) {
    const int o = __float_as_int(rc.w);
#pragma unroll
    for (int t = 0; t < NT; t++) cv[t] = t < nt ? ld_stream2(cin + (int64_t)t * a.M + o) : zero2;
  };
  float4 recA = lane < cnt ? ld_stream4(recp) : zrec;
  float4 recB = lane + C::PB < cnt ? ld_stream4(recp + C::PB) : zrec;
  float2 cA[NT], cB[NT];
#pragma unroll
  for (int t = 0; t < NT; t++) cA[t] = cB[t] = zero2;
  if (lane < cnt) ldc(recA, cA);
  for (int b0 = 0; b0 < cnt; b0 += C::PB) {
    const int nb = min(C::PB, cnt - b0);
    const float4 recC = b0 + 2 * C::PB + lane < cnt ? ld_stream4(recp + b0 + 2 * C::PB) : zrec;
    if (b0 + C::PB + lane < cnt) ldc(recB, cB);
    __syncwarp();
    if (lane < nb) {  // kernel vectors of point b0 + lane, once for all transforms
      float *row = rows + lane * R::ROW;
      const float px = recA.x, py = recA.y;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < R::CSO / 4; i++) reinterpret_cast<float4 *>(row)[i] = z4;
      const int isx = window_start(px, NS), isy = window_start(py, NS);
      float kx[2 * NP], ky[2 * NP];
      if (!tab.direct) {
        const float zx = fmaf(2.f, float(isx) - px, float(NS - 1));
        const float zy = fmaf(2.f, float(isy) - py, float(NS - 1));
        const float2 zx2 = make_float2(zx, zx), zy2 = make_float2(zy, zy);
        float2 ax[NP], ay[NP];
#pragma unroll
        for (int j = 0; j < NP; j++) ax[j] = ay[j] = make_float2(tab.c[0][2 * j], tab.c[0][2 * j + 1]);
        for (int k = 1; k < tab.ncoef; k++) {
#pragma unroll
          for (int j = 0; j < NP; j++) {
            const float2 cj = make_float2(tab.c[k][2 * j], tab.c[k][2 * j + 1]);
            ax[j] = fma2(ax[j], zx2, cj);
            ay[j] = fma2(ay[j], zy2, cj);
          }
        }
#pragma unroll
        for (int j = 0; j < NP; j++) {
          kx[2 * j] = ax[j].x; kx[2 * j + 1] = ax[j].y;
          ky[2 * j] = ay[j].x; ky[2 * j + 1] = ay[j].y;
        }
      } else {
        float tx[NS], ty[NS];
        eval_kernel<float, NS>(tx, float(isx) - px, tab);
        eval_kernel<float, NS>(ty, float(isy) - py, tab);
#pragma unroll
        for (int j = 0; j < NS; j++) { kx[j] = tx[j]; ky[j] = ty[j]; }
      }
      int xl = isx - xa;
      xl = xl < 0 ? 0 : (xl > C::WX - NS ? C::WX - NS : xl);
      int yl = isy - ya;
      yl = yl < 0 ? 0 : (yl > C::WY - NS ? C::WY - NS : yl);
#pragma unroll
      for (int j = 0; j < NS; j++) {
        row[R::KXO + xl + j] = kx[j];
        const int iy = yl + j;
        row[R::KYO + S * (iy & 3) + (iy >> 2)] = ky[j];
      }
#pragma unroll
      for (int t = 0; t < NT; t += 2)
        *reinterpret_cast<float4 *>(row + R::CSO + 2 * t) = make_float4(cA[t].x, cA[t].y, cA[t + 1].x, cA[t + 1].y);
    }
    __syncwarp();
    recA = recB;
    recB = recC;
#pragma unroll
    for (int t = 0; t < NT; t++) cA[t] = cB[t];
    int ro = 0;
#pragma unroll 2
    for (int p = 0; p < nb; p++) {
      const float2 kx2 = *reinterpret_cast<const float2 *>(myx + ro);
      const float4 ky4 = *reinterpret_cast<const float4 *>(myy + ro);
      const float kyv[4] = {ky4.x, ky4.y, ky4.z, ky4.w};
#pragma unroll
      for (int t = 0; t < NT; t += 2) {
        const float4 c2 = *reinterpret_cast<const float4 *>(rows + ro + R::CSO + 2 * t);
        const float2 ct[2] = {make_float2(c2.x, c2.y), make_float2(c2.z, c2.w)};
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const float2 w0 = mul2(ct[u], make_float2(kx2.x, kx2.x)), w1 = mul2(ct[u], make_float2(kx2.y, kx2.y));
#pragma unroll
          for (int s = 0; s < S; s++) {
            const float2 k = make_float2(kyv[s], kyv[s]);
            acc[t + u][s][0] = fma2(w0, k, acc[t + u][s][0]);
            acc[t + u][s][1] = fma2(w1, k, acc[t + u][s][1]);
          }
        }
      }
      ro += R::ROW;
    }
  }
#pragma unroll
  for (int t = 0; t < NT; t++) {
    if (t >= nt) break;
#pragma unroll
    for (int s = 0; s < S; s++) {
      const int gy = wrap_once(ya + 4 * s + r, nf1);
#pragma unroll
      for (int c = 0; c < CX; c++) {
        const int gx = wrap_once(xa + CX * q + c, nf0);
        if (acc[t][s][c].x != 0.f || acc[t][s][c].y != 0.f)
          red_add(fw + (int64_t)t * a.nftot + (int64_t)gy * nf0 + gx, acc[t][s][c]);
      }
    }
  }
}

// ==================================================================================== SPREAD, stacked, narrow window
// Second generation of the stacked spreader (BASELINE config 4: 64 transforms sharing 1e7 points).
// k_rt2_spread_nt above inherits the 16 x 16 window of the single-transform kernel: 49 of its 256
// cells lie in a point's stencil, so per point and transform it issues 2 FMUL2 + 8 FFMA2 of which
// 19 % are useful, and it gathers every strength on its own (one 128-byte line per 8 bytes).  Here:
//   * the window is the 8 x 12 one of the 3-D kernels (bins of 2 x 6 anchor cells at ns = 7): lane
//     (r, q) owns x cell q of rows {r, r+4, r+8}; per point and transform 1 FMUL2 + 3 FFMA2;
//   * NT = 8 transforms per pass (24 complex accumulators per lane), so the kernel vectors are
//     evaluated and loaded once per 8 transforms;
//   * the strengths are first re-laid out POINT-major, [M][8] (k_pack_strengths: one coalesced
//     pass over the batch), so the gather through idx brings 64 useful bytes per point and pass,
//     and it is made by cp.async straight into shared memory, one batch ahead.
#ifndef RT2S_S
#define RT2S_S 3
#endif
#ifndef RT2S_NT
#define RT2S_NT 8
#endif
template <int NS> struct Rt2sCfg {
  static constexpr int NT = RT2S_NT;
  static constexpr int S = RT2S_S;
  static constexpr int WX = 8, WY = 4 * S;
  static constexpr int H = NS / 2;
  static constexpr int BX = WX - NS + 1, BY = WY - NS + 1;
  static constexpr int PB = 32;
  static constexpr int CB = 2;         // batches whose weight rows stay in shared memory across the passes
  static constexpr int NP = (NS + 1) / 2;
  static constexpr int KXO = 0;        // 8 x weights (zero outside the stencil)
  static constexpr int KYO = 8;        // ky[r = row & 3][row >> 2], 4 x 4 (3 used)
  static constexpr int ROW = 28;       // 24 floats + 4: stride / 4 odd
  static constexpr int CSO = CB * PB * ROW;  // strengths: [2 buffers][PB points][NT complex]
  static constexpr int SMEM = (CSO + 2 * PB * 2 * NT) * (int)sizeof(float);
  static_assert(BX >= 1 && BY >= 1, "window too small");
};

// cpack[i][t] = c[t][i], t < cstride (zero beyond nt), cstride = NT * passes: a transpose of 32
// points x (up to) 64 transforms per CTA through shared memory.  Reads are 256-byte runs of one
// transform, writes the contiguous cstride * 8 bytes of each point (16 KB per CTA at 64
// transforms).  Without the tile -- every thread writing the 64-byte pieces of its own point, 512
// bytes apart from its neighbour's -- the pack ran at 3.4 TB/s instead of 6.5 (profiles/r02t).
constexpr int PK_T = 64;  // transforms per tile
__global__ void __launch_bounds__(256) k_pack_strengths(const float2 *__restrict__ c, int64_t M, int nt, int cstride,
                                                         float2 *__restrict__ cpack) {
  __shared__ float2 tile[32][PK_T + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int np = (int)min((int64_t)32, M - i0);
  for (int tc = 0; tc < cstride; tc += PK_T) {
    const int tw = min(PK_T, cstride - tc);  // multiple of 8
    if (tc) __syncthreads();
    for (int t = w; t < tw; t += 8)
      tile[lane][t] = (tc + t < nt && lane < np) ? __ldcs(c + (int64_t)(tc + t) * M + i0 + lane) : make_float2(0.f, 0.f);
    __syncthreads();
    const int h = tw / 2;  // float4 per point
    for (int j = threadIdx.x; j < np * h; j += 256) {
      const int p = j / h, e = j - p * h;
      const float2 u = tile[p][2 * e], v = tile[p][2 * e + 1];
      *reinterpret_cast<float4 *>(cpack + (i0 + p) * cstride + tc + 2 * e) = make_float4(u.x, u.y, v.x, v.y);
    }
  }
}

// One warp per subproblem, ALL nt stacked transforms of it, NT = 8 per pass.  A bin holds ~30
// points at BASELINE config 4, so with one launch per 8 transforms a CTA lived ~8 us of which the
// first 3 were the chain of dependent loads that finds its points (subproblem -> bin -> records ->
// strengths; 26 % of all stall samples sat on the kernel's first instructions, profiles/r02t) and
// the kernel vectors were re-evaluated for every group of 8.  Here the passes run back to back in
// the same warp as one software pipeline over (pass, batch) steps: the strengths of the next step
// are in flight (cp.async) while this one is spread, the weight rows of the first CB batches are
// evaluated once and stay in shared memory, and the 24 accumulators are flushed to the pass's 8
// fine grids at the end of each pass.
template <int NS>
__global__ void __launch_bounds__(32) k_rt2s_spread(const SwrArgs a, const __grid_constant__ HornerTable<float> tab,
                                                     const float2 *__restrict__ cpack, int nt, int cstride) {
  using C = Rt2sCfg<NS>;
  constexpr int S = C::S, NP = C::NP, NT = C::NT;
  extern __shared__ __align__(16) float swr_smem[];
  const int lane = threadIdx.x;
  int first, cnt, x0, y0;
  if (!swr_decode(a, blockIdx.x, first, cnt, x0, y0)) return;
  float *rows = swr_smem;
  const unsigned sm0 = (unsigned)__cvta_generic_to_shared(swr_smem);
  const int r = lane >> 3, q = lane & 7;
  const int xa = x0 - C::H, ya = y0 - C::H;
  const int nf0 = a.nf[0], nf1 = a.nf[1];
  const int nbat = (cnt + C::PB - 1) / C::PB, npass = (nt + NT - 1) / NT;
  const bool cached = nbat <= C::CB;

  float2 acc[NT][S];
#pragma unroll
  for (int t = 0; t < NT; t++)
#pragma unroll
    for (int s = 0; s < S; s++) acc[t][s] = make_float2(0.f, 0.f);

  const PtRec<float> *recp = a.rec + first + lane;
  const float4 zrec = make_float4(0.f, 0.f, 0.f, 0.f);
  // the 64 bytes of strengths (pass `ps`) of the point of record rc -> buffer `buf`, row `prow`
  auto issue_str = [&](int buf, const float4 &rc, bool valid, int prow, int ps) {
    if (valid) {
#ifdef RT2S_NO_GATHER  // experiment: strengths from one hot row instead of through idx
      const float2 *src = cpack + (int64_t)(lane) * cstride + ps * NT;
#else
      const float2 *src = cpack + (int64_t)__float_as_int(rc.w) * cstride + ps * NT;
#endif
      const unsigned dst = sm0 + (C::CSO + (buf * C::PB + prow) * 2 * NT) * 4;
#pragma unroll
      for (int k = 0; k < NT / 2; k++)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * k), "l"(src + 2 * k) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // Row of a point inside its batch: class-major (the order inside a bin is free).  Class 0: the y
  // window ends below row slot S-1, class 2: it starts above row slot 0, class 1: anything else --
  // the loops of classes 0 and 2 do not issue the FFMA2s of the row slot they cannot touch (2 of 3
  // per transform).  The order is fixed when the strengths of the step are requested (one step
  // ahead, they land in the row of their point), so the loops below have warp-uniform trip counts.
  auto order = [&](const float4 &rc, int n, int &pos, int &n0, int &n01) {
    pos = lane; n0 = 0; n01 = n;
    if constexpr (S == 3) {
      const bool act = lane < n;
      int yl = window_start(rc.y, NS) - ya;
      yl = yl < 0 ? 0 : (yl > C::WY - NS ? C::WY - NS : yl);
      const int cls = yl + NS <= 4 * (S - 1) ? 0 : (yl >= 4 ? 2 : 1);
      const unsigned b0 = __ballot_sync(0xffffffffu, act && cls == 0);
      const unsigned b1 = __ballot_sync(0xffffffffu, act && cls == 1);
      const unsigned b2 = __ballot_sync(0xffffffffu, act && cls == 2);
      const unsigned lt = (1u << lane) - 1u;
      n0 = __popc(b0);
      n01 = n0 + __popc(b1);
      pos = cls == 0 ? __popc(b0 & lt) : (cls == 1 ? n0 + __popc(b1 & lt) : n01 + __popc(b2 & lt));
    }
  };
  // this lane's cell of row slot s in the first fine grid; + t * nftot per transform
  float2 *cell[S];
  {
    const int gx = wrap_once(xa + q, nf0);
#pragma unroll
    for (int s = 0; s < S; s++) cell[s] = a.fw + (int64_t)wrap_once(ya + 4 * s + r, nf1) * nf0 + gx;
  }
  // steps k = pass * nbat + batch; (bA, pA) = this step, (bB, pB) the next, (bC, pC) the one after
  auto next_step = [&](int &b, int &ps) { if (++b == nbat) { b = 0; ps++; } };
  auto rec_of = [&](int b, int ps) { return ps < npass && b * C::PB + lane < cnt ? ld_stream4(recp + b * C::PB) : zrec; };
  auto count_of = [&](int b, int ps) { return ps < npass ? min(C::PB, cnt - b * C::PB) : 0; };
  int bA = 0, pA = 0, bB = 0, pB = 0;
  next_step(bB, pB);
  int bC = bB, pC = pB;
  next_step(bC, pC);
  float4 recA = rec_of(bA, pA);
  float4 recB = rec_of(bB, pB);
  int posA, n0A, n01A;
  order(recA, count_of(bA, pA), posA, n0A, n01A);
  issue_str(0, recA, lane < cnt, posA, 0);
  const unsigned ax = sm0 + (C::KXO + q) * 4, ay = sm0 + (C::KYO + 4 * r) * 4;
  for (int k = 0; pA < npass; k++) {
    const int nb = count_of(bA, pA);
    const float4 recC = rec_of(bC, pC);
    int posB, n0B, n01B;
    const int nbB = count_of(bB, pB);
    order(recB, nbB, posB, n0B, n01B);
    issue_str((k + 1) & 1, recB, lane < nbB, posB, pB);
    const int rbase = cached ? bA * C::PB : 0;  // first row of this batch
    __syncwarp();
    if (lane < nb && (pA == 0 || !cached)) {  // kernel vectors of the point, once for all transforms
      float *row = rows + (rbase + posA) * C::ROW;
      const float px = recA.x, py = recA.y;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 6; i++) reinterpret_cast<float4 *>(row)[i] = z4;
      const int isx = window_start(px, NS), isy = window_start(py, NS);
      float kx[2 * NP], ky[2 * NP];
      if (!tab.direct) {
        const float zx = fmaf(2.f, float(isx) - px, float(NS - 1));
        const float zy = fmaf(2.f, float(isy) - py, float(NS - 1));
        const float2 zx2 = make_float2(zx, zx), zy2 = make_float2(zy, zy);
        float2 ax2[NP], ay2[NP];
#pragma unroll
        for (int j = 0; j < NP; j++) ax2[j] = ay2[j] = make_float2(tab.c[0][2 * j], tab.c[0][2 * j + 1]);
        for (int kk = 1; kk < tab.ncoef; kk++) {
#pragma unroll
          for (int j = 0; j < NP; j++) {
            const float2 cj = make_float2(tab.c[kk][2 * j], tab.c[kk][2 * j + 1]);
            ax2[j] = fma2(ax2[j], zx2, cj);
            ay2[j] = fma2(ay2[j], zy2, cj);
          }
        }
#pragma unroll
        for (int j = 0; j < NP; j++) {
          kx[2 * j] = ax2[j].x; kx[2 * j + 1] = ax2[j].y;
          ky[2 * j] = ay2[j].x; ky[2 * j + 1] = ay2[j].y;
        }
      } else {
        float tx[NS], ty[NS];
        eval_kernel<float, NS>(tx, float(isx) - px, tab);
        eval_kernel<float, NS>(ty, float(isy) - py, tab);
#pragma unroll
        for (int j = 0; j < NS; j++) { kx[j] = tx[j]; ky[j] = ty[j]; }
      }
      int xl = isx - xa;
      xl = xl < 0 ? 0 : (xl > C::WX - NS ? C::WX - NS : xl);
      int yl = isy - ya;
      yl = yl < 0 ? 0 : (yl > C::WY - NS ? C::WY - NS : yl);
#pragma unroll
      for (int j = 0; j < NS; j++) {
        row[C::KXO + xl + j] = kx[j];
        const int iy = yl + j;
        row[C::KYO + 4 * (iy & 3) + (iy >> 2)] = ky[j];
      }
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // this step's strengths have landed
    __syncwarp();
    const unsigned cs = sm0 + (C::CSO + (k & 1) * C::PB * 2 * NT) * 4;
    const unsigned axb = ax + rbase * C::ROW * 4, ayb = ay + rbase * C::ROW * 4;
    // all NT transforms of the points [p0, p1) of this batch; per point S FMUL (kx * ky[s], shared by
    // the transforms) and per transform one FFMA2 per row slot: acc += c_t * (kx ky[s]) with the
    // strength pair as the vector operand and the weight as the scalar-broadcast one.  The operands
    // of the NEXT point (rows are consecutive across the class loops) are loaded while this one is
    // spread: at 4 warps per scheduler nothing else covers the shared-memory latency
    // (short-scoreboard stalls 2.4 per issue without it, profiles/r02t).
    struct Pt { float kx; float4 ky; float4 c[NT / 2]; };
    auto load_pt = [&](Pt &d, int p) {
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(d.kx) : "r"(axb + p * C::ROW * 4));
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d.ky.x), "=f"(d.ky.y), "=f"(d.ky.z), "=f"(d.ky.w) : "r"(ayb + p * C::ROW * 4));
#pragma unroll
      for (int t = 0; t < NT / 2; t++)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d.c[t].x), "=f"(d.c[t].y), "=f"(d.c[t].z), "=f"(d.c[t].w) : "r"(cs + (p * 2 * NT + 4 * t) * 4));
    };
    Pt pa, pb;  // pa = the next point to spread when a loop is entered
    load_pt(pa, 0);
    auto spread_pt = [&](auto clc, const Pt &cur) {
      constexpr int CLS = decltype(clc)::value;
      constexpr int S0 = CLS == 2 ? 1 : 0, S1 = CLS == 0 ? S - 1 : S;
      const float kyv[4] = {cur.ky.x, cur.ky.y, cur.ky.z, cur.ky.w};
      float2 kxy[S];
#pragma unroll
      for (int s = S0; s < S1; s++) {
        const float w = cur.kx * kyv[s];
        kxy[s] = make_float2(w, w);
      }
#pragma unroll
      for (int t = 0; t < NT; t += 2) {
        const float2 c0 = make_float2(cur.c[t / 2].x, cur.c[t / 2].y), c1 = make_float2(cur.c[t / 2].z, cur.c[t / 2].w);
#pragma unroll
        for (int s = S0; s < S1; s++) {
          acc[t][s] = fma2(c0, kxy[s], acc[t][s]);
          acc[t + 1][s] = fma2(c1, kxy[s], acc[t + 1][s]);
        }
      }
    };
    auto run = [&](auto clc, int p0, int p1) {  // two points per trip: the register sets swap roles, no copies
      int p = p0;
#pragma unroll 1
      for (; p + 1 < p1; p += 2) {
        load_pt(pb, p + 1);
        spread_pt(clc, pa);
        load_pt(pa, min(p + 2, C::PB - 1));
        spread_pt(clc, pb);
      }
      if (p < p1) {
        load_pt(pb, min(p + 1, C::PB - 1));
        spread_pt(clc, pa);
        pa = pb;
      }
    };
    if constexpr (S == 3) {
      run(std::integral_constant<int, 0>{}, 0, n0A);
      run(std::integral_constant<int, 1>{}, n0A, n01A);
      run(std::integral_constant<int, 2>{}, n01A, nb);
    } else {
      run(std::integral_constant<int, 1>{}, 0, nb);
    }
    if (bA == nbat - 1) {  // end of pass pA: its NT fine grids take the window
      const int tn = min(NT, nt - pA * NT);
      const int64_t g0 = (int64_t)pA * NT * a.nftot;
#pragma unroll
      for (int t = 0; t < NT; t++) {
        if (t < tn) {
#pragma unroll
          for (int s = 0; s < S; s++)
#ifdef RT2S_NO_RED   // experiment: how much of the kernel is the flush
            if (acc[t][s].x == 12345.f) red_add_nz(cell[s] + g0 + (int64_t)t * a.nftot, acc[t][s]);
#else
            red_add_nz(cell[s] + g0 + (int64_t)t * a.nftot, acc[t][s]);
#endif
        }
#pragma unroll
        for (int s = 0; s < S; s++) acc[t][s] = make_float2(0.f, 0.f);
      }
    }
    __syncwarp();
    recA = recB; recB = recC;
    posA = posB; n0A = n0B; n01A = n01B;
    bA = bB; pA = pB; bB = bC; pB = pC;
    next_step(bC, pC);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ==================================================================================== INTERP
template <int NS>
__global__ void __launch_bounds__(32 * Rt2Cfg<NS>::WARPS)
    k_rt2_interp(const SwrArgs a, const __grid_constant__ HornerTable<float> tab) {
  using C = Rt2Cfg<NS>;
  constexpr int S = C::S, CX = C::CX;
  extern __shared__ __align__(16) float swr_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int first, cnt, x0, y0;
  if (!swr_decode(a, blockIdx.x * C::WARPS + w, first, cnt, x0, y0)) return;
  float *rows = swr_smem + w * C::interp_warp_floats;
  float2 *RES = reinterpret_cast<float2 *>(rows + C::PB * C::ROW);
  float2 *res_w = RES + lane;                                        // + (t & 15) * 33 per point
  const float2 *res_r = RES + (lane & 15) * 33 + (lane >> 4) * 16;  // this lane's 16 terms of a half-batch sum
  float2 *cout = a.cout + (int64_t)blockIdx.y * a.M;
  const float2 *fw = a.fw + (int64_t)blockIdx.y * a.nftot;

  const int r = lane >> 3, q = lane & 7;
  const int xa = x0 - C::H, ya = y0 - C::H;
  const int nf0 = a.nf[0], nf1 = a.nf[1];
  float2 val[S][CX];
#pragma unroll
  for (int s = 0; s < S; s++) {
    const int gy = wrap_once(ya + 4 * s + r, nf1);
#pragma unroll
    for (int c = 0; c < CX; c++) val[s][c] = __ldg(fw + (int64_t)gy * nf0 + wrap_once(xa + CX * q + c, nf0));
  }

  Rt2Row pr;
  const float *myx = rows + C::KXO + 2 * CX * q;
  const float *myy = rows + C::KYO + S * r;
  const PtRec<float> *recp = a.rec + first + lane;
  const float4 zrec = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 recA = lane < cnt ? ld_stream4(recp) : zrec;
  for (int b0 = 0; b0 < cnt; b0 += C::PB) {
    const int nb = min(C::PB, cnt - b0);
    const float4 recB = b0 + C::PB + lane < cnt ? ld_stream4(recp + b0 + C::PB) : zrec;
    const int orig = __float_as_int(recA.w);
    float2 mine = make_float2(0.f, 0.f);
    __syncwarp();
    if (lane < nb) rt2_weights<NS>(tab, recA, make_float2(1.f, 1.f), xa, ya, rows + lane * C::ROW);
    __syncwarp();
    recA = recB;
    pr.load_x(myx, 0);
    pr.load_y(myy, 0);
    int ro = 0, t = 0;
    for (int half = 0; half * 16 < nb; half++) {
      const int tend = min(nb, 16 * half + 16);
      float2 *rw = res_w;
#pragma unroll 2
      for (; t < tend; t++) {
        const int ron = t + 1 < nb ? ro + C::ROW : ro;
        const float2 k0 = pr.cxp(0), k1 = pr.cxp(1);  // (kx, kx) of this lane's two cells
        float2 k[S];
#pragma unroll
        for (int s = 0; s < S; s++) k[s] = pr.kyp(s);
        pr.load_x(myx, ron);
        pr.load_y(myy, ron);
        float2 res = make_float2(0.f, 0.f);
#pragma unroll
        for (int s = 0; s < S; s++) {
          const float2 t0 = fma2(val[s][1], k1, mul2(val[s][0], k0));
          res = fma2(t0, k[s], res);
        }
        *rw = res;
        rw += 33;
        ro = ron;
      }
      // lane (row, part) sums half of row `row`; one butterfly step finishes it (swr_kernels.cuh)
      __syncwarp();
      {
        float2 s0 = res_r[0];
#pragma unroll
        for (int j = 1; j < 16; j++) s0 = add2(s0, res_r[j]);
        s0.x += __shfl_xor_sync(0xffffffffu, s0.x, 16);
        s0.y += __shfl_xor_sync(0xffffffffu, s0.y, 16);
        if ((lane >> 4) == half) mine = s0;
      }
      __syncwarp();
    }
    if (lane < nb) {
      float2 o = mine;
      if (a.scale) {
        const float2 sc = __ldg(a.scale + orig);
        o = make_float2(o.x * sc.x - o.y * sc.y, o.x * sc.y + o.y * sc.x);
      }
      cout[orig] = o;
    }
  }
}

// ==================================================================================== INTERP, stacked
// NT transforms sharing the points: NT register tiles per lane, the kernel vectors of a point are
// loaded once, the partial results of every transform go through their own padded tile and are
// reduced once per half batch.  Row layout as in the single-transform kernel ((kx, kx) pairs |
// ky), written by rt2_weights with cv = (1, 1).
template <int NS, int NT> struct Rt2NtiCfg {
  using B = Rt2Cfg<NS>;
  static constexpr size_t warp_floats = B::PB * B::ROW + (size_t)NT * 2 * 8 * 33;  // NT tiles of 8 points
  static constexpr size_t smem() { return B::WARPS * warp_floats * sizeof(float); }
};

template <int NS, int NT>
__global__ void __launch_bounds__(32 * Rt2Cfg<NS>::WARPS)
    k_rt2_interp_nt(const SwrArgs a, const __grid_constant__ HornerTable<float> tab, int ntr) {
  using C = Rt2Cfg<NS>;
  using R = Rt2NtiCfg<NS, NT>;
  constexpr int S = C::S, CX = C::CX;
  extern __shared__ __align__(16) float swr_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int first, cnt, x0, y0;
  if (!swr_decode(a, blockIdx.x * C::WARPS + w, first, cnt, x0, y0)) return;
  float *rows = swr_smem + w * R::warp_floats;
  float2 *RES = reinterpret_cast<float2 *>(rows + C::PB * C::ROW);   // [NT][8][33]
  float2 *res_w = RES + lane;
  const float2 *res_r = RES + (lane & 7) * 33 + (lane >> 3) * 8;  // this lane's 8 terms of a group sum
  const int t0 = blockIdx.y * NT;
  const int nt = min(NT, ntr - t0);
  float2 *cout = a.cout + (int64_t)t0 * a.M;
  const float2 *fw = a.fw + (int64_t)t0 * a.nftot;

  const int r = lane >> 3, q = lane & 7;
  const int xa = x0 - C::H, ya = y0 - C::H;
  const int nf0 = a.nf[0], nf1 = a.nf[1];
  float2 val[NT][S][CX];
#pragma unroll
  for (int t = 0; t < NT; t++)
#pragma unroll
    for (int s = 0; s < S; s++) {
      const int gy = wrap_once(ya + 4 * s + r, nf1);
#pragma unroll
      for (int c = 0; c < CX; c++)
        val[t][s][c] = t < nt ? __ldg(fw + (int64_t)t * a.nftot + (int64_t)gy * nf0 + wrap_once(xa + CX * q + c, nf0))
                              : make_float2(0.f, 0.f);
    }

  Rt2Row pr;
  const float *myx = rows + C::KXO + 2 * CX * q;
  const float *myy = rows + C::KYO + S * r;
  const PtRec<float> *recp = a.rec + first + lane;
  const float4 zrec = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 recA = lane < cnt ? ld_stream4(recp) : zrec;
  for (int b0 = 0; b0 < cnt; b0 += C::PB) {
    const int nb = min(C::PB, cnt - b0);
    const float4 recB = b0 + C::PB + lane < cnt ? ld_stream4(recp + b0 + C::PB) : zrec;
    const int orig = __float_as_int(recA.w);
    float2 mine[NT];
#pragma unroll
    for (int t = 0; t < NT; t++) mine[t] = make_float2(0.f, 0.f);
    __syncwarp();
    if (lane < nb) rt2_weights<NS>(tab, recA, make_float2(1.f, 1.f), xa, ya, rows + lane * C::ROW);
    __syncwarp();
    recA = recB;
    pr.load_x(myx, 0);
    pr.load_y(myy, 0);
    int ro = 0, p = 0;
    for (int grp = 0; grp * 8 < nb; grp++) {  // groups of 8 points: small result tiles (shared memory per warp)
      const int tend = min(nb, 8 * grp + 8);
      float2 *rw = res_w;
      for (; p < tend; p++) {
        const int ron = p + 1 < nb ? ro + C::ROW : ro;
        const float2 k0 = pr.cxp(0), k1 = pr.cxp(1);
        float2 k[S];
#pragma unroll
        for (int s = 0; s < S; s++) k[s] = pr.kyp(s);
        pr.load_x(myx, ron);
        pr.load_y(myy, ron);
#pragma unroll
        for (int t = 0; t < NT; t++) {
          float2 res = make_float2(0.f, 0.f);
#pragma unroll
          for (int s = 0; s < S; s++) {
            const float2 u = fma2(val[t][s][1], k1, mul2(val[t][s][0], k0));
            res = fma2(u, k[s], res);
          }
          rw[t * (8 * 33)] = res;
        }
        rw += 33;
        ro = ron;
      }
      __syncwarp();
#pragma unroll
      for (int t = 0; t < NT; t++) {
        float2 s0 = res_r[t * (8 * 33)];
#pragma unroll
        for (int j = 1; j < 8; j++) s0 = add2(s0, res_r[t * (8 * 33) + j]);
        s0.x += __shfl_xor_sync(0xffffffffu, s0.x, 8);
        s0.y += __shfl_xor_sync(0xffffffffu, s0.y, 8);
        s0.x += __shfl_xor_sync(0xffffffffu, s0.x, 16);
        s0.y += __shfl_xor_sync(0xffffffffu, s0.y, 16);
        if ((lane >> 3) == grp) mine[t] = s0;
      }
      __syncwarp();
    }
    if (lane < nb) {
#pragma unroll
      for (int t = 0; t < NT; t++)
        if (t < nt) cout[(int64_t)t * a.M + orig] = mine[t];
    }
  }
}

}  // namespace b2n

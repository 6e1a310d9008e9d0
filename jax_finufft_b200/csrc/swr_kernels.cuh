// swr_kernels.cuh -- sliding-window REGISTER spreading / interpolation, 3-D, float, ns <= 8.
// Replaces spread_3d_subprob / spread_3d_output_driven and interp_3d_nupts_driven / interp_3d_subprob
// (V/src/cuda/3d/spreadinterp3d.cuh:138-383, 556-712) on the headline path (3-D c64, eps >= 1e-7).
//
// The reference accumulates a subproblem in a shared-memory tile with 2*ns^3 float atomics per
// point; our tile kernels (tile_kernels.cuh) replaced the atomics by owned LDS/STS read-modify-
// writes but remain bound by shared-memory wavefronts (~70 per point).  Here the accumulators
// live in REGISTERS and the FP32 pipe is the binding unit, so the design goal is to waste as few
// lane-FMAs as possible on cells a point does not touch:
//   * points are binned by the ANCHOR cell u_d = window_start(x_d) + ns/2 (the cell whose
//     ns-wide stencil the point uses), so all points of an anchor cell touch the same ns cells
//     per dimension and a bin of b anchor cells touches exactly b + ns - 1 cells;
//   * a subproblem is a run of <= maxsub points of one bin of BX x BY x BZ anchor cells
//     (2 x 6 x 64 at ns = 7), sorted by anchor z (sort.cu sub-key); ONE WARP owns it;
//   * the warp covers the WX x WY = 8 x 12 cells those points can touch: lane (r, q) owns x cell
//     q (CX = 1; two cells for ns = 8) of rows {r, 4+r, 8+r} ("row slots") and, for each, a ring
//     of D = ns z planes -- 3 * ns complex accumulators per lane (42 registers at ns = 7);
//   * points arrive in anchor-z order, so the ring only moves forward: a plane that drops out
//     of the window is flushed with one red.global.add.v2.f32 per row slot (8 lanes = one 64-byte
//     row segment, executed in L2) and its registers are zeroed.  Ring slot = plane mod D, so no
//     register ever moves;
//   * per batch of 32 points, lane t evaluates the three kernel vectors of point t once and
//     parks them in (warp-private) shared memory in exactly the form the inner loop consumes:
//     strength * x-weight as a complex pair per window column, y weights in (row mod 4, row / 4)
//     order, z weights duplicated (w, w) in ring-slot order;
//   * per point the warp then issues 6 LDS and, per row slot the point's y window reaches (2.33
//     of 3 on average), 2 FMUL + ns fma.rn.f32x2 (FFMA2: one instruction per complex cell
//     update).  Lane efficiency 7/8 (x) * 7/(4 * 2.33) (y) * 7/7 (z) = 66 % at ns = 7.
// No shared or global atomics with return, no block barriers (warps are independent).
//
// Interpolation is the mirror image: the ring holds planes LOADED from the fine grid (64-byte
// row segments), each point costs (ns + 3) instructions per reachable row slot and one STS; the
// cross-lane sum is done once per batch through shared memory.
//
// Algorithmic HBM bytes: 16 (record) + 8 (strength, 32-B sector gather) per
// point + one pass over the fine grid.  Binding resource: FP32 pipe (FFMA2 = 2 pipe cycles).
#pragma once
#include "plan.h"

namespace b2n {

constexpr int SWR_BZ = 64;  // anchor z cells per bin (subproblems slide along them)

template <int NS> struct SwrCfg {
  static constexpr int D = NS;               // ring depth = z planes an anchor cell's points touch
  static constexpr int H = NS / 2;           // anchor u = window_start + H
  static constexpr int CX = NS <= 7 ? 1 : 2; // x cells per lane
  static constexpr int WX = 8 * CX;          // window extent in x
  static constexpr int BX = CX == 1 ? WX - NS + 1 : ((WX - NS + 1) & ~1);  // bin extent (anchor cells)
  static constexpr int S = 3;                // row slots per lane
  static constexpr int WY = 4 * S;           // window extent in y
  static constexpr int BY = WY - NS + 1;
  static constexpr int BZ = SWR_BZ;
  static constexpr int PB = 32;              // points per weight batch (one per lane)
  // per-point shared-memory row, in floats
  static constexpr int KXO = 0;              // WX pairs: spread (c.re*kx, c.im*kx); interp (kx, kx)
  static constexpr int KYO = 2 * WX;         // ky[row & 3][row >> 2], 4 x 4
  static constexpr int KZO = KYO + 16;       // D pairs (kz, kz) in ring-slot order, then META
  static constexpr int KZW = (2 * D + 2 + 3) & ~3;
  static constexpr int MTO = KZO + KZW - 2;  // META = {first plane of the point's window, row-slot mask}
  static constexpr int ROW0 = KZO + KZW;
  // stride/4 odd: lane-strided 16-byte accesses of 8 consecutive lanes hit 8 distinct bank groups
  static constexpr int ROW = (ROW0 / 4) % 2 == 1 ? ROW0 : ROW0 + 4;
  static constexpr int WARPS = 4;
  static constexpr int OFF = D * 1024;       // makes (plane + OFF) % D non-negative
  static constexpr int MINB = NS <= 5 ? 5 : (NS <= 7 ? 4 : 2);  // CTAs per SM the register budget allows
  static constexpr size_t smem_bytes() { return (size_t)WARPS * PB * ROW * sizeof(float); }
  static_assert(BX >= 1 && BY >= 1, "window too small");
  static_assert(CX == 1 || (BX % 2 == 0 && H % 2 == 0), "paired x cells must stay 16-byte aligned");
};

struct SwrArgs {
  const PtRec<float> *rec;
  const int32_t *bin_start, *sp_off, *sp_bin;
  const float2 *cin;     // spread: strengths [ntr][M]
  float2 *cout;          // interp: outputs   [ntr][M]
  const float2 *scale;   // optional per-point factor (type-3 prephase / deconv), by original index
  float2 *fw;            // fine grid(s) [ntr][nftot]
  int64_t M, nftot;
  int nf[3], bin[3], nbin[3];
  int64_t nbins;
  int maxsub;
};

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
// acc += a * b, two packed floats at once (sm_100: FFMA2)
__device__ __forceinline__ void fma2(unsigned long long &acc, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// streaming loads: the records and the strength gather are read once and must not push the
// fine-grid lines (the REDs' / ring loads' working set) out of L2
__device__ __forceinline__ float4 ld_stream4(const void *p) {
  return __ldcs(reinterpret_cast<const float4 *>(p));
}
__device__ __forceinline__ float2 ld_stream2(const float2 *p) { return __ldcs(p); }

// ---- anchor cells (shared with sort.cu) ------------------------------------------------------------
// u = window_start(x') + ns/2 in [0, nf]; u == nf (x' within ns/2 - floor(ns/2) .. of the seam)
// is the periodic image of u = 0: the coordinate is stored shifted by -nf so that the kernel
// recomputes the same window from the record alone.
__device__ __forceinline__ int swr_anchor(float &xr, int ns, int nf) {
  const int h = ns >> 1;
  int u = window_start(xr, ns) + h;
  if (u >= nf) {
    xr -= (float)nf;
    u = window_start(xr, ns) + h;
  }
  return u < 0 ? 0 : (u >= nf ? nf - 1 : u);
}

// ---- subproblem decode ---------------------------------------------------------------------------
__device__ __forceinline__ bool swr_decode(const SwrArgs &a, int sp, int &first, int &cnt, int &x0,
                                           int &y0) {
  if (sp >= a.sp_off[a.nbins]) return false;
  const int b = a.sp_bin[sp];
  const int s = sp - a.sp_off[b];
  first = a.bin_start[b] + s * a.maxsub;
  cnt = min(a.maxsub, a.bin_start[b + 1] - first);
  x0 = (b % a.nbin[0]) * a.bin[0];
  y0 = ((b / a.nbin[0]) % a.nbin[1]) * a.bin[1];
  return cnt > 0;
}

// ---- phase 1: lane t parks the weights of point p (see SwrCfg for the row layout) ----------------
template <int NS, bool SPREAD>
__device__ __forceinline__ void swr_weights(const HornerTable<float> &tab, const float4 pr4, float2 cv,
                                            int xa, int ya, float *row) {
  // pr4 = the point record; cv = strength (spread) or (1, 1) (interp: the x row holds (kx, kx))
  using C = SwrCfg<NS>;
  const float px = pr4.x, py = pr4.y, pz = pr4.z;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < (C::KYO + 16) / 4; i++) reinterpret_cast<float4 *>(row)[i] = z4;
  // the three kernel vectors in one Horner sweep (each table coefficient is fetched once)
  const int isx = window_start(px, NS), isy = window_start(py, NS), isz = window_start(pz, NS);
  float kx[NS], ky[NS], kz[NS];
  if (!tab.direct) {
    const float zx = fmaf(2.f, float(isx) - px, float(NS - 1));
    const float zy = fmaf(2.f, float(isy) - py, float(NS - 1));
    const float zz = fmaf(2.f, float(isz) - pz, float(NS - 1));
#pragma unroll
    for (int j = 0; j < NS; j++) kx[j] = ky[j] = kz[j] = tab.c[0][j];
    for (int k = 1; k < tab.ncoef; k++) {
#pragma unroll
      for (int j = 0; j < NS; j++) {
        const float cj = tab.c[k][j];
        kx[j] = fmaf(kx[j], zx, cj);
        ky[j] = fmaf(ky[j], zy, cj);
        kz[j] = fmaf(kz[j], zz, cj);
      }
    }
  } else {
    eval_kernel<float, NS>(kx, float(isx) - px, tab);
    eval_kernel<float, NS>(ky, float(isy) - py, tab);
    eval_kernel<float, NS>(kz, float(isz) - pz, tab);
  }
  {
    int xl = isx - xa;
    xl = xl < 0 ? 0 : (xl > C::WX - NS ? C::WX - NS : xl);
    float2 *dst = reinterpret_cast<float2 *>(row + C::KXO) + xl;
#pragma unroll
    for (int j = 0; j < NS; j++) dst[j] = make_float2(cv.x * kx[j], cv.y * kx[j]);
  }
  int mask;
  {
    int yl = isy - ya;
    yl = yl < 0 ? 0 : (yl > C::WY - NS ? C::WY - NS : yl);
#pragma unroll
    for (int j = 0; j < NS; j++) {
      const int iy = yl + j;
      row[C::KYO + 4 * (iy & 3) + (iy >> 2)] = ky[j];
    }
    const int slo = yl >> 2, shi = (yl + NS - 1) >> 2;
    mask = ((2 << shi) - 1) & ~((1 << slo) - 1);
  }
  {
    int slot = (isz + C::OFF) % C::D;
    float2 *dst = reinterpret_cast<float2 *>(row + C::KZO);
#pragma unroll
    for (int j = 0; j < NS; j++) {
      dst[slot] = make_float2(kz[j], kz[j]);
      slot = slot + 1 == C::D ? 0 : slot + 1;
    }
    *reinterpret_cast<int2 *>(row + C::MTO) = make_int2(isz, mask);
  }
}

// z weights (duplicated pairs, ring-slot order) + META of point row `row`
template <int NS>
__device__ __forceinline__ void swr_load_kz(const float *row, unsigned long long (&kzp)[SwrCfg<NS>::D],
                                            int &zw, int &mask) {
  using C = SwrCfg<NS>;
  float4 v[C::KZW / 4];
#pragma unroll
  for (int i = 0; i < C::KZW / 4; i++) v[i] = *reinterpret_cast<const float4 *>(row + C::KZO + 4 * i);
#pragma unroll
  for (int k = 0; k < C::D; k++)
    kzp[k] = (k & 1) ? pack2(v[k / 2].z, v[k / 2].w) : pack2(v[k / 2].x, v[k / 2].y);
  const float4 m = v[C::KZW / 4 - 1];
  zw = __float_as_int(m.z);
  mask = __float_as_int(m.w);
}

// ==================================================================================== SPREAD
template <int NS>
__global__ void __launch_bounds__(32 * SwrCfg<NS>::WARPS, SwrCfg<NS>::MINB)
    k_swr_spread(const SwrArgs a, const __grid_constant__ HornerTable<float> tab) {
  using C = SwrCfg<NS>;
  constexpr int D = C::D, S = C::S, CX = C::CX;
  extern __shared__ __align__(16) float swr_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int first, cnt, x0, y0;
  if (!swr_decode(a, blockIdx.x * C::WARPS + w, first, cnt, x0, y0)) return;
  float *rows = swr_smem + w * (C::PB * C::ROW);
  const float2 *cin = a.cin + (int64_t)blockIdx.y * a.M;
  float2 *fw = a.fw + (int64_t)blockIdx.y * a.nftot;

  const int r = lane >> 3, q = lane & 7;
  const int xa = x0 - C::H, ya = y0 - C::H;
  const int nf0 = a.nf[0], nf1 = a.nf[1], nf2 = a.nf[2];
  int rowoff[S];
#pragma unroll
  for (int s = 0; s < S; s++)
    rowoff[s] = wrap_once(ya + 4 * s + r, nf1) * nf0 + wrap_once(xa + CX * q, nf0);
  const int64_t pstride = (int64_t)nf0 * nf1;

  unsigned long long acc[S][CX][D];
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int c = 0; c < CX; c++)
#pragma unroll
      for (int k = 0; k < D; k++) acc[s][c][k] = 0ull;

  // flush ring slot `slot` (holding plane p) to the fine grid and clear it
  auto flush = [&](int p, int slot) {
    const int gz = p < 0 ? p + nf2 : (p >= nf2 ? p - nf2 : p);
    float2 *pl = fw + (int64_t)gz * pstride;
#pragma unroll
    for (int k = 0; k < D; k++) {
      if (slot == k) {
#pragma unroll
        for (int s = 0; s < S; s++) {
          if constexpr (CX == 1) {
            if (acc[s][0][k] & 0x7fffffff7fffffffull) red_add(pl + rowoff[s], unpack2(acc[s][0][k]));
          } else {
            if ((acc[s][0][k] | acc[s][1][k]) & 0x7fffffff7fffffffull) {
              const float2 v0 = unpack2(acc[s][0][k]), v1 = unpack2(acc[s][1][k]);
              red_add4(reinterpret_cast<float4 *>(pl + rowoff[s]), make_float4(v0.x, v0.y, v1.x, v1.y));
            }
          }
#pragma unroll
          for (int c = 0; c < CX; c++) acc[s][c][k] = 0ull;
        }
      }
    }
  };

  int cur = 0x40000000;  // first plane held by the ring (sentinel: ring empty)
  const float *myx = rows + C::KXO + 2 * CX * q;
  const float *myy = rows + C::KYO + 4 * r;
  // Two-deep software pipeline over batches of 32 points: the record of batch b+2 and the
  // strength of batch b+1 (whose address comes from the record of b+1) are in flight while the
  // warp spreads batch b, so neither DRAM round trip is exposed.
  const PtRec<float> *recp = a.rec + first + lane;
  const float4 zrec = make_float4(0.f, 0.f, 0.f, 0.f);
  auto ldc = [&](const float4 &rc) {
    const int o = __float_as_int(rc.w);
    float2 cv = ld_stream2(cin + o);
    if (a.scale) {
      const float2 sc = __ldg(a.scale + o);
      cv = make_float2(cv.x * sc.x - cv.y * sc.y, cv.x * sc.y + cv.y * sc.x);
    }
    return cv;
  };
  float4 recA = lane < cnt ? ld_stream4(recp) : zrec;
  float4 recB = lane + C::PB < cnt ? ld_stream4(recp + C::PB) : zrec;
  float2 cA = lane < cnt ? ldc(recA) : make_float2(0.f, 0.f);
  for (int b0 = 0; b0 < cnt; b0 += C::PB) {
    const int nb = min(C::PB, cnt - b0);
    const float4 recC = b0 + 2 * C::PB + lane < cnt ? ld_stream4(recp + b0 + 2 * C::PB) : zrec;
    const float2 cB = b0 + C::PB + lane < cnt ? ldc(recB) : make_float2(0.f, 0.f);
    __syncwarp();
    if (lane < nb) swr_weights<NS, true>(tab, recA, cA, xa, ya, rows + lane * C::ROW);
    __syncwarp();
    recA = recB;
    recB = recC;
    cA = cB;
#pragma unroll 1
    for (int t = 0; t < nb; t++) {
      const int ro = t * C::ROW;
      unsigned long long kzp[D];
      int zw, mask;
      swr_load_kz<NS>(rows + ro, kzp, zw, mask);
      if (zw != cur) {
        if (cur != 0x40000000) {
          const int nfl = (zw > cur && zw < cur + D) ? zw - cur : D;  // out of order = full flush
          int slot = (cur + C::OFF) % D;
          for (int i = 0; i < nfl; i++) {
            flush(cur + i, slot);
            slot = slot + 1 == D ? 0 : slot + 1;
          }
        }
        cur = zw;
      }
      float2 cx[CX];
      if constexpr (CX == 1) {
        cx[0] = *reinterpret_cast<const float2 *>(myx + ro);
      } else {
        const float4 v = *reinterpret_cast<const float4 *>(myx + ro);
        cx[0] = make_float2(v.x, v.y);
        cx[1] = make_float2(v.z, v.w);
      }
      const float4 ky4 = *reinterpret_cast<const float4 *>(myy + ro);
      const float ky[4] = {ky4.x, ky4.y, ky4.z, ky4.w};
      auto do_slot = [&](int s) {
#pragma unroll
        for (int c = 0; c < CX; c++) {
          const unsigned long long wv = pack2(cx[c].x * ky[s], cx[c].y * ky[s]);
#pragma unroll
          for (int k = 0; k < D; k++) fma2(acc[s][c][k], wv, kzp[k]);
        }
      };
      // ns >= 5 rows of a 12-row window always reach the middle row slot
      if (mask & 1) do_slot(0);
      if (NS >= 5 || (mask & 2)) do_slot(1);
      if (mask & 4) do_slot(2);
    }
  }
  if (cur != 0x40000000) {
    int slot = (cur + C::OFF) % D;
    for (int i = 0; i < D; i++) {
      flush(cur + i, slot);
      slot = slot + 1 == D ? 0 : slot + 1;
    }
  }
}

// ==================================================================================== INTERP
// Per-warp shared memory: the weight rows + RES[8][32] float2 partial results of a group of 8
// points (point t, lane j at column (j + t) & 31: conflict-free for the per-point store and for
// the per-group sums).
template <int NS> struct SwrInterpSmem {
  static constexpr size_t warp_floats = SwrCfg<NS>::PB * SwrCfg<NS>::ROW + 2 * 8 * 32;
  static constexpr size_t bytes() { return SwrCfg<NS>::WARPS * warp_floats * sizeof(float); }
};

template <int NS>
__global__ void __launch_bounds__(32 * SwrCfg<NS>::WARPS, SwrCfg<NS>::MINB)
    k_swr_interp(const SwrArgs a, const __grid_constant__ HornerTable<float> tab) {
  using C = SwrCfg<NS>;
  constexpr int D = C::D, S = C::S, CX = C::CX;
  extern __shared__ __align__(16) float swr_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int first, cnt, x0, y0;
  if (!swr_decode(a, blockIdx.x * C::WARPS + w, first, cnt, x0, y0)) return;
  float *rows = swr_smem + w * SwrInterpSmem<NS>::warp_floats;
  unsigned long long *RES = reinterpret_cast<unsigned long long *>(rows + C::PB * C::ROW);
  float2 *cout = a.cout + (int64_t)blockIdx.y * a.M;
  const float2 *fw = a.fw + (int64_t)blockIdx.y * a.nftot;

  const int r = lane >> 3, q = lane & 7;
  const int xa = x0 - C::H, ya = y0 - C::H;
  const int nf0 = a.nf[0], nf1 = a.nf[1], nf2 = a.nf[2];
  int rowoff[S];
#pragma unroll
  for (int s = 0; s < S; s++)
    rowoff[s] = wrap_once(ya + 4 * s + r, nf1) * nf0 + wrap_once(xa + CX * q, nf0);
  const int64_t pstride = (int64_t)nf0 * nf1;

  unsigned long long val[S][CX][D];
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int c = 0; c < CX; c++)
#pragma unroll
      for (int k = 0; k < D; k++) val[s][c][k] = 0ull;

  const float2 *ldp[S];  // this lane's cell of each row slot, plane 0
#pragma unroll
  for (int s = 0; s < S; s++) ldp[s] = fw + rowoff[s];
  // fetch this lane's cells of plane p
  auto fetch = [&](int p, unsigned long long (&v)[S][CX]) {
    const int gz = p < 0 ? p + nf2 : (p >= nf2 ? p - nf2 : p);
    const int64_t po = (int64_t)gz * pstride;
#pragma unroll
    for (int s = 0; s < S; s++) {
      if constexpr (CX == 1) {
        const float2 t = __ldg(ldp[s] + po);
        v[s][0] = pack2(t.x, t.y);
      } else {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(ldp[s] + po));
        v[s][0] = pack2(t.x, t.y);
        v[s][1] = pack2(t.z, t.w);
      }
    }
  };
  auto put = [&](int slot, const unsigned long long (&v)[S][CX]) {
#pragma unroll
    for (int k = 0; k < D; k++) {
      if (slot == k) {
#pragma unroll
        for (int s = 0; s < S; s++)
#pragma unroll
          for (int c = 0; c < CX; c++) val[s][c][k] = v[s][c];
      }
    }
  };
  unsigned long long pre[S][CX];  // plane pre_p, fetched one ring step ahead
  int pre_p = 0x40000000;

  int cur = 0x40000000;  // first plane held by the ring (sentinel: ring empty)
  const float *myx = rows + C::KXO + 2 * CX * q;
  const float *myy = rows + C::KYO + 4 * r;
  const PtRec<float> *recp = a.rec + first + lane;
  const float4 zrec = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 recA = lane < cnt ? ld_stream4(recp) : zrec;
  for (int b0 = 0; b0 < cnt; b0 += C::PB) {
    const int nb = min(C::PB, cnt - b0);
    const float4 recB = b0 + C::PB + lane < cnt ? ld_stream4(recp + b0 + C::PB) : zrec;
    const int orig = __float_as_int(recA.w);
    unsigned long long mine = 0ull;  // interpolated value of point b0 + lane
    __syncwarp();
    if (lane < nb) swr_weights<NS, false>(tab, recA, make_float2(1.f, 1.f), xa, ya, rows + lane * C::ROW);
    __syncwarp();
    recA = recB;
#pragma unroll 1
    for (int t = 0; t < nb; t++) {
      const int ro = t * C::ROW;
      unsigned long long kzp[D];
      int zw, mask;
      swr_load_kz<NS>(rows + ro, kzp, zw, mask);
      if (zw != cur) {
        int pfirst = zw, n = D;
        if (zw > cur && zw < cur + D) {  // (never true for the empty-ring sentinel)
          pfirst = cur + D;
          n = zw - cur;
        }
        int slot = (pfirst + C::OFF) % D;
        for (int i = 0; i < n; i++) {
          if (pfirst + i == pre_p) {
            put(slot, pre);
          } else {
            unsigned long long v[S][CX];
            fetch(pfirst + i, v);
            put(slot, v);
          }
          slot = slot + 1 == D ? 0 : slot + 1;
        }
        cur = zw;
        pre_p = zw + D;
        fetch(pre_p, pre);
      }
      float2 kx[CX];  // (kx, kx)
      if constexpr (CX == 1) {
        kx[0] = *reinterpret_cast<const float2 *>(myx + ro);
      } else {
        const float4 v = *reinterpret_cast<const float4 *>(myx + ro);
        kx[0] = make_float2(v.x, v.y);
        kx[1] = make_float2(v.z, v.w);
      }
      const float4 ky4 = *reinterpret_cast<const float4 *>(myy + ro);
      const float ky[4] = {ky4.x, ky4.y, ky4.z, ky4.w};
      unsigned long long res = 0ull;
      auto do_slot = [&](int s) {
#pragma unroll
        for (int c = 0; c < CX; c++) {
          unsigned long long t0 = mul2(val[s][c][0], kzp[0]);
#pragma unroll
          for (int k = 1; k < D; k++) fma2(t0, val[s][c][k], kzp[k]);
          fma2(res, t0, pack2(kx[c].x * ky[s], kx[c].y * ky[s]));
        }
      };
      if (mask & 1) do_slot(0);
      if (NS >= 5 || (mask & 2)) do_slot(1);
      if (mask & 4) do_slot(2);
      RES[(t & 7) * 32 + ((lane + t) & 31)] = res;
      if ((t & 7) == 7 || t == nb - 1) {
        // group of <= 8 points done: lane j sums quarter j>>3 of row j&7, two butterfly steps
        // finish the row; point 8g + row lives in lane 8g + row, which keeps the total
        __syncwarp();
        const int row8 = lane & 7, part = lane >> 3;
        unsigned long long s0 = 0ull;
#pragma unroll
        for (int j = 0; j < 8; j++) s0 = add2(s0, RES[row8 * 32 + ((part * 8 + j + row8) & 31)]);
        s0 = add2(s0, __shfl_xor_sync(0xffffffffu, s0, 8));
        s0 = add2(s0, __shfl_xor_sync(0xffffffffu, s0, 16));
        if (part == (t >> 3)) mine = s0;
        __syncwarp();
      }
    }
    if (lane < nb) {
      float2 o = unpack2(mine);
      if (a.scale) {
        const float2 sc = __ldg(a.scale + orig);
        o = make_float2(o.x * sc.x - o.y * sc.y, o.x * sc.y + o.y * sc.x);
      }
      cout[orig] = o;
    }
  }
}

}  // namespace b2n

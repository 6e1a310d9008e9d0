// swr_kernels.cuh -- sliding-window REGISTER kernels, 3-D, float, ns <= 8: the shared geometry
// (SwrCfg, anchors, batch ordering) and the INTERPOLATOR k_swr_interp; the spreader is
// k_swr2_spread (swr2_kernels.cuh).  The description below covers both; the spreader's phase-chain
// first generation (k_swr_spread, round 1: 7.34 ms at C3) was replaced in round 2.
// Replaces spread_3d_subprob / spread_3d_output_driven and interp_3d_nupts_driven / interp_3d_subprob
// (V/src/cuda/3d/spreadinterp3d.cuh:138-383, 556-712) on the headline path (3-D c64, eps >= 1e-7).
//
// The reference accumulates a subproblem in a shared-memory tile with 2*ns^3 float atomics per
// point; our tile kernels (tile_kernels.cuh) replaced the atomics by owned LDS/STS read-modify-
// writes but remain bound by shared-memory wavefronts (~70 per point).  Here the accumulators
// live in REGISTERS and instruction issue / the FP32 pipe are the binding units:
//   * points are binned by the ANCHOR cell u_d = window_start(x_d) + ns/2 (the cell whose
//     ns-wide stencil the point uses), so all points of an anchor cell touch the same ns cells
//     per dimension and a bin of b anchor cells touches exactly b + ns - 1 cells;
//   * a subproblem is a run of <= maxsub points of one bin of BX x BY x BZ anchor cells
//     (2 x 6 x 64 at ns = 7), sorted by anchor z (sort.cu sub-key); ONE WARP owns it;
//   * the warp covers the WX x WY = 8 x 12 cells those points can touch: lane (r, q) owns x cell
//     q (CX = 1; two cells for ns = 8) of rows {r, 4+r, 8+r} ("row slots") and, for each, a ring
//     of D = ns z planes -- 3 * ns complex accumulators per lane (42 registers at ns = 7);
//   * points arrive in anchor-z order, so the ring only moves forward: the plane that drops out
//     of the window is retired with one red.global.add.v2.f32 per row slot (8 lanes = one 64-byte
//     row segment, executed in L2) and its registers are zeroed;
//   * STATIC PHASES.  Ring slot of plane p = (p - base) mod D.  The per-subproblem control flow
//     is a chain of D copies of { consume the points whose window starts at plane cur; retire
//     plane cur; cur++ }, entered through one switch per batch of 32 points and falling through
//     from phase to phase, so every accumulator index in the FMA block and in the retire step is
//     a literal: no select chains, no slot rotation of the z weights.  (The first version kept
//     one copy of the loop and selected the slot to flush at run time: ~100 executed
//     instructions per retired plane, 17 % of the kernel -- profiles/r01e_*.)
//   * per batch of 32 points, lane t evaluates the three kernel vectors of point t once (Horner,
//     two intervals per FFMA2) and parks them in warp-private shared memory in the form the
//     inner loop consumes: strength * x-weight as a complex pair per window column, y weights in
//     (row mod 4, row / 4) order, z weights in plane order, stored ONCE each: FFMA2 / FMUL2 take
//     a scalar-broadcast operand (R.F32), so no (w, w) pairs are needed -- and an LDS.128 costs
//     four shared-memory wavefronts per warp whatever it broadcasts;
//   * per point the warp then issues 4 LDS, 3 FMUL2 and 3 * ns FFMA2 (one instruction per complex
//     cell update), straight-line: row slots a point's y window does not reach multiply by zero
//     weights instead of branching or predicating (both measured slower).  Each
//     piece of a point's weights is reloaded from the NEXT point's row right after its last use
//     ("rolling" loads), so the loads have a whole FMA block to land;
//   * one warp per CTA: warps are independent, and a CTA slot is recycled the moment its
//     subproblem ends.
// No shared or global atomics with return, no block barriers (warps are independent).
//
// Interpolation is the mirror image: the ring holds planes LOADED from the fine grid (64-byte
// row segments, the next plane prefetched one step ahead); the cross-lane sum is done once per
// half batch (16 points) through a padded shared-memory tile, outside the phase chain.
//
// FFMA2/FMUL2 are emitted through the sm_100 intrinsics (__ffma2_rn): with inline-asm "+l"
// operands ptxas routed half of the accumulators through temporaries (20 MOVs per point).
//
// Algorithmic HBM bytes: 16 (record) + 8 (strength, 32-B sector gather) per
// point + one pass over the fine grid.
#pragma once
#include <type_traits>

#include "plan.h"

namespace b2n {

#ifndef SWR_S
#define SWR_S 3
#endif
// 1: class-specialised spread loops (points of a batch ordered by (plane, y class); the row slot a
// class never touches is not computed): type-1 step 12.46 -> 12.26 ms at C3.  0 rebuilds the plain chain.
#ifndef SWR_YCLASS
#define SWR_YCLASS 1
#endif
#ifndef SWR_YCLASS_INTERP
#define SWR_YCLASS_INTERP 0
#endif
#ifndef SWR_BZ_N
#define SWR_BZ_N 64
#endif
constexpr int SWR_BZ = SWR_BZ_N;  // anchor z cells per bin (subproblems slide along them)
constexpr int SWR_EMPTY = 0x40000000;

template <int NS> struct SwrCfg {
  static constexpr int D = NS;               // ring depth = z planes an anchor cell's points touch
  static constexpr int H = NS / 2;           // anchor u = window_start + H
  static constexpr int CX = NS <= 7 ? 1 : 2; // x cells per lane
  static constexpr int WX = 8 * CX;          // window extent in x
  static constexpr int BX = CX == 1 ? WX - NS + 1 : ((WX - NS + 1) & ~1);  // bin extent (anchor cells)
  static constexpr int S = SWR_S;            // row slots per lane
  static constexpr int WY = 4 * S;           // window extent in y
  static constexpr int BY = WY - NS + 1;
  static constexpr int BZ = SWR_BZ;
  static constexpr int PB = 32;              // points per weight batch (one per lane)
  static constexpr int NP = (NS + 1) / 2;    // Horner interval pairs
  // per-point shared-memory row, in floats
  static constexpr int KXO = 0;              // WX pairs: spread (c.re*kx, c.im*kx); interp (kx, kx)
  static constexpr int KYO = 2 * WX;         // ky[row & 3][row >> 2], 4 x 4
  static constexpr int KZO = KYO + 16;       // D z weights in plane order (FFMA2 broadcasts a scalar operand), META last
  static constexpr int KZW = (D + 1 + 3) & ~3;
  static constexpr int MTO = KZO + KZW - 1;  // META = first plane of the point's window
  static constexpr int ROW0 = KZO + KZW;
  // stride/4 odd: lane-strided 16-byte accesses of 8 consecutive lanes hit 8 distinct bank groups
  static constexpr int ROW = (ROW0 / 4) % 2 == 1 ? ROW0 : ROW0 + 4;
  // one warp per CTA: warps are independent, and a CTA's slot is then recycled the moment its
  // subproblem ends instead of waiting for the slowest of four (measured at C3: interp 9.85 ->
  // 9.0 ms, spread -0.7 %; two warps per CTA gave nothing)
  static constexpr int WARPS = 1;
#ifdef SWR_MINB
  static constexpr int MINB = SWR_MINB;
#else
  static constexpr int MINB = (NS <= 5 ? 5 : (NS <= 7 ? 4 : 2)) * 4;
#endif  // CTAs per SM the register budget allows
  static constexpr size_t smem_bytes() { return (size_t)WARPS * PB * ROW * sizeof(float); }
  static_assert(BX >= 1 && BY >= 1, "window too small");
  static_assert(CX == 1 || (BX % 2 == 0 && H % 2 == 0), "paired x cells must stay 16-byte aligned");
  static_assert(D <= 8, "the phase chain below is written out for D <= 8");
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

// packed FP32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2, one issue slot for two lanes' worth)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }

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
// pr4 = the point record; cv = strength (spread) or (1, 1) (interp: the x row holds (kx, kx))
template <int NS>
__device__ __forceinline__ void swr_weights(const HornerTable<float> &tab, const float4 pr4, float2 cv,
                                            int xa, int ya, float *row, int meta_shift = 0, int meta_cls = 0) {
  using C = SwrCfg<NS>;
  constexpr int NP = C::NP;
  const float px = pr4.x, py = pr4.y, pz = pr4.z;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < (C::KYO + 16) / 4; i++) reinterpret_cast<float4 *>(row)[i] = z4;
  const int isx = window_start(px, NS), isy = window_start(py, NS), isz = window_start(pz, NS);
  float kx[2 * NP], ky[2 * NP], kz[2 * NP];
  if (!tab.direct) {
    // the three kernel vectors in one Horner sweep, two intervals per FFMA2; each coefficient
    // pair is fetched once (warp-uniform constant-bank load).  Column NS of an odd-NS table is 0.
    const float zx = fmaf(2.f, float(isx) - px, float(NS - 1));
    const float zy = fmaf(2.f, float(isy) - py, float(NS - 1));
    const float zz = fmaf(2.f, float(isz) - pz, float(NS - 1));
    const float2 zx2 = make_float2(zx, zx), zy2 = make_float2(zy, zy), zz2 = make_float2(zz, zz);
    float2 ax[NP], ay[NP], az[NP];
#pragma unroll
    for (int j = 0; j < NP; j++) ax[j] = ay[j] = az[j] = make_float2(tab.c[0][2 * j], tab.c[0][2 * j + 1]);
    for (int k = 1; k < tab.ncoef; k++) {
#pragma unroll
      for (int j = 0; j < NP; j++) {
        const float2 cj = make_float2(tab.c[k][2 * j], tab.c[k][2 * j + 1]);
        ax[j] = fma2(ax[j], zx2, cj);
        ay[j] = fma2(ay[j], zy2, cj);
        az[j] = fma2(az[j], zz2, cj);
      }
    }
#pragma unroll
    for (int j = 0; j < NP; j++) {
      kx[2 * j] = ax[j].x; kx[2 * j + 1] = ax[j].y;
      ky[2 * j] = ay[j].x; ky[2 * j + 1] = ay[j].y;
      kz[2 * j] = az[j].x; kz[2 * j + 1] = az[j].y;
    }
  } else {
    float tx[NS], ty[NS], tz[NS];
    eval_kernel<float, NS>(tx, float(isx) - px, tab);
    eval_kernel<float, NS>(ty, float(isy) - py, tab);
    eval_kernel<float, NS>(tz, float(isz) - pz, tab);
#pragma unroll
    for (int j = 0; j < NS; j++) { kx[j] = tx[j]; ky[j] = ty[j]; kz[j] = tz[j]; }
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
#pragma unroll
    for (int j = 0; j < NS; j++) {
      const int iy = yl + j;
      row[C::KYO + 4 * (iy & 3) + (iy >> 2)] = ky[j];
    }
  }
  {
    float kzm[C::KZW];
#pragma unroll
    for (int j = 0; j < C::KZW; j++) kzm[j] = j < NS ? kz[j] : 0.f;
    kzm[C::KZW - 1] = __int_as_float((isz << meta_shift) + meta_cls);
#pragma unroll
    for (int i = 0; i < C::KZW / 4; i++)
      *reinterpret_cast<float4 *>(row + C::KZO + 4 * i) = make_float4(kzm[4 * i], kzm[4 * i + 1], kzm[4 * i + 2], kzm[4 * i + 3]);
  }
}

// Position and y class of lane's point inside its batch of nb points, for the class-specialised
// loops: the batch arrives ordered by window plane, so only runs of equal planes are permuted
// (class 0 first).  Class 0: the y window ends below row slot S-1; class 2: it starts above row
// slot 0; class 1: anything.  A batch that is not ordered by plane (periodic seam) keeps its order
// and is all class 1.
template <int NS>
__device__ __forceinline__ void swr_batch_order(const float4 &rec, int nb, int ya, int lane, int &pos, int &cls) {
  using C = SwrCfg<NS>;
  pos = lane;
  cls = 1;
  if constexpr (C::S == 3) {
    const bool act = lane < nb;
    const int isz = act ? window_start(rec.z, NS) : 0x3fffffff - lane;
    int yl = window_start(rec.y, NS) - ya;
    yl = yl < 0 ? 0 : (yl > C::WY - NS ? C::WY - NS : yl);
    const int prev = __shfl_up_sync(0xffffffffu, isz, 1);
    const bool sorted = !__any_sync(0xffffffffu, act && lane > 0 && isz < prev);
    if (sorted) {
      cls = yl + NS <= 4 * (C::S - 1) ? 0 : (yl >= 4 ? 2 : 1);
      const unsigned peers = __match_any_sync(0xffffffffu, isz);
      const unsigned b0 = __ballot_sync(0xffffffffu, act && cls == 0) & peers;
      const unsigned b1 = __ballot_sync(0xffffffffu, act && cls == 1) & peers;
      const unsigned b2 = __ballot_sync(0xffffffffu, act && cls == 2) & peers;
      const unsigned lt = (1u << lane) - 1u;
      const int rank = cls == 0 ? __popc(b0 & lt)
                                : (cls == 1 ? __popc(b0) + __popc(b1 & lt) : __popc(b0) + __popc(b1) + __popc(b2 & lt));
      pos = __ffs(peers) - 1 + rank;
    }
  }
}

// One point's weights as the inner loop holds them in registers: this lane's x weight(s), its
// y weights, the (kz, kz) pairs in plane order and the window's first plane (META).  The loops
// below keep ONE such set live and reload each piece from the NEXT point's row right after its
// last use ("rolling" loads): every LDS has a whole FMA block to land, a point's row is read
// exactly once even across phase changes, and no second register set is needed.
template <int NS> struct SwrRow {
  using C = SwrCfg<NS>;
  static constexpr int NV = C::KZW / 4;
  float4 kv[NV];       // kz 4i .. 4i+3 (the last one ends with META)
  float2 cx[C::CX];
  float4 ky;
  __device__ __forceinline__ void load_xy(const float *myx, const float *myy, int ro) {
    if constexpr (C::CX == 1) {
      cx[0] = *reinterpret_cast<const float2 *>(myx + ro);
    } else {
      const float4 v = *reinterpret_cast<const float4 *>(myx + ro);
      cx[0] = make_float2(v.x, v.y);
      cx[1] = make_float2(v.z, v.w);
    }
    ky = *reinterpret_cast<const float4 *>(myy + ro);
  }
  __device__ __forceinline__ void load_kv(const float *rows, int ro, int i) {
    kv[i] = *reinterpret_cast<const float4 *>(rows + ro + C::KZO + 4 * i);
  }
  __device__ __forceinline__ float2 kz(int j) const {  // (kz, kz): a scalar-broadcast FFMA2 operand
    const float4 v = kv[j / 4];
    const float k = (j & 3) == 0 ? v.x : ((j & 3) == 1 ? v.y : ((j & 3) == 2 ? v.z : v.w));
    return make_float2(k, k);
  }
  __device__ __forceinline__ int zw() const { return __float_as_int(kv[NV - 1].w); }
  __device__ __forceinline__ float2 kyv(int s) const {
    const float k = s == 0 ? ky.x : (s == 1 ? ky.y : (s == 2 ? ky.z : ky.w));
    return make_float2(k, k);
  }
};

// ==================================================================================== INTERP
// Per-warp shared memory: the weight rows + RES[16][33] float2 partial results of half a batch
// (point t, lane j at [t & 15][j]; the odd row stride makes both the per-point store and the
// per-half sums -- lane (row, part) adds columns part*16 .. part*16+15 of its row -- free of bank
// conflicts with constant per-lane offsets).  The half-batch boundary is treated like a batch
// boundary of the phase chain, so the inner loop carries no "group full?" test.
// + the plane staging ring (SWR_STAGE): SWR_STG planes of [row slot][lane] float2 (x CX), filled by
// cp.async (LDGSTS) SWR_STG ring steps ahead of their use.  The first version kept ONE plane of
// look-ahead in registers (`pre`) and waited on it: long-scoreboard stalls ~1 per issue, 15 % of
// all samples on the register moves that consume it (profiles/r01h) -- a plane is consumed every
// ~2800 cycles, less than the latency of a fine-grid line under this kernel's own random 32-byte
// output scatter.  Bulk/TMA copies (cp.async.bulk[.tensor]) need 16-byte aligned rows; a window
// row starts at an odd cell (x0 - ns/2: 8-byte aligned) and wraps periodically, so the per-lane
// 8-byte LDGSTS form is the one that fits.  Each lane reads back only what it copied itself:
// cp.async.wait_group is all the synchronisation needed.
#ifndef SWR_STAGE
#define SWR_STAGE 1
#endif
#ifndef SWR_STG
#define SWR_STG 2  // planes in flight: 2 -> 8.60 ms, 4 -> 8.72, 6 -> 8.82 at C3 (more stages = less L1 for the plane lines)
#endif
template <int NS> struct SwrInterpSmem {
  static constexpr size_t res_floats = SwrCfg<NS>::PB * SwrCfg<NS>::ROW + 2 * 16 * 33;  // multiple of 4 floats: rows stay 16-byte aligned
  static constexpr size_t stg_floats = SWR_STAGE ? (size_t)SWR_STG * SwrCfg<NS>::S * 32 * 2 * SwrCfg<NS>::CX : 0;
  static constexpr size_t warp_floats = res_floats + stg_floats;
  static constexpr size_t bytes() { return SwrCfg<NS>::WARPS * warp_floats * sizeof(float); }
};

__device__ __forceinline__ unsigned swr_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void swr_cp_async8(unsigned d, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(g) : "memory");
}
__device__ __forceinline__ void swr_cp_async16(unsigned d, const void *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
}
__device__ __forceinline__ void swr_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void swr_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

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
  float2 *RES = reinterpret_cast<float2 *>(rows + C::PB * C::ROW);
  float2 *res_w = RES + lane;                                        // + (t & 15) * 33 per point
  const float2 *res_r = RES + (lane & 15) * 33 + (lane >> 4) * 16;  // this lane's 16 terms of a half-batch sum
  float2 *cout = a.cout + (int64_t)blockIdx.y * a.M;

  const int r = lane >> 3, q = lane & 7;
  const int xa = x0 - C::H, ya = y0 - C::H;
  const int nf0 = a.nf[0], nf1 = a.nf[1], nf2 = a.nf[2];
  const int64_t pstride = (int64_t)nf0 * nf1;
  const float2 *cell[S];  // this lane's cell of each row slot in plane 0
#pragma unroll
  for (int s = 0; s < S; s++)
    cell[s] = a.fw + (int64_t)blockIdx.y * a.nftot +
              (wrap_once(ya + 4 * s + r, nf1) * nf0 + wrap_once(xa + CX * q, nf0));

  int cur = SWR_EMPTY;  // first plane held by the ring
  int ph = 0;           // ring slot of plane cur
  // last plane a point of this bin can touch (anchor z0 + BZ - 1, window start - H, D planes)
  const int plast = (a.sp_bin[blockIdx.x * C::WARPS + w] / (a.nbin[0] * a.nbin[1])) * a.bin[2] + C::BZ - 1 - C::H + D - 1;
  float2 val[S][CX][D];  // ring of loaded planes
  auto plane_off = [&](int p) {
    int gz = p < 0 ? p + nf2 : (p >= nf2 ? p - nf2 : p);
    if (gz >= nf2) gz %= nf2;  // look-ahead beyond one period (tiny grids)
    return (int64_t)gz * pstride;
  };
  auto fetch = [&](int p, float2 (&v)[S][CX]) {
    const int64_t po = plane_off(p);
#pragma unroll
    for (int s = 0; s < S; s++) {
      if constexpr (CX == 1) {
        v[s][0] = __ldg(cell[s] + po);
      } else {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(cell[s] + po));
        v[s][0] = make_float2(t.x, t.y);
        v[s][1] = make_float2(t.z, t.w);
      }
    }
  };
#if SWR_STAGE
  constexpr unsigned STG_PLANE = S * 32 * CX * 8;  // bytes per stage
  const unsigned stg0 = swr_smem_u32(rows + SwrInterpSmem<NS>::res_floats) + lane * CX * 8;
  unsigned sga = stg0;  // stage that holds plane cur + D
  // copy this lane's cells of plane p into stage g (one commit group per plane, possibly empty:
  // nothing is fetched beyond the last plane a point of this bin can touch)
  auto stage_issue = [&](unsigned g, int p) {
    if (p <= plast) {
      const int64_t po = plane_off(p);
#pragma unroll
      for (int s = 0; s < S; s++) {
        if constexpr (CX == 1) swr_cp_async8(g + s * 32 * CX * 8, cell[s] + po);
        else swr_cp_async16(g + s * 32 * CX * 8, cell[s] + po);
      }
    }
    swr_cp_async_commit();
  };
  // ring slot PH <- the oldest staged plane (= plane cur + D); its stage is re-issued SWR_STG planes on
  auto stage_take = [&](auto phc) {
    constexpr int PH = decltype(phc)::value;
    swr_cp_async_wait<SWR_STG - 1>();
#pragma unroll
    for (int s = 0; s < S; s++) {
      if constexpr (CX == 1) {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(sga + s * 32 * CX * 8));
        val[s][0][PH] = v;
      } else {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sga + s * 32 * CX * 8));
        val[s][0][PH] = make_float2(v.x, v.y);
        val[s][1][PH] = make_float2(v.z, v.w);
      }
    }
    stage_issue(sga, cur + D + SWR_STG);
    sga = sga + STG_PLANE == stg0 + SWR_STG * STG_PLANE ? stg0 : sga + STG_PLANE;
  };
#else
  float2 pre[S][CX];     // plane cur + D, fetched one ring step ahead
#endif
  // (re)load the whole ring for a window starting at plane p0; phases re-based so that ph = 0
  auto refill = [&](int p0) {
#if SWR_STAGE
    swr_cp_async_wait<0>();
#endif
#pragma unroll
    for (int k = 0; k < D; k++) {
      float2 v[S][CX];
      fetch(p0 + k, v);
#pragma unroll
      for (int s = 0; s < S; s++)
#pragma unroll
        for (int c = 0; c < CX; c++) val[s][c][k] = v[s][c];
    }
#if SWR_STAGE
#pragma unroll
    for (int g = 0; g < SWR_STG; g++) stage_issue(stg0 + g * STG_PLANE, p0 + D + g);
    sga = stg0;
#else
    fetch(p0 + D, pre);
#endif
  };
  // interpolated value (this lane's share) of the point held in `pr`, then roll `pr` on to the
  // row at `ron`; PH = ring slot of the first plane of the point's window
  SwrRow<NS> pr;
  const float *myx = rows + C::KXO + 2 * CX * q;
  const float *myy = rows + C::KYO + 4 * r;
  auto point = [&](auto phc, auto clc, int ron) {
    constexpr int PH = decltype(phc)::value;
    constexpr int CLS = decltype(clc)::value;  // see swr_batch_order: the row slot a class never touches is skipped
    constexpr int S0 = CLS == 2 ? 1 : 0, S1 = CLS == 0 ? S - 1 : S;
    // value = kx * sum_s ky[s] * (sum_k plane[s][k] * kz[k]): the x and y weights enter AFTER the
    // plane sums as scalar-broadcast operands (S FMUL2/FFMA2 + one FMUL2 per cell column) instead
    // of as S precomputed (kx * ky[s]) pairs (S FMUL2 more per point, held across the FMA block),
    // so their rolling load moves behind the final dot product.
    float2 part[S][CX];
#pragma unroll
    for (int i = 0; i < SwrRow<NS>::NV; i++) {
#pragma unroll
      for (int j = 4 * i; j < 4 * i + 4; j++) {
        if (j < D) {
          const float2 kzj = pr.kz(j);
#pragma unroll
          for (int s = S0; s < S1; s++)
#pragma unroll
            for (int c = 0; c < CX; c++)
              part[s][c] = j == 0 ? mul2(val[s][c][(PH + j) % D], kzj)
                                  : fma2(val[s][c][(PH + j) % D], kzj, part[s][c]);
        }
      }
      pr.load_kv(rows, ron, i);
    }
    float2 res;
#pragma unroll
    for (int c = 0; c < CX; c++) {
      float2 pc = mul2(part[S0][c], pr.kyv(S0));
#pragma unroll
      for (int s = S0 + 1; s < S1; s++) pc = fma2(part[s][c], pr.kyv(s), pc);
      const float2 kxc = make_float2(pr.cx[c].x, pr.cx[c].x);
      res = c == 0 ? mul2(pc, kxc) : fma2(pc, kxc, res);
    }
    pr.load_xy(myx, myy, ron);
    return res;
  };

  const PtRec<float> *recp = a.rec + first + lane;
  const float4 zrec = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 recA = lane < cnt ? ld_stream4(recp) : zrec;
  for (int b0 = 0; b0 < cnt; b0 += C::PB) {
    const int nb = min(C::PB, cnt - b0);
    const float4 recB = b0 + C::PB + lane < cnt ? ld_stream4(recp + b0 + C::PB) : zrec;
    float2 mine = make_float2(0.f, 0.f);  // interpolated value of the point at batch position `lane`
    __syncwarp();
#if SWR_YCLASS_INTERP
    int pos, cls;
    swr_batch_order<NS>(recA, nb, ya, lane, pos, cls);
    // original index by batch POSITION, parked in the pad column of RES (never read or written by the sums)
    int *opad = reinterpret_cast<int *>(RES + 32);
    if (lane < nb) {
      opad[(pos & 15) * 66 + (pos >> 4)] = __float_as_int(recA.w);
      swr_weights<NS>(tab, recA, make_float2(1.f, 1.f), xa, ya, rows + pos * C::ROW, 2, cls);
    }
    __syncwarp();
    const int orig = lane < nb ? opad[(lane & 15) * 66 + (lane >> 4)] : 0;
#else
    const int orig = __float_as_int(recA.w);
    if (lane < nb) swr_weights<NS>(tab, recA, make_float2(1.f, 1.f), xa, ya, rows + lane * C::ROW);
    __syncwarp();
#endif
    recA = recB;
    int t = 0, ro = 0;
    pr.load_xy(myx, myy, 0);
#pragma unroll
    for (int i = 0; i < SwrRow<NS>::NV; i++) pr.load_kv(rows, 0, i);
#if SWR_YCLASS_INTERP
#define SWR_ZP(m) ((m) >> 2) /* META = plane * 4 + class */
#define SWR_INTERP_POINTS(PH)                                                                   \
      SWR_INTERP_CLASS(PH, 0)                                                                   \
      SWR_INTERP_CLASS(PH, 1)                                                                   \
      SWR_INTERP_CLASS(PH, 2)
#define SWR_INTERP_CLASS(PH, CL)                                                                \
      while (t < tend && zw == 4 * cur + CL) {                                                  \
        const int ron = t + 1 < nb ? ro + C::ROW : ro;                                          \
        *rw = point(std::integral_constant<int, PH>{}, std::integral_constant<int, CL>{}, ron); \
        rw += 33;                                                                               \
        t++;                                                                                    \
        ro = ron;                                                                               \
        zw = pr.zw();                                                                           \
      }
#else
#define SWR_ZP(m) (m)
#define SWR_INTERP_POINTS(PH)                                                                   \
      while (t < tend && zw == cur) {                                                           \
        const int ron = t + 1 < nb ? ro + C::ROW : ro;                                          \
        *rw = point(std::integral_constant<int, PH>{}, std::integral_constant<int, 1>{}, ron);  \
        rw += 33;                                                                               \
        t++;                                                                                    \
        ro = ron;                                                                               \
        zw = pr.zw();                                                                           \
      }
#endif
    int zw = pr.zw();
    bool fill = cur == SWR_EMPTY;
    for (int half = 0; half * 16 < nb; half++) {
      const int tend = min(nb, 16 * half + 16);
      float2 *rw = res_w;
    reenter:
      if (fill) {  // first point, or a gap wider than the ring (or disorder): one shared copy
        cur = SWR_ZP(zw);
        refill(cur);
        ph = 0;
        fill = false;
      }
      for (;;) {
        switch (ph) {
#if SWR_STAGE
#define SWR_INTERP_ADVANCE(PH)                                                                  \
      stage_take(std::integral_constant<int, PH>{});                                            \
      cur++;
#else
#define SWR_INTERP_ADVANCE(PH)                                                                  \
      _Pragma("unroll") for (int s = 0; s < S; s++)                                             \
          _Pragma("unroll") for (int c = 0; c < CX; c++) val[s][c][PH] = pre[s][c];             \
      cur++;                                                                                    \
      fetch(cur + D, pre);
#endif
#define SWR_INTERP_PHASE(PH)                                                                    \
  case PH:                                                                                      \
    if constexpr (PH < D) {                                                                     \
      SWR_INTERP_POINTS(PH)                                                                     \
      if (t >= tend) {                                                                          \
        ph = PH;                                                                                \
        goto half_done;                                                                         \
      }                                                                                         \
      if ((unsigned)(SWR_ZP(zw) - cur) >= (unsigned)D || SWR_ZP(zw) + D - 1 > plast) {          \
        fill = true;                                                                            \
        goto reenter;                                                                           \
      }                                                                                         \
      /* advance one plane: slot PH takes plane cur + D, the next one is requested */           \
      SWR_INTERP_ADVANCE(PH)                                                                    \
    }
          SWR_INTERP_PHASE(0)
          SWR_INTERP_PHASE(1)
          SWR_INTERP_PHASE(2)
          SWR_INTERP_PHASE(3)
          SWR_INTERP_PHASE(4)
          SWR_INTERP_PHASE(5)
          SWR_INTERP_PHASE(6)
          SWR_INTERP_PHASE(7)
#undef SWR_INTERP_PHASE
#undef SWR_INTERP_ADVANCE
#undef SWR_INTERP_POINTS
#undef SWR_INTERP_CLASS
#undef SWR_ZP
          default: break;
        }
        ph = 0;
      }
    half_done:
      // lane (row, part) sums half of row `row`; one butterfly step finishes it.  Point
      // 16 * half + row belongs to lane 16 * half + row = the lane with part == half.
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
#if SWR_STAGE
  swr_cp_async_wait<0>();
#endif
}

}  // namespace b2n

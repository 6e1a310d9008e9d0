// swr_kernels.cuh -- sliding-window REGISTER spreading / interpolation, 3-D, float, ns <= 8.
// Replaces spread_3d_subprob / spread_3d_output_driven and interp_3d_nupts_driven / interp_3d_subprob
// (V/src/cuda/3d/spreadinterp3d.cuh:138-383, 556-712) on the headline path (3-D c64, eps >= 1e-7).
//
// The reference accumulates a subproblem in a shared-memory tile with 2*ns^3 float atomics per
// point; our tile kernels (tile_kernels.cuh) replaced the atomics by owned LDS/STS read-modify-
// writes but remain bound by shared-memory wavefronts (~70 per point).  Here the accumulators
// live in REGISTERS:
//   * a subproblem is a run of <= maxsub points of one bin of BX x BY x BZ = 8 x (12-ns) x 64
//     fine-grid cells, sorted by z cell (sort.cu sub-key); ONE WARP owns it;
//   * the warp covers the 16 (x) x 12 (y) cells any of those points can touch; lane (r, q) owns
//     x cells {2q, 2q+1} of rows {r, 4+r, 8+r} and, for each, a ring of D = ns+1 z planes:
//     6 * D complex accumulators per lane (96 registers at ns = 7);
//   * points arrive in z-cell order, so the ring only moves forward: the plane that drops out
//     of the window is flushed with red.global.add.v4.f32 (two complex cells per lane, executed
//     in L2) and its registers are zeroed.  Ring slot = plane mod D, so no register ever moves;
//   * per batch of 32 points, lane t evaluates the three kernel vectors of point t once and
//     parks them zero-padded in (warp-private) shared memory: x weights at their window offset,
//     y weights in (row mod 4, row / 4) order, z weights in ring-slot order;
//   * per point the warp then issues 5 LDS and <= 6*D fma.rn.f32x2 (FFMA2: one instruction per
//     complex cell update, the real z weight broadcast as the scalar operand).  Row slots the
//     point's y window cannot reach are skipped (warp-uniform branch).
// No shared or global atomics with return, no block barriers (warps are independent).
//
// Interpolation is the mirror image: the ring holds planes LOADED from the fine grid (coalesced
// 128-byte rows), each point costs 6*(D+1) FFMA2 and a warp reduction.
//
// Algorithmic HBM bytes: 16 (record) + 8 (strength, 32-B sector gather) per
// point + one pass over the fine grid.  Binding resource: FMA issue.
#pragma once
#include "plan.h"

namespace b2n {

template <int NS> struct SwrCfg {
  static constexpr int D = NS + 1;           // ring depth (z planes a z cell's points can touch)
  static constexpr int H = NS / 2;           // first plane of cell zc's window: zc - H
  static constexpr int HXE = (H + 1) & ~1;   // low-side x halo, even (16-byte aligned pairs)
  static constexpr int BX = 8, WX = 16;      // bin / window extent in x
  static constexpr int S = 3;                // row slots per lane
  static constexpr int WY = 4 * S;           // window extent in y
  static constexpr int BY = WY - NS;         // bin extent in y
  static constexpr int BZ = 64;              // bin extent in z (subproblems slide along it)
  static constexpr int PB = 32;              // points per weight batch (one per lane)
  static constexpr int ROW = 20;             // words per point per array: 16 used + 4 pad makes
                                             // lane-strided STS.128 / LDS.128 conflict-free
  static constexpr int WARPS = 4;
  static constexpr int OFF = D * 64;         // makes (plane + OFF) % D non-negative
  static constexpr size_t smem_bytes() { return (size_t)WARPS * 3 * PB * ROW * sizeof(float); }
  static_assert(BX + NS - H + HXE <= WX, "x window too narrow");
  static_assert(D <= 12, "ring too deep for the 12-word kz row");
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

// ---- phase 1: lane t parks the weights of point p0 + t --------------------------------------------
// KX[t][i]            x weight of window column i (0 outside the point's ns columns)
// KY[t][4*(i&3)+i/4]  y weight of window row i
// KZ[t][slot]         z weight of the plane living in ring slot `slot`;  KZ[t][12..15] = META
// META = {zw (first plane of the point's z cell window), row-slot mask, strength.re, strength.im}
template <int NS, bool SPREAD>
__device__ __forceinline__ void swr_weights(const SwrArgs &a, const HornerTable<float> &tab, int p,
                                            int xa, int ya, const float2 *cin, float *KXt,
                                            float *KYt, float *KZt, int &orig) {
  using C = SwrCfg<NS>;
  const PtRec<float> pr = a.rec[p];
  orig = pr.idx;
  float2 cv = make_float2(1.f, 0.f);
  if (SPREAD) {
    cv = cin[pr.idx];
    if (a.scale) {
      const float2 sc = a.scale[pr.idx];
      cv = make_float2(cv.x * sc.x - cv.y * sc.y, cv.x * sc.y + cv.y * sc.x);
    }
  }
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    reinterpret_cast<float4 *>(KXt)[i] = z4;
    reinterpret_cast<float4 *>(KYt)[i] = z4;
    if (i < 3) reinterpret_cast<float4 *>(KZt)[i] = z4;
  }
  float ker[NS];
  {
    const int is = window_start(pr.x, NS);
    eval_kernel<float, NS>(ker, float(is) - pr.x, tab);
    const int xl = is - xa;
#pragma unroll
    for (int j = 0; j < NS; j++) KXt[xl + j] = ker[j];
  }
  int mask;
  {
    const int is = window_start(pr.y, NS);
    eval_kernel<float, NS>(ker, float(is) - pr.y, tab);
    const int yl = is - ya;
#pragma unroll
    for (int j = 0; j < NS; j++) {
      const int iy = yl + j;
      KYt[4 * (iy & 3) + (iy >> 2)] = ker[j];
    }
    const int slo = yl >> 2, shi = (yl + NS - 1) >> 2;
    mask = ((2 << shi) - 1) & ~((1 << slo) - 1);
  }
  {
    const int is = window_start(pr.z, NS);
    eval_kernel<float, NS>(ker, float(is) - pr.z, tab);
    int slot = (is + C::OFF) % C::D;
#pragma unroll
    for (int j = 0; j < NS; j++) {
      KZt[slot] = ker[j];
      slot = slot + 1 == C::D ? 0 : slot + 1;
    }
    float4 m;
    m.x = __int_as_float((int)pr.z - C::H);
    m.y = __int_as_float(mask);
    m.z = cv.x;
    m.w = cv.y;
    reinterpret_cast<float4 *>(KZt)[3] = m;
  }
}

// ==================================================================================== SPREAD
template <int NS>
__global__ void __launch_bounds__(128, 3) k_swr_spread(const SwrArgs a,
                                                       const __grid_constant__ HornerTable<float> tab) {
  using C = SwrCfg<NS>;
  constexpr int D = C::D, S = C::S;
  extern __shared__ __align__(16) float swr_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int first, cnt, x0, y0;
  if (!swr_decode(a, blockIdx.x * C::WARPS + w, first, cnt, x0, y0)) return;
  float *KX = swr_smem + w * (3 * C::PB * C::ROW);
  float *KY = KX + C::PB * C::ROW;
  float *KZ = KY + C::PB * C::ROW;
  const float2 *cin = a.cin + (int64_t)blockIdx.y * a.M;
  float2 *fw = a.fw + (int64_t)blockIdx.y * a.nftot;

  const int r = lane >> 3, q = lane & 7;
  const int xa = x0 - C::HXE, ya = y0 - C::H;
  const int nf0 = a.nf[0], nf1 = a.nf[1], nf2 = a.nf[2];
  int rowoff[S];
#pragma unroll
  for (int s = 0; s < S; s++)
    rowoff[s] = wrap_once(ya + 4 * s + r, nf1) * nf0 + wrap_once(xa + 2 * q, nf0);
  const int64_t pstride = (int64_t)nf0 * nf1;

  unsigned long long acc[S][2][D];
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int k = 0; k < D; k++) acc[s][0][k] = acc[s][1][k] = 0ull;

  // flush ring slot `slot` (holding plane p) to the fine grid and clear it
  auto flush = [&](int p, int slot) {
    const int gz = p < 0 ? p + nf2 : (p >= nf2 ? p - nf2 : p);
    float2 *pl = fw + (int64_t)gz * pstride;
#pragma unroll
    for (int k = 0; k < D; k++) {
      if (slot == k) {
#pragma unroll
        for (int s = 0; s < S; s++) {
          if ((acc[s][0][k] | acc[s][1][k]) & 0x7fffffff7fffffffull) {
            const float2 v0 = unpack2(acc[s][0][k]), v1 = unpack2(acc[s][1][k]);
            red_add4(reinterpret_cast<float4 *>(pl + rowoff[s]), make_float4(v0.x, v0.y, v1.x, v1.y));
          }
          acc[s][0][k] = acc[s][1][k] = 0ull;
        }
      }
    }
  };

  int cur = 0;       // first plane held by the ring
  bool open = false;  // ring holds data
  for (int b0 = 0; b0 < cnt; b0 += C::PB) {
    const int nb = min(C::PB, cnt - b0);
    __syncwarp();
    if (lane < nb) {
      int orig;
      swr_weights<NS, true>(a, tab, first + b0 + lane, xa, ya, cin, KX + lane * C::ROW,
                            KY + lane * C::ROW, KZ + lane * C::ROW, orig);
    }
    __syncwarp();
    for (int t = 0; t < nb; t++) {
      const float4 m = *reinterpret_cast<const float4 *>(KZ + t * C::ROW + 12);
      const int zw = __float_as_int(m.x), mask = __float_as_int(m.y);
      if (zw != cur || !open) {
        if (open) {
          const int nfl = (zw > cur && zw < cur + D) ? zw - cur : D;  // backwards = full flush
          int slot = (cur + C::OFF) % D;
          for (int i = 0; i < nfl; i++) {
            flush(cur + i, slot);
            slot = slot + 1 == D ? 0 : slot + 1;
          }
        }
        cur = zw;
        open = true;
      }
      const float2 kx = *reinterpret_cast<const float2 *>(KX + t * C::ROW + 2 * q);
      const float4 ky4 = *reinterpret_cast<const float4 *>(KY + t * C::ROW + 4 * r);
      float kz[12];
#pragma unroll
      for (int i = 0; i < (D + 3) / 4; i++) {
        const float4 v = *reinterpret_cast<const float4 *>(KZ + t * C::ROW + 4 * i);
        kz[4 * i] = v.x; kz[4 * i + 1] = v.y; kz[4 * i + 2] = v.z; kz[4 * i + 3] = v.w;
      }
      const unsigned long long c2 = pack2(m.z, m.w);
      const unsigned long long cx0 = mul2(c2, pack2(kx.x, kx.x));
      const unsigned long long cx1 = mul2(c2, pack2(kx.y, kx.y));
      const float ky[4] = {ky4.x, ky4.y, ky4.z, ky4.w};
#pragma unroll
      for (int s = 0; s < S; s++) {
        if (mask & (1 << s)) {
          const unsigned long long kyy = pack2(ky[s], ky[s]);
          const unsigned long long w0 = mul2(cx0, kyy), w1 = mul2(cx1, kyy);
#pragma unroll
          for (int k = 0; k < D; k++) {
            const unsigned long long kk = pack2(kz[k], kz[k]);
            fma2(acc[s][0][k], w0, kk);
            fma2(acc[s][1][k], w1, kk);
          }
        }
      }
    }
  }
  if (open) {
    int slot = (cur + C::OFF) % D;
    for (int i = 0; i < D; i++) {
      flush(cur + i, slot);
      slot = slot + 1 == D ? 0 : slot + 1;
    }
  }
}


// ==================================================================================== INTERP
// Per-warp shared memory: the three weight arrays + RES[32][32] float2 partial results
// (point t, lane j at column (j + t) & 31: conflict-free for the per-point store and for the
// per-batch row sums).
template <int NS> struct SwrInterpSmem {
  static constexpr size_t warp_floats = 3 * SwrCfg<NS>::PB * SwrCfg<NS>::ROW + 2 * 32 * 32;
  static constexpr size_t bytes() { return SwrCfg<NS>::WARPS * warp_floats * sizeof(float); }
};

template <int NS>
__global__ void __launch_bounds__(128, 3) k_swr_interp(const SwrArgs a,
                                                       const __grid_constant__ HornerTable<float> tab) {
  using C = SwrCfg<NS>;
  constexpr int D = C::D, S = C::S;
  extern __shared__ __align__(16) float swr_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int first, cnt, x0, y0;
  if (!swr_decode(a, blockIdx.x * C::WARPS + w, first, cnt, x0, y0)) return;
  float *KX = swr_smem + w * SwrInterpSmem<NS>::warp_floats;
  float *KY = KX + C::PB * C::ROW;
  float *KZ = KY + C::PB * C::ROW;
  unsigned long long *RES = reinterpret_cast<unsigned long long *>(KZ + C::PB * C::ROW);
  float2 *cout = a.cout + (int64_t)blockIdx.y * a.M;
  const float2 *fw = a.fw + (int64_t)blockIdx.y * a.nftot;

  const int r = lane >> 3, q = lane & 7;
  const int xa = x0 - C::HXE, ya = y0 - C::H;
  const int nf0 = a.nf[0], nf1 = a.nf[1], nf2 = a.nf[2];
  int rowoff[S];
#pragma unroll
  for (int s = 0; s < S; s++)
    rowoff[s] = wrap_once(ya + 4 * s + r, nf1) * nf0 + wrap_once(xa + 2 * q, nf0);
  const int64_t pstride = (int64_t)nf0 * nf1;

  unsigned long long val[S][2][D];
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int k = 0; k < D; k++) val[s][0][k] = val[s][1][k] = 0ull;

  // load plane p of the fine grid into ring slot `slot`
  auto load = [&](int p, int slot) {
    const int gz = p < 0 ? p + nf2 : (p >= nf2 ? p - nf2 : p);
    const float2 *pl = fw + (int64_t)gz * pstride;
    float4 v[S];
#pragma unroll
    for (int s = 0; s < S; s++) v[s] = __ldg(reinterpret_cast<const float4 *>(pl + rowoff[s]));
#pragma unroll
    for (int k = 0; k < D; k++) {
      if (slot == k) {
#pragma unroll
        for (int s = 0; s < S; s++) {
          val[s][0][k] = pack2(v[s].x, v[s].y);
          val[s][1][k] = pack2(v[s].z, v[s].w);
        }
      }
    }
  };

  int cur = 0;
  bool open = false;
  for (int b0 = 0; b0 < cnt; b0 += C::PB) {
    const int nb = min(C::PB, cnt - b0);
    int orig = 0;
    __syncwarp();
    if (lane < nb)
      swr_weights<NS, false>(a, tab, first + b0 + lane, xa, ya, nullptr, KX + lane * C::ROW,
                             KY + lane * C::ROW, KZ + lane * C::ROW, orig);
    __syncwarp();
    for (int t = 0; t < nb; t++) {
      const float4 m = *reinterpret_cast<const float4 *>(KZ + t * C::ROW + 12);
      const int zw = __float_as_int(m.x), mask = __float_as_int(m.y);
      if (zw != cur || !open) {
        int pfirst = zw, n = D;
        if (open && zw > cur && zw < cur + D) {
          pfirst = cur + D;
          n = zw - cur;
        }
        int slot = (pfirst + C::OFF) % D;
        for (int i = 0; i < n; i++) {
          load(pfirst + i, slot);
          slot = slot + 1 == D ? 0 : slot + 1;
        }
        cur = zw;
        open = true;
      }
      const float2 kx = *reinterpret_cast<const float2 *>(KX + t * C::ROW + 2 * q);
      const float4 ky4 = *reinterpret_cast<const float4 *>(KY + t * C::ROW + 4 * r);
      float kz[12];
#pragma unroll
      for (int i = 0; i < (D + 3) / 4; i++) {
        const float4 v = *reinterpret_cast<const float4 *>(KZ + t * C::ROW + 4 * i);
        kz[4 * i] = v.x; kz[4 * i + 1] = v.y; kz[4 * i + 2] = v.z; kz[4 * i + 3] = v.w;
      }
      const float ky[4] = {ky4.x, ky4.y, ky4.z, ky4.w};
      unsigned long long res = 0ull;
#pragma unroll
      for (int s = 0; s < S; s++) {
        if (mask & (1 << s)) {
          unsigned long long t0 = 0ull, t1 = 0ull;
#pragma unroll
          for (int k = 0; k < D; k++) {
            const unsigned long long kk = pack2(kz[k], kz[k]);
            fma2(t0, val[s][0][k], kk);
            fma2(t1, val[s][1][k], kk);
          }
          const float w0 = ky[s] * kx.x, w1 = ky[s] * kx.y;
          fma2(res, t0, pack2(w0, w0));
          fma2(res, t1, pack2(w1, w1));
        }
      }
      RES[t * 32 + ((lane + t) & 31)] = res;
    }
    __syncwarp();
    if (lane < nb) {
      unsigned long long sum = 0ull;
      const unsigned long long one = pack2(1.f, 1.f);
#pragma unroll 8
      for (int j = 0; j < 32; j++) fma2(sum, RES[lane * 32 + ((j + lane) & 31)], one);
      float2 o = unpack2(sum);
      if (a.scale) {
        const float2 sc = a.scale[orig];
        o = make_float2(o.x * sc.x - o.y * sc.y, o.x * sc.y + o.y * sc.x);
      }
      cout[orig] = o;
    }
  }
}

}  // namespace b2n

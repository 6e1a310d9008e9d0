// tile_kernels.cuh -- shared-memory tile spreading (type 1/3) and interpolation (type 2) for 2-D
// and 3-D on sm_100a.  Replaces spread_{2,3}d_subprob / spread_{2,3}d_output_driven and
// interp_{2,3}d_nupts_driven / interp_{2,3}d_subprob
// (V/src/cuda/3d/spreadinterp3d.cuh:138-383,556-712; V/src/cuda/2d/spreadinterp2d.cuh:116-448).
//
// Why not the reference's design: its SM spreader does 2*ns^d atomicAdd on shared floats per
// point, and on sm_100a a shared-memory float atomicAdd is a CAS spin loop (SASS:
// LDS + FADD + ATOMS.CAST.SPIN), i.e. >= 3 LSU ops per scalar update plus retries.  Here NO
// atomics touch shared memory:
//   * one CTA (8 warps) per subproblem = (bin, chunk of <= maxsub points), tile = bin + halo;
//   * the kernel weights of a batch of 64 points are evaluated once (4 threads per point, one
//     per dimension) and parked in shared memory, x-weights pre-shifted into a 16-byte-aligned
//     window so that every tile access below is one LDS.128 / STS.128;
//   * 3-D: every z-plane of the tile is OWNED by exactly one warp (plane mod 8).  For each point
//     a warp updates the ns x (ns+1) patch of its plane(s): lane = (row, 16-byte column) ->
//     LDS.128, 4 FFMA, STS.128, conflict-free (row stride = 16 words mod 32).  Ownership makes
//     the read-modify-write race-free without atomics or block barriers inside a batch;
//   * 2-D: every warp owns a private replica of the tile and takes every 8th point; replicas
//     are summed at flush;
//   * flush: red.global.add.v4.f32 (REDG.128, two complex cells per lane) straight to the fine
//     grid in L2, skipping all-zero vectors.
// Interpolation mirrors this: tile staged once with LDG.128 -> STS.128 (wrap-aware), one warp
// per point, lanes = (row, column pair), ns plane steps of LDS.128 + 4 FFMA, warp-shuffle reduce.
//
// Algorithmic HBM bytes per point (float, 3-D): coords 12 + idx 4 + c gather 8 (32-B sector)
// + share of the 8*nf grid write.  The binding resource is shared-memory bandwidth:
// ns^3 cells * 8 B * (read+write) per point.
#pragma once
#include "plan.h"

namespace b2n {

template <typename T, int NS> struct TileCfg {
  static constexpr bool F32 = sizeof(T) == 4;
  static constexpr int CPL = F32 ? 2 : 1;                    // complex cells per 16-byte lane access
  static constexpr int XW = F32 ? ((NS + 2) / 2) * 2 : NS;   // aligned x window holding ns cells
  static constexpr int LPR = XW / CPL;                       // lanes per patch row
  static constexpr int RPP = 32 / LPR;                       // patch rows per warp pass
  static constexpr int NPASS = (NS + RPP - 1) / RPP;
  static constexpr int HX = F32 ? ((NS / 2 + 1) / 2) * 2 : NS / 2;  // low-side x halo (even in f32)
  static constexpr int H = NS / 2;                           // low-side halo in y, z
  static constexpr int NW = 8;                               // warps per CTA
  static constexpr int PB = 64;                              // points per weight batch
  __host__ __device__ static constexpr int tx(int bx) {
    int t = bx + XW;
    if (F32 && LPR == 4)
      while (t % 16 != 8) t += 2;  // row stride == 16 words (mod 32): conflict-free quarter-warps
    return t;
  }
  // shared bytes of the per-batch weight area
  __host__ __device__ static constexpr size_t batch_bytes() {
    return (size_t)PB * (XW * sizeof(T) + NS * sizeof(T) + NS * sizeof(cpx<T>) + 16 + sizeof(cpx<T>));
  }
};

template <typename T> struct TileArgs {
  const PtRec<T> *rec;         // sorted points: folded coords + original index
  const int32_t *bin_start, *sp_off, *sp_bin;
  const cpx<T> *cin;           // spread: strengths [ntr][M]
  cpx<T> *cout;                // interp: outputs   [ntr][M]
  const cpx<T> *scale;         // optional per-point factor (type-3 prephase / deconv), by orig index
  cpx<T> *fw;                  // fine grid(s) [ntr][nftot]
  int64_t M, nftot;
  int nf[3], bin[3], nbin[3];
  int64_t nbins;
  int maxsub;
  int TX, TY, TZ;
};

template <typename T> __device__ __forceinline__ cpx<T> cmul(cpx<T> a, cpx<T> b) {
  cpx<T> r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}

// ---- subproblem decode common to all tile kernels --------------------------------------------
template <typename T>
__device__ __forceinline__ bool decode_subproblem(const TileArgs<T> &a, int &first, int &cnt,
                                                   int &xo, int &yo, int &zo) {
  const int sp = blockIdx.x;
  if (sp >= a.sp_off[a.nbins]) return false;
  const int b = a.sp_bin[sp];
  const int s = sp - a.sp_off[b];
  first = a.bin_start[b] + s * a.maxsub;
  cnt = min(a.maxsub, a.bin_start[b + 1] - first);
  xo = (b % a.nbin[0]) * a.bin[0];
  yo = ((b / a.nbin[0]) % a.nbin[1]) * a.bin[1];
  zo = (b / (a.nbin[0] * a.nbin[1])) * a.bin[2];
  return true;
}

// ---- phase 1: weights of one batch, 4 threads per point (sub = dimension) --------------------
// KX[t][XW]   x weights shifted into the aligned window (zeros elsewhere)
// KY[t][NS]   y weights
// KZ[t][NS]   z weights times strength (spread, complex) / z weights (interp, .x only)
// META[t]     int4 {xa, yl, zl, orig index}
template <typename T, int NS, int DIM, bool SPREAD>
__device__ __forceinline__ void batch_weights(const TileArgs<T> &a, const HornerTable<T> &tab,
                                               int p0, int nb, int xo, int yo, int zo,
                                               const cpx<T> *cin, T *KX, T *KY, cpx<T> *KZ,
                                               int4 *META) {
  using C = TileCfg<T, NS>;
  const int t = threadIdx.x >> 2, sub = threadIdx.x & 3;
  if (t >= nb) return;
  const int p = p0 + t;
  T ker[NS];
  if (sub == 0) {
    const T xr = a.rec[p].x;
    const int is = window_start(xr, NS);
    eval_kernel<T, NS>(ker, T(is) - xr, tab);
    const int xl = is - xo + C::HX;
    const int xa = xl & ~(C::CPL - 1);
    const int sh = xl - xa;
    T *dst = KX + t * C::XW;
#pragma unroll
    for (int i = 0; i < C::XW; i++) {
      T v = T(0);
#pragma unroll
      for (int j = 0; j < NS; j++)
        if (i == j + sh) v = ker[j];
      dst[i] = v;
    }
    META[t].x = xa;
  } else if (sub == 1) {
    const T yr = a.rec[p].y;
    const int is = window_start(yr, NS);
    eval_kernel<T, NS>(ker, T(is) - yr, tab);
#pragma unroll
    for (int j = 0; j < NS; j++) KY[t * NS + j] = ker[j];
    META[t].y = is - yo + C::H;
  } else if (sub == 2) {
    const int j0 = (int)a.rec[p].idx;
    META[t].w = j0;
    cpx<T> cv;
    cv.x = T(1);
    cv.y = T(0);
    if (SPREAD) {
      cv = cin[j0];
      if (a.scale) cv = cmul<T>(cv, a.scale[j0]);
    }
    if (DIM == 3) {
      const T zr = a.rec[p].z;
      const int is = window_start(zr, NS);
      eval_kernel<T, NS>(ker, T(is) - zr, tab);
      META[t].z = is - zo + C::H;
#pragma unroll
      for (int j = 0; j < NS; j++) {
        cpx<T> v;
        v.x = ker[j] * cv.x;
        v.y = ker[j] * cv.y;
        KZ[t * NS + j] = v;
      }
    } else {
      META[t].z = 0;
      KZ[t * NS] = cv;
    }
  }
}

template <typename T> struct Vec16;  // 16-byte shared/global vector of the tile
template <> struct Vec16<float> { using type = float4; };
template <> struct Vec16<double> { using type = double2; };

// acc(16 B of tile) += (w0, w1) x kz   [float: two cells; double: one cell, w1 unused]
__device__ __forceinline__ void patch_fma(float4 &v, float w0, float w1, float2 k) {
  v.x = fmaf(w0, k.x, v.x);
  v.y = fmaf(w0, k.y, v.y);
  v.z = fmaf(w1, k.x, v.z);
  v.w = fmaf(w1, k.y, v.w);
}
__device__ __forceinline__ void patch_fma(double2 &v, double w0, double, double2 k) {
  v.x = fma(w0, k.x, v.x);
  v.y = fma(w0, k.y, v.y);
}

__device__ __forceinline__ void acc_fma(float4 &A, float k, const float4 &B) {
  A.x = fmaf(k, B.x, A.x);
  A.y = fmaf(k, B.y, A.y);
  A.z = fmaf(k, B.z, A.z);
  A.w = fmaf(k, B.w, A.w);
}
__device__ __forceinline__ void acc_fma(double2 &A, double k, const double2 &B) {
  A.x = fma(k, B.x, A.x);
  A.y = fma(k, B.y, A.y);
}
__device__ __forceinline__ void acc_dot(const float4 &A, float w0, float w1, float &re, float &im) {
  re += w0 * A.x + w1 * A.z;
  im += w0 * A.y + w1 * A.w;
}
__device__ __forceinline__ void acc_dot(const double2 &A, double w0, double, double &re, double &im) {
  re += w0 * A.x;
  im += w0 * A.y;
}

__device__ __forceinline__ void load_w01(const float *KXt, int q, float ky, float &w0, float &w1) {
  const float2 k2 = *reinterpret_cast<const float2 *>(KXt + 2 * q);
  w0 = k2.x * ky;
  w1 = k2.y * ky;
}
__device__ __forceinline__ void load_w01(const double *KXt, int q, double ky, double &w0, double &w1) {
  w0 = KXt[q] * ky;
  w1 = 0.0;
}

// ---- flush helpers ------------------------------------------------------------------------------
__device__ __forceinline__ bool nonzero(const float4 &v) {
  return (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f) | (v.w != 0.f);
}
__device__ __forceinline__ bool nonzero(const double2 &v) { return (v.x != 0.0) | (v.y != 0.0); }
__device__ __forceinline__ void red16(float2 *cell, const float4 &v) {
  red_add4(reinterpret_cast<float4 *>(cell), v);
}
__device__ __forceinline__ void red16(double2 *cell, const double2 &v) { red_add(cell, v); }
__device__ __forceinline__ void vadd(float4 &a, const float4 &b) {
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}
__device__ __forceinline__ void vadd(double2 &a, const double2 &b) { a.x += b.x; a.y += b.y; }

// ==================================================================================== SPREAD 3-D
template <typename T, int NS>
__global__ void __launch_bounds__(256) k_spread3d(const TileArgs<T> a,
                                                   const __grid_constant__ HornerTable<T> tab) {
  using C = TileCfg<T, NS>;
  using V = typename Vec16<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int first, cnt, xo, yo, zo;
  if (!decode_subproblem(a, first, cnt, xo, yo, zo)) return;
  const int TX = a.TX, TY = a.TY, TZ = a.TZ;
  const int ncell = TX * TY * TZ;
  cpx<T> *tile = reinterpret_cast<cpx<T> *>(smem_raw);
  T *KX = reinterpret_cast<T *>(tile + ncell);
  T *KY = KX + C::PB * C::XW;
  cpx<T> *KZ = reinterpret_cast<cpx<T> *>(KY + C::PB * NS);
  int4 *META = reinterpret_cast<int4 *>(KZ + C::PB * NS);
  const cpx<T> *cin = a.cin + (int64_t)blockIdx.y * a.M;
  cpx<T> *fw = a.fw + (int64_t)blockIdx.y * a.nftot;

  {  // zero the tile
    V z = {};
    V *tv = reinterpret_cast<V *>(tile);
    for (int i = threadIdx.x; i < ncell / C::CPL; i += 256) tv[i] = z;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = lane / C::LPR, q = lane % C::LPR;
  const bool lane_on = r < C::RPP;

  for (int b0 = 0; b0 < cnt; b0 += C::PB) {
    const int nb = min(C::PB, cnt - b0);
    __syncthreads();  // tile zeroed / previous batch fully consumed
    batch_weights<T, NS, 3, true>(a, tab, first + b0, nb, xo, yo, zo, cin, KX, KY, KZ, META);
    __syncthreads();
    for (int t = 0; t < nb; t++) {
      const int4 m = META[t];
      const T *KXt = KX + t * C::XW;
      // planes of [zl, zl+NS) owned by this warp: P == w (mod 8)
      for (int P = m.z + ((w - m.z) & 7); P < m.z + NS; P += 8) {
        const cpx<T> kz = KZ[t * NS + (P - m.z)];
#pragma unroll
        for (int ps = 0; ps < C::NPASS; ps++) {
          const int row = ps * C::RPP + r;
          if (lane_on && row < NS) {
            T w0, w1;
            load_w01(KXt, q, KY[t * NS + row], w0, w1);
            V *ptr = reinterpret_cast<V *>(tile + ((P * TY + m.y + row) * TX + m.x + q * C::CPL));
            V v = *ptr;
            patch_fma(v, w0, w1, kz);
            *ptr = v;
          }
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // flush: tile -> fine grid, periodic wrap, 16 B per lane
  const int nvx = TX / C::CPL;
  for (int i = threadIdx.x; i < nvx * TY * TZ; i += 256) {
    const int vx = i % nvx, yy = (i / nvx) % TY, zz = i / (nvx * TY);
    int gx = xo - C::HX + vx * C::CPL, gy = yo - C::H + yy, gz = zo - C::H + zz;
    if (gx > a.nf[0] + C::H || gy > a.nf[1] + C::H || gz > a.nf[2] + C::H) continue;
    const V v = *reinterpret_cast<const V *>(tile + ((zz * TY + yy) * TX + vx * C::CPL));
    if (!nonzero(v)) continue;
    gx = wrap_once(gx, a.nf[0]);
    gy = wrap_once(gy, a.nf[1]);
    gz = wrap_once(gz, a.nf[2]);
    red16(fw + ((int64_t)gz * a.nf[1] + gy) * a.nf[0] + gx, v);
  }
}

// ==================================================================================== SPREAD 2-D
template <typename T, int NS>
__global__ void __launch_bounds__(256) k_spread2d(const TileArgs<T> a,
                                                   const __grid_constant__ HornerTable<T> tab) {
  using C = TileCfg<T, NS>;
  using V = typename Vec16<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int first, cnt, xo, yo, zo;
  if (!decode_subproblem(a, first, cnt, xo, yo, zo)) return;
  const int TX = a.TX, TY = a.TY;
  const int ncell = TX * TY;  // per replica
  cpx<T> *tile = reinterpret_cast<cpx<T> *>(smem_raw);
  T *KX = reinterpret_cast<T *>(tile + ncell * C::NW);
  T *KY = KX + C::PB * C::XW;
  cpx<T> *KZ = reinterpret_cast<cpx<T> *>(KY + C::PB * NS);
  int4 *META = reinterpret_cast<int4 *>(KZ + C::PB * NS);
  const cpx<T> *cin = a.cin + (int64_t)blockIdx.y * a.M;
  cpx<T> *fw = a.fw + (int64_t)blockIdx.y * a.nftot;
  {
    V z = {};
    V *tv = reinterpret_cast<V *>(tile);
    for (int i = threadIdx.x; i < ncell * C::NW / C::CPL; i += 256) tv[i] = z;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = lane / C::LPR, q = lane % C::LPR;
  const bool lane_on = r < C::RPP;
  cpx<T> *mine = tile + w * ncell;  // this warp's private replica

  for (int b0 = 0; b0 < cnt; b0 += C::PB) {
    const int nb = min(C::PB, cnt - b0);
    __syncthreads();
    batch_weights<T, NS, 2, true>(a, tab, first + b0, nb, xo, yo, zo, cin, KX, KY, KZ, META);
    __syncthreads();
    for (int t = w; t < nb; t += C::NW) {
      const int4 m = META[t];
      const T *KXt = KX + t * C::XW;
      const cpx<T> cv = KZ[t * NS];
#pragma unroll
      for (int ps = 0; ps < C::NPASS; ps++) {
        const int row = ps * C::RPP + r;
        if (lane_on && row < NS) {
          T w0, w1;
          load_w01(KXt, q, KY[t * NS + row], w0, w1);
          V *ptr = reinterpret_cast<V *>(mine + ((m.y + row) * TX + m.x + q * C::CPL));
          V v = *ptr;
          patch_fma(v, w0, w1, cv);
          *ptr = v;
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  const int nvx = TX / C::CPL;
  for (int i = threadIdx.x; i < nvx * TY; i += 256) {
    const int vx = i % nvx, yy = i / nvx;
    int gx = xo - C::HX + vx * C::CPL, gy = yo - C::H + yy;
    if (gx > a.nf[0] + C::H || gy > a.nf[1] + C::H) continue;
    V v = *reinterpret_cast<const V *>(tile + (yy * TX + vx * C::CPL));
#pragma unroll
    for (int k = 1; k < C::NW; k++)
      vadd(v, *reinterpret_cast<const V *>(tile + k * ncell + (yy * TX + vx * C::CPL)));
    if (!nonzero(v)) continue;
    gx = wrap_once(gx, a.nf[0]);
    gy = wrap_once(gy, a.nf[1]);
    red16(fw + (int64_t)gy * a.nf[0] + gx, v);
  }
}

// ==================================================================================== INTERP
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, int NS, int DIM>
__global__ void __launch_bounds__(256) k_interp(const TileArgs<T> a,
                                                 const __grid_constant__ HornerTable<T> tab) {
  using C = TileCfg<T, NS>;
  using V = typename Vec16<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int first, cnt, xo, yo, zo;
  if (!decode_subproblem(a, first, cnt, xo, yo, zo)) return;
  const int TX = a.TX, TY = a.TY, TZ = DIM == 3 ? a.TZ : 1;
  const int ncell = TX * TY * TZ;
  cpx<T> *tile = reinterpret_cast<cpx<T> *>(smem_raw);
  T *KX = reinterpret_cast<T *>(tile + ncell);
  T *KY = KX + C::PB * C::XW;
  cpx<T> *KZ = reinterpret_cast<cpx<T> *>(KY + C::PB * NS);
  int4 *META = reinterpret_cast<int4 *>(KZ + C::PB * NS);
  cpx<T> *OUT = reinterpret_cast<cpx<T> *>(META + C::PB);
  cpx<T> *cout = a.cout + (int64_t)blockIdx.y * a.M;
  const cpx<T> *fw = a.fw + (int64_t)blockIdx.y * a.nftot;

  // stage the tile (bin + halo) from the fine grid: LDG.128 -> STS.128, periodic wrap
  const int nvx = TX / C::CPL;
  for (int i = threadIdx.x; i < nvx * TY * TZ; i += 256) {
    const int vx = i % nvx, yy = (i / nvx) % TY, zz = i / (nvx * TY);
    int gx = xo - C::HX + vx * C::CPL, gy = yo - C::H + yy, gz = DIM == 3 ? zo - C::H + zz : 0;
    V v = {};
    if (gx <= a.nf[0] + C::H && gy <= a.nf[1] + C::H && (DIM < 3 || gz <= a.nf[2] + C::H)) {
      gx = wrap_once(gx, a.nf[0]);
      gy = wrap_once(gy, a.nf[1]);
      if (DIM == 3) gz = wrap_once(gz, a.nf[2]);
      v = __ldg(reinterpret_cast<const V *>(fw + ((int64_t)gz * a.nf[1] + gy) * a.nf[0] + gx));
    }
    *reinterpret_cast<V *>(tile + ((zz * TY + yy) * TX + vx * C::CPL)) = v;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = lane / C::LPR, q = lane % C::LPR;
  const bool lane_on = r < C::RPP;

  for (int b0 = 0; b0 < cnt; b0 += C::PB) {
    const int nb = min(C::PB, cnt - b0);
    __syncthreads();
    batch_weights<T, NS, DIM, false>(a, tab, first + b0, nb, xo, yo, zo, nullptr, KX, KY, KZ, META);
    __syncthreads();
    for (int t = w; t < nb; t += C::NW) {
      const int4 m = META[t];
      const T *KXt = KX + t * C::XW;
      T re = T(0), im = T(0);
#pragma unroll
      for (int ps = 0; ps < C::NPASS; ps++) {
        const int row = ps * C::RPP + r;
        if (lane_on && row < NS) {
          T w0, w1;
          load_w01(KXt, q, KY[t * NS + row], w0, w1);
          const cpx<T> *base = tile + ((m.z * TY + m.y + row) * TX + m.x + q * C::CPL);
          V acc = {};
          if (DIM == 3) {
#pragma unroll
            for (int k = 0; k < NS; k++)
              acc_fma(acc, KZ[t * NS + k].x, *reinterpret_cast<const V *>(base + k * TY * TX));
          } else {
            acc = *reinterpret_cast<const V *>(base);
          }
          acc_dot(acc, w0, w1, re, im);
        }
      }
      re = warp_sum(re);
      im = warp_sum(im);
      if (lane == 0) {
        cpx<T> o;
        o.x = re;
        o.y = im;
        OUT[t] = o;
      }
    }
    __syncthreads();
    if (threadIdx.x < nb) {
      const int j0 = META[threadIdx.x].w;
      cpx<T> o = OUT[threadIdx.x];
      if (a.scale) o = cmul<T>(o, a.scale[j0]);
      cout[j0] = o;
    }
  }
}

}  // namespace b2n

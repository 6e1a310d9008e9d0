// swr2_kernels.cuh -- second generation of the sliding-window REGISTER kernels (3-D, float, ns <= 8).
// Same decomposition as swr_kernels.cuh (one warp per subproblem, 8 x 12 cell window, a ring of
// D = ns z planes in registers, points in anchor-z order), different control structure:
//
//   * ABSOLUTE RING SLOTS.  Plane p (counted from a per-subproblem base zb) lives in ring slot
//     p mod D for the whole subproblem.  The lane that evaluates a point's kernel vectors stores
//     the z weights already rotated into slot order (slot (p0 + j) mod D holds kz[j]), so the
//     FMA block is ONE straight-line piece of code -- acc[s][k] += wv[s] * kzslot[k], k literal --
//     for every point, whatever plane its window starts at.  The first generation reached literal
//     accumulator indices by instantiating the whole loop body D times (a chain of phases entered
//     through a switch): 7x the code (instruction-cache misses, `no_instruction` stalls) and ~9
//     control instructions per point to move between phases.
//   * SENTINEL-TERMINATED RUNS.  Every row carries META = ((plane << 1 | half) << 2 | y class) in
//     the unused fourth column of its y block, so it arrives with the y weights; each batch is
//     ordered by (plane, class) and ends at a sentinel row (META = -1).  The inner loops are
//     `while (meta == key) point<class>()`: one compare and one branch per point, no run tables.
//     Plane changes (retire the planes that left the window) are handled between the runs.
//     (Tried and measured slower on B200, DESIGN.md 4.2: run boundaries and class masks gathered
//     with three ballots per batch so that the loops have warp-uniform trip counts -- ptxas then
//     places every rolling LDS directly in front of its use and at 4-5 warps per scheduler the
//     exposed latency is not covered: 7.02 -> 7.65 ms at C3.)
//   The absolute-slot INTERPOLATOR of this generation was slower than the phase chain of
//   swr_kernels.cuh and has been removed; what it taught -- staging the next planes
//   through shared memory with cp.async -- lives on in k_swr_interp.
#pragma once
#include "swr_kernels.cuh"

namespace b2n {


// Row of one point in shared memory (floats).  Differences from SwrCfg: META sits in the unused
// fourth column of the y block (one copy per lane row, so it arrives with the y weights, first
// thing in the FMA block), and the z block holds only the D weights, in RING-SLOT order.
template <int NS> struct Swr2Cfg : SwrCfg<NS> {
  using B = SwrCfg<NS>;
  static constexpr int KXO = 0;                 // WX pairs: spread (c.re*kx, c.im*kx); interp (kx, kx)
  static constexpr int KYO = 2 * B::WX;         // [r = row & 3]{ky(r), ky(4 + r), ky(8 + r), META}
  static constexpr int KZO = KYO + 16;          // D z weights by ring slot
  static constexpr int KZW = (B::D + 3) & ~3;
  static constexpr int ROW0 = KZO + KZW;
  static constexpr int ROW = (ROW0 / 4) % 2 == 1 ? ROW0 : ROW0 + 4;  // stride/4 odd (see SwrCfg)
  static constexpr int NV = KZW / 4;
  // CTAs per SM of the spreader.  With the look-ahead in shared memory the ns = 6, 7 kernels fit 96
  // registers without spills: 18 per SM (C3: 6.81 ms; 16 -> 7.11, 20 -> 6.97)
#ifdef SWR2_MINB
  static constexpr int MINB2 = SWR2_MINB;
#else
  static constexpr int MINB2 = (NS == 6 || NS == 7) ? 18 : B::MINB;
#endif
  static constexpr int NROWS = B::PB + 1;       // + one row: the rolling loads of the last point read a row ahead
  static constexpr int ZBIAS = 16 * B::D;       // planes are counted from zb = z0 - H - ZBIAS: always > 0
  static constexpr int SPREAD_FLOATS = NROWS * ROW;
  // spread: rows + the look-ahead rings (3 x 32 records of 16 B, 2 x 32 strengths of 8 B)
  static constexpr int SPREAD_SMEM = (SPREAD_FLOATS + 3 * 128 + 2 * 64) * (int)sizeof(float);
  static_assert(B::S <= 3, "META uses the fourth y column");
};

__device__ __forceinline__ bool swr2_decode(const SwrArgs &a, int sp, int &first, int &cnt, int &x0, int &y0, int &z0) {
  if (sp >= a.sp_off[a.nbins]) return false;
  const int b = a.sp_bin[sp];
  const int s = sp - a.sp_off[b];
  first = a.bin_start[b] + s * a.maxsub;
  cnt = min(a.maxsub, a.bin_start[b + 1] - first);
  const int bxy = a.nbin[0] * a.nbin[1];
  const int bz = b / bxy, r = b - bz * bxy;
  const int by = r / a.nbin[0];
  x0 = (r - by * a.nbin[0]) * a.bin[0];
  y0 = by * a.bin[1];
  z0 = bz * a.bin[2];
  return cnt > 0;
}

// shared-memory accesses by 32-bit shared address: no generic-to-shared conversion in the loops
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds64(unsigned a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}

__device__ __forceinline__ void sts64(unsigned a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts32(unsigned a, int v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ int lds32(unsigned a) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}

__device__ __forceinline__ void cp_async8_u32(unsigned d, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_u32(unsigned d, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// the three kernel vectors of one point in one Horner sweep, two intervals per FFMA2; NC > 0:
// coefficient count known at compile time (fully unrolled, no loop bookkeeping)
template <int NS, int NC>
__device__ __forceinline__ void swr2_horner(const HornerTable<float> &tab, float zx, float zy, float zz,
                                            float (&kx)[2 * SwrCfg<NS>::NP], float (&ky)[2 * SwrCfg<NS>::NP],
                                            float (&kz)[2 * SwrCfg<NS>::NP]) {
  constexpr int NP = SwrCfg<NS>::NP;
  const float2 zx2 = make_float2(zx, zx), zy2 = make_float2(zy, zy), zz2 = make_float2(zz, zz);
  float2 ax[NP], ay[NP], az[NP];
#pragma unroll
  for (int j = 0; j < NP; j++) ax[j] = ay[j] = az[j] = make_float2(tab.c[0][2 * j], tab.c[0][2 * j + 1]);
  auto step = [&](int k) {
#pragma unroll
    for (int j = 0; j < NP; j++) {
      const float2 cj = make_float2(tab.c[k][2 * j], tab.c[k][2 * j + 1]);
      ax[j] = fma2(ax[j], zx2, cj);
      ay[j] = fma2(ay[j], zy2, cj);
      az[j] = fma2(az[j], zz2, cj);
    }
  };
  if constexpr (NC > 0) {
#pragma unroll
    for (int k = 1; k < NC; k++) step(k);
  } else {
    for (int k = 1; k < tab.ncoef; k++) step(k);
  }
#pragma unroll
  for (int j = 0; j < NP; j++) {
    kx[2 * j] = ax[j].x; kx[2 * j + 1] = ax[j].y;
    ky[2 * j] = ay[j].x; ky[2 * j + 1] = ay[j].y;
    kz[2 * j] = az[j].x; kz[2 * j + 1] = az[j].y;
  }
}

// lane t parks the weights of its point in `row` (layout: Swr2Cfg).  The z weights go to their
// ring slots: plane prel + j (prel = first plane of the window, counted from the subproblem's
// base) -> row[KZO + (prel + j) mod D].  meta = ((prel << 1 | half) << 2) | cls.
template <int NS>
__device__ __forceinline__ void swr2_weights(const HornerTable<float> &tab, const float4 pr4, float2 cv, int xa,
                                             int ya, int zb, float *row, int half, int cls) {
  using C = Swr2Cfg<NS>;
  constexpr int NP = C::NP, D = C::D;
  const float px = pr4.x, py = pr4.y, pz = pr4.z;
  const int isx = window_start(px, NS), isy = window_start(py, NS), isz = window_start(pz, NS);
  const int prel = isz - zb;  // > 0 (ZBIAS)
  {
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 m4 = make_float4(0.f, 0.f, 0.f, __int_as_float((((prel << 1) | half) << 2) | cls));
#pragma unroll
    for (int i = 0; i < C::KYO / 4; i++) reinterpret_cast<float4 *>(row)[i] = z4;
#pragma unroll
    for (int i = 0; i < 4; i++) reinterpret_cast<float4 *>(row + C::KYO)[i] = m4;
  }
  float kx[2 * NP], ky[2 * NP], kz[2 * NP];
  if (!tab.direct) {
    const float zx = fmaf(2.f, float(isx) - px, float(NS - 1));
    const float zy = fmaf(2.f, float(isy) - py, float(NS - 1));
    const float zz = fmaf(2.f, float(isz) - pz, float(NS - 1));
    // coefficient count horner_fit (hostmath.cpp) settles on for float tables at sigma = 2; any
    // other table (sigma = 1.25, ...) takes the run-time loop
    constexpr int NC = NS == 2 ? 6 : NS == 3 ? 5 : NS == 4 ? 8 : NS == 5 ? 7 : 9;
    if (tab.ncoef == NC) swr2_horner<NS, NC>(tab, zx, zy, zz, kx, ky, kz);
    else swr2_horner<NS, 0>(tab, zx, zy, zz, kx, ky, kz);
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
    const unsigned sl = (unsigned)prel % (unsigned)D;
    float *kzr = row + C::KZO;
#pragma unroll
    for (int j = 0; j < NS; j++) kzr[min(sl + j, sl + j - D)] = kz[j];  // (sl + j) mod D: the wrapped term underflows
  }
}

// One point's row as the inner loops hold it: this lane's x weight(s), its y weights + META, the
// D (kz, kz) scalars by ring slot.  The loops keep ONE such set live and reload each piece from
// the NEXT row right after its last use ("rolling" loads).
template <int NS> struct Swr2Row {
  using C = Swr2Cfg<NS>;
  static constexpr int NV = C::NV;
  float4 kv[NV];
  float2 cx[C::CX];
  float4 ky;  // .w = META
  __device__ __forceinline__ void load_xy(unsigned ax, unsigned ay) {
    if constexpr (C::CX == 1) {
      cx[0] = lds64(ax);
    } else {
      const float4 v = lds128(ax);
      cx[0] = make_float2(v.x, v.y);
      cx[1] = make_float2(v.z, v.w);
    }
    ky = lds128(ay);
  }
  __device__ __forceinline__ void load_kv(unsigned az, int i) { kv[i] = lds128(az + 16 * i); }
  __device__ __forceinline__ void load_all(unsigned ax, unsigned ay, unsigned az) {
    load_xy(ax, ay);
#pragma unroll
    for (int i = 0; i < NV; i++) load_kv(az, i);
  }
  __device__ __forceinline__ float2 kz(int j) const {  // (kz, kz): a scalar-broadcast FFMA2 operand
    const float4 v = kv[j / 4];
    const float k = (j & 3) == 0 ? v.x : ((j & 3) == 1 ? v.y : ((j & 3) == 2 ? v.z : v.w));
    return make_float2(k, k);
  }
  __device__ __forceinline__ int meta() const { return __float_as_int(ky.w); }
  __device__ __forceinline__ float2 kyv(int s) const {
    const float k = s == 0 ? ky.x : (s == 1 ? ky.y : ky.z);
    return make_float2(k, k);
  }
};

// ==================================================================================== SPREAD
template <int NS, bool SCALED>
__global__ void __launch_bounds__(32, Swr2Cfg<NS>::MINB2)
    k_swr2_spread(const SwrArgs a, const __grid_constant__ HornerTable<float> tab) {
  using C = Swr2Cfg<NS>;
  constexpr int D = C::D, S = C::S, CX = C::CX, NV = C::NV;
  extern __shared__ __align__(16) float swr_smem[];
  const int lane = threadIdx.x;
  int first, cnt, x0, y0, z0;
  if (!swr2_decode(a, blockIdx.x, first, cnt, x0, y0, z0)) return;
  float *rows = swr_smem;
  const float2 *cin = a.cin + (int64_t)blockIdx.y * a.M;

  const int r = lane >> 3, q = lane & 7;
  const int xa = x0 - C::H, ya = y0 - C::H, zb = z0 - C::H - C::ZBIAS;
  const int nf0 = a.nf[0], nf1 = a.nf[1], nf2 = a.nf[2];
  const int64_t pstride = (int64_t)nf0 * nf1;

  float2 acc[S][CX][D];
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int c = 0; c < CX; c++)
#pragma unroll
      for (int k = 0; k < D; k++) acc[s][c][k] = make_float2(0.f, 0.f);

  // Retire the first plane of the ring (relative plane `cur`, ring slot `slot`): RED this lane's
  // cells, clear the slot, move on by one plane.  pz[s] = this lane's cell of row slot s in that
  // plane, carried along (+ one plane per retire, periodic wrap) instead of being rebuilt from a
  // 64-bit multiply per plane.
  float2 *pz[S];
  int gz = 0;  // fine-grid plane of relative plane cur
  auto seek = [&](int prel) {  // position pz on relative plane prel
    gz = prel + zb;
    gz = gz < 0 ? gz + nf2 : gz;
    if (gz >= nf2) gz = gz - nf2 < nf2 ? gz - nf2 : gz % nf2;
    float2 *base = a.fw + (int64_t)blockIdx.y * a.nftot + (int64_t)gz * pstride + wrap_once(xa + CX * q, nf0);
#pragma unroll
    for (int s = 0; s < S; s++) pz[s] = base + wrap_once(ya + 4 * s + r, nf1) * nf0;  // once per subproblem (or gap)
  };
  auto retire = [&](int slot) {
    auto one = [&](auto kc) {
      constexpr int K = decltype(kc)::value;
      if constexpr (K < D) {
#pragma unroll
        for (int s = 0; s < S; s++) {
          if constexpr (CX == 1) {
            red_add(pz[s], acc[s][0][K]);
          } else {
            red_add4(reinterpret_cast<float4 *>(pz[s]),
                     make_float4(acc[s][0][K].x, acc[s][0][K].y, acc[s][1][K].x, acc[s][1][K].y));
          }
#pragma unroll
          for (int c = 0; c < CX; c++) acc[s][c][K] = make_float2(0.f, 0.f);
        }
      }
    };
    switch (slot) {
      case 0: one(std::integral_constant<int, 0>{}); break;
      case 1: one(std::integral_constant<int, 1>{}); break;
      case 2: one(std::integral_constant<int, 2>{}); break;
      case 3: one(std::integral_constant<int, 3>{}); break;
      case 4: one(std::integral_constant<int, 4>{}); break;
      case 5: one(std::integral_constant<int, 5>{}); break;
      case 6: one(std::integral_constant<int, 6>{}); break;
      default: one(std::integral_constant<int, 7>{}); break;
    }
    // (a never-taken branch for the periodic wrap instead of these selects: no measurable gain, r02x)
    const int64_t step = gz + 1 == nf2 ? -(int64_t)(nf2 - 1) * pstride : pstride;
    gz = gz + 1 == nf2 ? 0 : gz + 1;
#pragma unroll
    for (int s = 0; s < S; s++) pz[s] += step;
  };

  Swr2Row<NS> pr;
  // this lane's pieces of the row being consumed (32-bit shared addresses; + ROW per point)
  const unsigned ax0 = smem_u32(rows + C::KXO + 2 * CX * q), ay0 = smem_u32(rows + C::KYO + 4 * r),
                 az0 = smem_u32(rows + C::KZO);
  unsigned ax = ax0, ay = ay0, az = az0;
  constexpr unsigned RB = C::ROW * sizeof(float);
  // all ns^2 x ns cell updates of the point held in `pr`, then roll `pr` on to the next row.
  // CLS: 0 = the point's y window ends below row slot S-1, 2 = it starts above row slot 0,
  // 1 = anything: the untouched slot's FMUL2 + D FFMA2 are not issued at all
  // The rolling loads address the row K rows ahead of (ax, ay, az) with an immediate offset; the
  // pointers move once per two points (advance), far from the loads that use them -- bumped right
  // behind an LDS, the add waits for the load to release its address register (short-scoreboard
  // stalls on every pointer increment, profiles/r02k source view).
  // The adds are volatile: left to itself the compiler computed the three pointers of every exit
  // path ahead of its branch (3 IADD3 per point, on the ALU pipe that runs at one instruction per two
  // cycles) and copied them back on the way out; now a run pays them once, where it ends
  // (type-1 step 10.92 -> 10.72 ms at C3 together with four points per trip, profiles/r02ab).
  auto advance = [&](int k) {
    const unsigned d = k * RB;
    asm volatile("add.u32 %0, %0, %1;" : "+r"(ax) : "r"(d));
    asm volatile("add.u32 %0, %0, %1;" : "+r"(ay) : "r"(d));
    asm volatile("add.u32 %0, %0, %1;" : "+r"(az) : "r"(d));
  };
  auto point = [&](auto clc, auto kc) {
    constexpr int CLS = decltype(clc)::value;
    constexpr unsigned K = decltype(kc)::value;
    constexpr int S0 = CLS == 2 ? 1 : 0, S1 = CLS == 0 ? S - 1 : S;
    float2 wv[S][CX];
#pragma unroll
    for (int s = S0; s < S1; s++)
#pragma unroll
      for (int c = 0; c < CX; c++) wv[s][c] = mul2(pr.cx[c], pr.kyv(s));
    pr.load_xy(ax + K * RB, ay + K * RB);
#pragma unroll
    for (int i = 0; i < NV; i++) {
#pragma unroll
      for (int j = 4 * i; j < 4 * i + 4; j++) {
        if (j < D) {
          const float2 kzj = pr.kz(j);
#pragma unroll
          for (int s = S0; s < S1; s++)
#pragma unroll
            for (int c = 0; c < CX; c++) acc[s][c][j] = fma2(wv[s][c], kzj, acc[s][c][j]);
        }
      }
      pr.load_kv(az + K * RB, i);
    }
  };
  const unsigned sm0 = smem_u32(swr_smem);
  auto sentinel = [&](int row) {  // META = -1 in all four y-block copies of `row`
    if (lane < 4) sts32(sm0 + (row * C::ROW + C::KYO + 4 * lane + 3) * 4, -1);
  };

  int cur = SWR_EMPTY;  // first (relative) plane held by the ring
  int slot = 0;         // ring slot of plane cur = cur mod D
  // Two-deep pipeline over batches of 32 points, entirely in shared memory (cp.async): the record
  // of batch b+2 and the strength of batch b+1 (whose address comes from the record of b+1) are in
  // flight while the warp spreads batch b.  (Held in registers, as in the first generation, the
  // look-ahead costs 14 registers -- and when the register budget is tightened ptxas spills exactly
  // those, right behind their loads: every batch then waits out a DRAM round trip, profiles/r02h.)
  const PtRec<float> *recp = a.rec + first + lane;
  const unsigned rq = sm0 + (C::SPREAD_FLOATS + 4 * lane) * 4;        // record ring: 3 slots x 32 x 16 B
  const unsigned cq = sm0 + (C::SPREAD_FLOATS + 3 * 128 + 2 * lane) * 4;  // strength ring: 2 slots x 32 x 8 B
  auto issue_rec = [&](int b) {  // batch index b -> slot b % 3
    if (b * C::PB + lane < cnt) cp_async16_u32(rq + (b % 3) * 512, recp + b * C::PB);
    cp_async_commit();
  };
  auto issue_str = [&](int b) {  // needs the record of batch b in shared memory
    if (b * C::PB + lane < cnt) cp_async8_u32(cq + (b & 1) * 256, cin + lds32(rq + (b % 3) * 512 + 12));
    cp_async_commit();
  };
  issue_rec(0);
  issue_rec(1);
  cp_async_wait<1>();
  issue_str(0);
  sentinel(C::PB);
  int bi = 0;
  for (int b0 = 0; b0 < cnt; b0 += C::PB, bi++) {
    const int nb = min(C::PB, cnt - b0);
    issue_rec(bi + 2);
    cp_async_wait<1>();   // record bi+1 and strength bi have landed
    issue_str(bi + 1);
    const float4 recA = lane < nb ? lds128(rq + (bi % 3) * 512) : make_float4(0.f, 0.f, 0.f, 0.f);
    float2 cA = lane < nb ? lds64(cq + (bi & 1) * 256) : make_float2(0.f, 0.f);
    __syncwarp();
    int pos, cls;
#if SWR_YCLASS
    swr_batch_order<NS>(recA, nb, ya, lane, pos, cls);
#else
    pos = lane; cls = 1;
#endif
    if (lane < nb) {
      if constexpr (SCALED) {  // type 3: strength times the prephase of the point
        const float2 sA = __ldg(a.scale + __float_as_int(recA.w));
        cA = make_float2(cA.x * sA.x - cA.y * sA.y, cA.x * sA.y + cA.y * sA.x);
      }
      swr2_weights<NS>(tab, recA, cA, xa, ya, zb, rows + pos * C::ROW, 0, cls);
    }
    if (nb < C::PB) sentinel(nb);
    __syncwarp();
    ax = ax0; ay = ay0; az = az0;
    pr.load_all(ax, ay, az);
    for (;;) {
      int mt = pr.meta();
      if (mt < 0) break;  // sentinel row
      const int z = mt >> 3;
      if (z != cur) {
        if (cur != SWR_EMPTY) {  // retire the planes below the new window (all D after a gap / disorder)
          const unsigned d = (unsigned)(z - cur);
          const int nr = d >= (unsigned)D ? D : (int)d;
#pragma unroll 1
          for (int i = 0; i < nr; i++) {
            retire(slot);
            slot = slot + 1 == D ? 0 : slot + 1;
            cur++;
          }
        }
        if (cur != z) {  // first point of the subproblem, or a jump over empty planes
          cur = z;
          slot = (int)((unsigned)z % (unsigned)D);
          seek(z);
        }
      }
      const int k0 = mt & ~3;
#ifndef SWR2_UNROLL4
#define SWR2_UNROLL4 1
#endif
#if SWR2_UNROLL4
#define SWR2_STEP(CL, K)                                                                        \
        point(std::integral_constant<int, CL>{}, std::integral_constant<unsigned, K>{});         \
        mt = pr.meta();                                                                          \
        if (mt != k0 + CL) {                                                                     \
          advance(K);                                                                            \
          break;                                                                                 \
        }
#define SWR2_RUN(CL)                                                                            \
      while (mt == k0 + CL) { /* four points per trip; (ax, ay, az) = the row held in pr */      \
        SWR2_STEP(CL, 1)                                                                         \
        SWR2_STEP(CL, 2)                                                                         \
        SWR2_STEP(CL, 3)                                                                         \
        point(std::integral_constant<int, CL>{}, std::integral_constant<unsigned, 4>{});         \
        advance(4);                                                                              \
        mt = pr.meta();                                                                          \
      }
#else
#define SWR2_RUN(CL)                                                                            \
      while (mt == k0 + CL) { /* two points per trip; (ax, ay, az) = the row held in pr */       \
        point(std::integral_constant<int, CL>{}, std::integral_constant<unsigned, 1>{});         \
        mt = pr.meta();                                                                          \
        if (mt != k0 + CL) {                                                                     \
          advance(1);                                                                            \
          break;                                                                                 \
        }                                                                                        \
        point(std::integral_constant<int, CL>{}, std::integral_constant<unsigned, 2>{});         \
        advance(2);                                                                              \
        mt = pr.meta();                                                                          \
      }
#endif
      SWR2_RUN(0)
      SWR2_RUN(1)
      SWR2_RUN(2)
#undef SWR2_RUN
#undef SWR2_STEP
    }
  }
  if (cur != SWR_EMPTY)
#pragma unroll 1
    for (int i = 0; i < D; i++) {
      retire(slot);
      slot = slot + 1 == D ? 0 : slot + 1;
    }
  cp_async_wait<0>();
}

}  // namespace b2n

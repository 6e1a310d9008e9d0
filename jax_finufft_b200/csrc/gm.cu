// gm.cu -- "global memory" spread / interp: one thread per (sorted) point, runtime kernel width,
// any dimension.  Replaces spread_{1,2,3}d_nupts_driven / interp_{1,2,3}d_nupts_driven
// (V/src/cuda/3d/spreadinterp3d.cuh:86-135,556-608 and the 1d/2d twins).  Used for 1-D, for
// gpu_method=1, and whenever a tile does not fit shared memory.  Spreading adds with vector
// reductions executed in L2 (red.global.add.v2.f32 -> REDG.E.ADD.F32x2), never with a returning
// atomic; points are visited in bin-sorted order so neighbouring threads hit neighbouring lines.
#include "plan.h"

namespace b2n {

template <typename T> struct GmArgs {
  const PtRec<T> *rec;
  const cpx<T> *cin;
  cpx<T> *cout;
  const cpx<T> *scale;
  cpx<T> *fw;
  int64_t M, nftot;
  int nf[3];
  int dim, ns;
};

template <typename T> __device__ __forceinline__ cpx<T> cmulg(cpx<T> a, cpx<T> b) {
  cpx<T> r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}

template <typename T, bool SPREAD>
__global__ void __launch_bounds__(128) k_gm(const GmArgs<T> a,
                                             const __grid_constant__ HornerTable<T> tab) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.M) return;
  const int ns = a.ns;
  T k1[MAX_NS], k2[MAX_NS], k3[MAX_NS];
  int s1, s2 = 0, s3 = 0;
  const PtRec<T> pr = a.rec[p];
  {
    const T xr = pr.x;
    s1 = window_start(xr, ns);
    eval_kernel_rt<T>(k1, ns, T(s1) - xr, tab);
  }
  if (a.dim > 1) {
    const T yr = pr.y;
    s2 = window_start(yr, ns);
    eval_kernel_rt<T>(k2, ns, T(s2) - yr, tab);
  }
  if (a.dim > 2) {
    const T zr = pr.z;
    s3 = window_start(zr, ns);
    eval_kernel_rt<T>(k3, ns, T(s3) - zr, tab);
  }
  const int64_t j0 = pr.idx;
  const int n2 = a.dim > 1 ? ns : 1, n3 = a.dim > 2 ? ns : 1;
  cpx<T> *fw = a.fw + (int64_t)blockIdx.y * a.nftot;
  cpx<T> cv;
  cv.x = cv.y = T(0);
  if (SPREAD) {
    cv = a.cin[(int64_t)blockIdx.y * a.M + j0];
    if (a.scale) cv = cmulg<T>(cv, a.scale[j0]);
  }
  for (int c3 = 0; c3 < n3; c3++) {
    const int iz = a.dim > 2 ? wrap_once(s3 + c3, a.nf[2]) : 0;
    const T w3 = a.dim > 2 ? k3[c3] : T(1);
    for (int c2 = 0; c2 < n2; c2++) {
      const int iy = a.dim > 1 ? wrap_once(s2 + c2, a.nf[1]) : 0;
      const T w23 = a.dim > 1 ? w3 * k2[c2] : w3;
      cpx<T> *row = fw + ((int64_t)iz * a.nf[1] + iy) * a.nf[0];
      for (int c1 = 0; c1 < ns; c1++) {
        const int ix = wrap_once(s1 + c1, a.nf[0]);
        const T w = k1[c1] * w23;
        if (SPREAD) {
          cpx<T> v;
          v.x = cv.x * w;
          v.y = cv.y * w;
          red_add(row + ix, v);
        } else {
          const cpx<T> g = row[ix];
          cv.x = fma(g.x, w, cv.x);
          cv.y = fma(g.y, w, cv.y);
        }
      }
    }
  }
  if (!SPREAD) {
    if (a.scale) cv = cmulg<T>(cv, a.scale[j0]);
    a.cout[(int64_t)blockIdx.y * a.M + j0] = cv;
  }
}

template <typename T, bool SPREAD>
static int launch_gm(Plan<T> &p, const cpx<T> *cin, cpx<T> *cout, const cpx<T> *scale, cpx<T> *fw,
                     int ntr) {
  if (p.pts.M == 0) return 0;
  GmArgs<T> a;
  a.rec = p.pts.rec;
  a.cin = cin;
  a.cout = cout;
  a.scale = scale;
  a.fw = fw;
  a.M = p.pts.M;
  a.nftot = p.nftot;
  for (int d = 0; d < 3; d++) a.nf[d] = (int)p.nf[d];
  a.dim = p.dim;
  a.ns = p.ns;
  dim3 grid((unsigned)cdiv(p.pts.M, 128), (unsigned)ntr);
  k_gm<T, SPREAD><<<grid, 128, 0, p.stream>>>(a, p.tab);  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  return 0;
}

template <typename T>
int spread_gm(Plan<T> &p, const cpx<T> *c, const cpx<T> *prescale, cpx<T> *fw, int ntr) {
  return launch_gm<T, true>(p, c, nullptr, prescale, fw, ntr);
}
template <typename T>
int interp_gm(Plan<T> &p, cpx<T> *c, const cpx<T> *postscale, const cpx<T> *fw, int ntr) {
  return launch_gm<T, false>(p, nullptr, c, postscale, const_cast<cpx<T> *>(fw), ntr);
}

template int spread_gm<float>(Plan<float> &, const cpx<float> *, const cpx<float> *, cpx<float> *, int);
template int spread_gm<double>(Plan<double> &, const cpx<double> *, const cpx<double> *, cpx<double> *, int);
template int interp_gm<float>(Plan<float> &, cpx<float> *, const cpx<float> *, const cpx<float> *, int);
template int interp_gm<double>(Plan<double> &, cpx<double> *, const cpx<double> *, const cpx<double> *, int);

}  // namespace b2n

// tile_launch.cuh -- ns dispatch + launch of the tile kernels for one (T, DIM); included by the
// per-(precision, dimension) translation units so the 15 kernel widths compile in parallel.
#pragma once
#include <algorithm>

#include "tile_kernels.cuh"

namespace b2n {

template <typename T, int DIM> size_t tile_smem_dim(int ns, const int *bin);
template <typename T, int DIM>
int launch_spread_dim(Plan<T> &p, const cpx<T> *c, const cpx<T> *prescale, cpx<T> *fw, int ntr);
template <typename T, int DIM>
int launch_interp_dim(Plan<T> &p, cpx<T> *c, const cpx<T> *postscale, const cpx<T> *fw, int ntr);

template <typename T, int NS, int DIM> struct TileGeom {
  using C = TileCfg<T, NS>;
  static void dims(const int *bin, int &TX, int &TY, int &TZ) {
    TX = C::tx(bin[0]);
    TY = bin[1] + NS;
    TZ = DIM == 3 ? bin[2] + NS : 1;
  }
  static size_t spread_bytes(const int *bin) {
    int TX, TY, TZ;
    dims(bin, TX, TY, TZ);
    size_t cells = (size_t)TX * TY * TZ * (DIM == 2 ? C::NW : 1);
    return cells * sizeof(cpx<T>) + C::batch_bytes();
  }
  static size_t interp_bytes(const int *bin) {
    int TX, TY, TZ;
    dims(bin, TX, TY, TZ);
    return (size_t)TX * TY * TZ * sizeof(cpx<T>) + C::batch_bytes();
  }
};

template <typename T> inline void fill_args(Plan<T> &p, TileArgs<T> &a) {
  a.rec = p.pts.rec;
  a.bin_start = p.pts.bin_start;
  a.sp_off = p.pts.sp_off;
  a.sp_bin = p.pts.sp_bin;
  a.M = p.pts.M;
  a.nftot = p.nftot;
  for (int d = 0; d < 3; d++) {
    a.nf[d] = (int)p.nf[d];
    a.bin[d] = p.bin[d];
    a.nbin[d] = p.nbin[d];
  }
  a.nbins = p.nbins;
  a.maxsub = p.maxsub;
}

template <typename T, int NS, int DIM>
int launch_spread_ns(Plan<T> &p, const cpx<T> *c, const cpx<T> *prescale, cpx<T> *fw, int ntr) {
  TileArgs<T> a;
  fill_args(p, a);
  a.cin = c;
  a.cout = nullptr;
  a.scale = prescale;
  a.fw = fw;
  TileGeom<T, NS, DIM>::dims(p.bin, a.TX, a.TY, a.TZ);
  const size_t smem = TileGeom<T, NS, DIM>::spread_bytes(p.bin);
  dim3 grid((unsigned)p.pts.sp_cap, (unsigned)ntr);
  if constexpr (DIM == 3) {
    B2N_CUDA_OK(cudaFuncSetAttribute(k_spread3d<T, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_spread3d<T, NS><<<grid, 256, smem, p.stream>>>(a, p.tab);  B2N_LAUNCHED(1);
  } else {
    B2N_CUDA_OK(cudaFuncSetAttribute(k_spread2d<T, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_spread2d<T, NS><<<grid, 256, smem, p.stream>>>(a, p.tab);  B2N_LAUNCHED(1);
  }
  B2N_LAUNCH_OK();
  return 0;
}

template <typename T, int NS, int DIM>
int launch_interp_ns(Plan<T> &p, cpx<T> *c, const cpx<T> *postscale, const cpx<T> *fw, int ntr) {
  TileArgs<T> a;
  fill_args(p, a);
  a.cin = nullptr;
  a.cout = c;
  a.scale = postscale;
  a.fw = const_cast<cpx<T> *>(fw);
  TileGeom<T, NS, DIM>::dims(p.bin, a.TX, a.TY, a.TZ);
  const size_t smem = TileGeom<T, NS, DIM>::interp_bytes(p.bin);
  auto kern = k_interp<T, NS, DIM>;
  B2N_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)p.pts.sp_cap, (unsigned)ntr);
  kern<<<grid, 256, smem, p.stream>>>(a, p.tab);  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  return 0;
}

// compile-time recursion over the kernel width
template <typename T, int DIM, int NS> struct NsDispatch {
  static int spread(Plan<T> &p, const cpx<T> *c, const cpx<T> *pre, cpx<T> *fw, int ntr) {
    if (p.ns == NS) return launch_spread_ns<T, NS, DIM>(p, c, pre, fw, ntr);
    return NsDispatch<T, DIM, NS + 1>::spread(p, c, pre, fw, ntr);
  }
  static int interp(Plan<T> &p, cpx<T> *c, const cpx<T> *post, const cpx<T> *fw, int ntr) {
    if (p.ns == NS) return launch_interp_ns<T, NS, DIM>(p, c, post, fw, ntr);
    return NsDispatch<T, DIM, NS + 1>::interp(p, c, post, fw, ntr);
  }
  static size_t smem(int ns, const int *bin) {
    if (ns == NS)
      return std::max(TileGeom<T, NS, DIM>::spread_bytes(bin), TileGeom<T, NS, DIM>::interp_bytes(bin));
    return NsDispatch<T, DIM, NS + 1>::smem(ns, bin);
  }
};
template <typename T, int DIM> struct NsDispatch<T, DIM, MAX_NS + 1> {
  static int spread(Plan<T> &, const cpx<T> *, const cpx<T> *, cpx<T> *, int) { return B2N_ERR_METHOD_NOTVALID; }
  static int interp(Plan<T> &, cpx<T> *, const cpx<T> *, const cpx<T> *, int) { return B2N_ERR_METHOD_NOTVALID; }
  static size_t smem(int, const int *) { return 0; }
};

#define B2N_DECLARE_TILE(T, DIM)                                                                \
  template <> size_t tile_smem_dim<T, DIM>(int ns, const int *bin);                             \
  template <>                                                                                   \
  int launch_spread_dim<T, DIM>(Plan<T> & p, const cpx<T> *c, const cpx<T> *pre, cpx<T> *fw,    \
                                int ntr);                                                       \
  template <>                                                                                   \
  int launch_interp_dim<T, DIM>(Plan<T> & p, cpx<T> *c, const cpx<T> *post, const cpx<T> *fw,   \
                                int ntr);

#define B2N_INSTANTIATE_TILE(T, DIM)                                                            \
  template <> size_t tile_smem_dim<T, DIM>(int ns, const int *bin) {                            \
    return NsDispatch<T, DIM, MIN_NS>::smem(ns, bin);                                           \
  }                                                                                             \
  template <>                                                                                   \
  int launch_spread_dim<T, DIM>(Plan<T> & p, const cpx<T> *c, const cpx<T> *pre, cpx<T> *fw,    \
                                int ntr) {                                                      \
    return NsDispatch<T, DIM, MIN_NS>::spread(p, c, pre, fw, ntr);                              \
  }                                                                                             \
  template <>                                                                                   \
  int launch_interp_dim<T, DIM>(Plan<T> & p, cpx<T> *c, const cpx<T> *post, const cpx<T> *fw,   \
                                int ntr) {                                                      \
    return NsDispatch<T, DIM, MIN_NS>::interp(p, c, post, fw, ntr);                             \
  }

}  // namespace b2n

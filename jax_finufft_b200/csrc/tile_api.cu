// tile_api.cu -- dimension dispatch of the tile spread/interp kernels.
#include "tile_launch.cuh"

namespace b2n {
B2N_DECLARE_TILE(float, 2)
B2N_DECLARE_TILE(float, 3)
B2N_DECLARE_TILE(double, 2)
B2N_DECLARE_TILE(double, 3)

template <typename T> size_t tile_smem_bytes(int dim, int ns, const int *bin) {
  if (dim == 2) return tile_smem_dim<T, 2>(ns, bin);
  if (dim == 3) return tile_smem_dim<T, 3>(ns, bin);
  return 0;
}
template <typename T>
int spread_tile(Plan<T> &p, const cpx<T> *c, const cpx<T> *prescale, cpx<T> *fw, int ntr) {
  if (p.pts.M == 0 || p.pts.sp_cap == 0) return 0;
  if (p.dim == 2) return launch_spread_dim<T, 2>(p, c, prescale, fw, ntr);
  if (p.dim == 3) return launch_spread_dim<T, 3>(p, c, prescale, fw, ntr);
  return B2N_ERR_METHOD_NOTVALID;
}
template <typename T>
int interp_tile(Plan<T> &p, cpx<T> *c, const cpx<T> *postscale, const cpx<T> *fw, int ntr) {
  if (p.pts.M == 0 || p.pts.sp_cap == 0) return 0;
  if (p.dim == 2) return launch_interp_dim<T, 2>(p, c, postscale, fw, ntr);
  if (p.dim == 3) return launch_interp_dim<T, 3>(p, c, postscale, fw, ntr);
  return B2N_ERR_METHOD_NOTVALID;
}
template size_t tile_smem_bytes<float>(int, int, const int *);
template size_t tile_smem_bytes<double>(int, int, const int *);
template int spread_tile<float>(Plan<float> &, const cpx<float> *, const cpx<float> *, cpx<float> *, int);
template int spread_tile<double>(Plan<double> &, const cpx<double> *, const cpx<double> *, cpx<double> *, int);
template int interp_tile<float>(Plan<float> &, cpx<float> *, const cpx<float> *, const cpx<float> *, int);
template int interp_tile<double>(Plan<double> &, cpx<double> *, const cpx<double> *, const cpx<double> *, int);
}  // namespace b2n

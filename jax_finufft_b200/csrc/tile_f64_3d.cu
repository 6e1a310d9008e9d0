// tile_f64_3d.cu -- 3-D double instantiations (ns = 2..16) of the tile spread/interp kernels.
#include "tile_launch.cuh"
namespace b2n {
B2N_INSTANTIATE_TILE(double, 3)
}

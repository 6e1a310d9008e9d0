// tile_f32_2d.cu -- 2-D float instantiations (ns = 2..16) of the tile spread/interp kernels.
#include "tile_launch.cuh"
namespace b2n {
B2N_INSTANTIATE_TILE(float, 2)
}

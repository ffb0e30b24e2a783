// binning.cu — duplicate (Gaussian,tile) pair emission, tile-range identification, key materialisation.
//
// Reference behaviour restated (diff-langsurf-rasterizer/cuda_rasterizer/rasterizer_impl.cu):
//   duplicateWithKeys  :70-111  key = (tile_y*grid_x + tile_x) << 32 | depth bits, value = Gaussian index,
//                               tiles enumerated row-major (y outer, x inner) over getRect's rectangle
//   identifyTileRanges :116-138 ranges[tile] = (first, last+1) in the sorted list, (0,0) for untouched tiles
//
// Here the pairs are emitted in DEPTH-SORTED Gaussian order (the depth passes of the LSD sort already
// happened on the P Gaussians, see sort.cu), so only the tile id has to be carried as the sort key.
#include "kernels.cuh"

namespace lsx {

namespace {

__global__ void __launch_bounds__(256) emit_tile_pairs_kernel(int P, const uint32_t* __restrict__ sorted_idx,
                                                              const uint32_t* __restrict__ offsets,
                                                              const float2* __restrict__ means2D,
                                                              const int* __restrict__ radii, uint32_t grid_x,
                                                              uint32_t grid_y, uint32_t* __restrict__ tile_keys,
                                                              uint32_t* __restrict__ vals) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const uint32_t g = sorted_idx[k];
    const int radius = radii[g];
    if (!(radius > 0)) return;
    uint32_t off = offsets[k];
    uint2 rmin, rmax;
    tile_rect(means2D[g], radius, rmin, rmax, grid_x, grid_y);
    for (uint32_t y = rmin.y; y < rmax.y; ++y) {
        for (uint32_t x = rmin.x; x < rmax.x; ++x) {
            tile_keys[off] = y * grid_x + x;
            vals[off] = g;
            ++off;
        }
    }
}

__global__ void __launch_bounds__(256) tile_ranges_kernel(int R, const uint32_t* __restrict__ keys,
                                                          uint2* __restrict__ ranges) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t t = keys[i];
    if (i == 0) {
        ranges[t].x = 0;
    } else {
        const uint32_t prev = keys[i - 1];
        if (prev != t) {
            ranges[prev].y = i;
            ranges[t].x = i;
        }
    }
    if (i == R - 1) ranges[t].y = R;
}

__global__ void __launch_bounds__(256) debug_keys_kernel(int num_tiles, const uint2* __restrict__ ranges,
                                                         const uint32_t* __restrict__ point_list,
                                                         const float* __restrict__ depths,
                                                         uint64_t* __restrict__ keys_out) {
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const uint2 r = ranges[t];
        for (uint32_t i = r.x + threadIdx.x; i < r.y; i += blockDim.x)
            keys_out[i] = ((uint64_t)t << 32) | (uint64_t)__float_as_uint(depths[point_list[i]]);
    }
}

}  // namespace

int launch_emit_tile_pairs(int P, const uint32_t* sorted_idx, const uint32_t* offsets, const float2* means2D,
                           const int* radii, uint32_t grid_x, uint32_t grid_y, uint32_t* tile_keys, uint32_t* vals,
                           cudaStream_t stream, bool debug) {
    if (P <= 0) return 0;
    emit_tile_pairs_kernel<<<ceil_div(P, 256), 256, 0, stream>>>(P, sorted_idx, offsets, means2D, radii, grid_x, grid_y,
                                                                 tile_keys, vals);
    LSX_KERNEL_OK(stream, debug);
    return 0;
}

int launch_tile_ranges(int R, const uint32_t* sorted_tile_keys, uint2* ranges, int num_tiles, cudaStream_t stream,
                       bool debug) {
    LSX_CUDA_OK(cudaMemsetAsync(ranges, 0, (size_t)num_tiles * sizeof(uint2), stream));
    if (R <= 0) return 0;
    tile_ranges_kernel<<<ceil_div(R, 256), 256, 0, stream>>>(R, sorted_tile_keys, ranges);
    LSX_KERNEL_OK(stream, debug);
    return 0;
}

int launch_debug_keys(int num_tiles, const uint2* ranges, const uint32_t* point_list, const float* depths,
                      uint64_t* keys_out, cudaStream_t stream) {
    if (num_tiles <= 0) return 0;
    const int blocks = num_tiles < 4096 ? num_tiles : 4096;
    debug_keys_kernel<<<blocks, 256, 0, stream>>>(num_tiles, ranges, point_list, depths, keys_out);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

}  // namespace lsx

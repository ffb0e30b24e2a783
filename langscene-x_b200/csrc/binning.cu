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

// Warp-cooperative expansion: a warp owns 32 consecutive depth-sorted Gaussians, whose pairs form ONE contiguous run of
// the output (offsets are the exclusive scan in that order).  Lane j of each step writes output element j of the run
// after locating its source Gaussian by a 5-step binary search over the lanes' run-local prefixes (shuffles), so every
// store instruction covers 128 contiguous bytes.  (Thread-per-Gaussian emission wrote ~7 pairs per thread to 32
// different runs per instruction: 9.3 M store sectors for 42 MB of payload, LSU-queue bound.)
__global__ void __launch_bounds__(256) emit_tile_pairs_kernel(int P, const uint32_t* __restrict__ sorted_idx,
                                                              const uint32_t* __restrict__ offsets,
                                                              const float2* __restrict__ means2D,
                                                              const int* __restrict__ radii, uint32_t grid_x,
                                                              uint32_t grid_y, uint32_t* __restrict__ tile_keys,
                                                              uint32_t* __restrict__ vals, const uint32_t cap,
                                                              const uint32_t* __restrict__ total_pairs) {
    constexpr unsigned kFull = 0xffffffffu;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const int k0 = k - (int)lane;  // warp-uniform
    // The output arrays hold `cap` >= R slots (R rounded up to 64, or a speculative capacity chosen before R was known on
    // the host): slots R .. cap-1 get the key 0xffffffff, which sorts behind every tile id on any digit range, and are
    // never referenced by a tile range.  If R > cap (speculation too small) the surplus pairs are dropped here and the
    // host, which learns R a little later, repeats the binning with the exact size.
    {
        const uint32_t R = __ldg(total_pairs);
        for (uint32_t i = R + (uint32_t)k; i < cap; i += gridDim.x * blockDim.x) {
            tile_keys[i] = 0xffffffffu;
            vals[i] = 0u;
        }
    }
    if (k0 >= P) return;
    uint32_t g = 0, w = 0, count = 0;
    uint2 rmin = make_uint2(0u, 0u), rmax;
    if (k < P) {
        g = sorted_idx[k];
        const int radius = radii[g];
        if (radius > 0) {
            tile_rect(means2D[g], radius, rmin, rmax, grid_x, grid_y);
            w = rmax.x - rmin.x;
            count = w * (rmax.y - rmin.y);
        }
    }
    const uint32_t base = offsets[k0];
    uint32_t pre = count;  // inclusive warp scan of the counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, pre, o);
        if (lane >= (unsigned)o) pre += t;
    }
    const uint32_t total = __shfl_sync(kFull, pre, 31);
    pre -= count;  // exclusive
    for (uint32_t j0 = 0; j0 < total; j0 += 32) {
        const uint32_t j = j0 + lane;
        // source = the last lane whose prefix is <= j (lanes without pairs share the prefix of their successor)
        uint32_t s = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const uint32_t ps = __shfl_sync(kFull, pre, (int)(s + step));
            if (ps <= j) s += step;
        }
        const uint32_t sg = __shfl_sync(kFull, g, (int)s);
        const uint32_t sw = __shfl_sync(kFull, w, (int)s);
        const uint32_t sx = __shfl_sync(kFull, rmin.x, (int)s);
        const uint32_t sy = __shfl_sync(kFull, rmin.y, (int)s);
        const uint32_t sp = __shfl_sync(kFull, pre, (int)s);
        if (j < total && base + j < cap) {
            LSX_CHECK_INDEX(sg, P, "emitted Gaussian id");
            const uint32_t t = j - sp;
            const uint32_t row = t / sw, col = t - row * sw;  // row-major over the rectangle, y outer (duplicateWithKeys)
            tile_keys[base + j] = (sy + row) * grid_x + (sx + col);
            vals[base + j] = sg;
        }
    }
}

// four consecutive keys per thread (one 16-B load + the preceding key)
// R = number of slots (real pairs followed by 0xffffffff padding keys, which belong to no tile)
__global__ void __launch_bounds__(256) tile_ranges_kernel(int R, const uint32_t* __restrict__ keys,
                                                          uint2* __restrict__ ranges, const uint32_t num_tiles) {
    pdl_wait();
    const long long base = 4ll * (blockIdx.x * (long long)blockDim.x + threadIdx.x);
    if (base >= R) return;
    uint32_t k[4];
    if (base + 3 < R) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(keys + base));
        k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) k[j] = (base + j < R) ? keys[base + j] : 0u;
    }
    uint32_t prev = base > 0 ? __ldg(keys + base - 1) : 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const long long i = base + j;
        if (i < R) {
            const uint32_t t = k[j];
            if (i == 0) {
                if (t < num_tiles) ranges[t].x = 0;
            } else if (prev != t) {
                if (prev < num_tiles) ranges[prev].y = (uint32_t)i;
                if (t < num_tiles) ranges[t].x = (uint32_t)i;
            }
            if (i == R - 1 && t < num_tiles) ranges[t].y = (uint32_t)R;
            prev = t;
        }
    }
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
                           uint32_t capacity, const uint32_t* total_pairs, cudaStream_t stream, bool debug) {
    if (P <= 0) return 0;
    emit_tile_pairs_kernel<<<ceil_div(P, 256), 256, 0, stream>>>(P, sorted_idx, offsets, means2D, radii, grid_x, grid_y,
                                                                 tile_keys, vals, capacity, total_pairs);
    LSX_KERNEL_OK(stream, debug);
    return 0;
}

int launch_tile_ranges(int R, const uint32_t* sorted_tile_keys, uint2* ranges, int num_tiles, cudaStream_t stream,
                       bool debug) {
    LSX_CUDA_OK(cudaMemsetAsync(ranges, 0, (size_t)num_tiles * sizeof(uint2), stream));
    if (R <= 0) return 0;
    launch_pdl(tile_ranges_kernel, ceil_div(ceil_div(R, 4), 256), 256, 0, stream, R, sorted_tile_keys, ranges, (uint32_t)num_tiles);
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

// stats.cu — workload counters of one rendered view, counted ON THE DEVICE from the scratch a forward call left
// behind ("report S, B, R with every number").
//
//   S      sum over pixels of n_contrib: the (pixel, list entry) tests the reference's render loops execute per pass
//          (forward.cu:335-407 / backward.cu:524-676 walk the tile list up to the pixel's last contributor)
//   B      (pixel, entry) pairs that are actually blended: entry below the pixel's last contributor, power <= 0 and
//          alpha >= 1/255 — the same predicate, from the same expressions, as render_bwd.cu
//   V      (8x4 block, entry) visits of this library's backward pass: per block, its compacted list up to the
//          deepest contributor of any of its pixels
//   Vb     visits in which at least one of the 32 pixels blends
//   L      total length of the per-block compacted lists (what the forward walks at most)
//   Hmax   what-if counter for 4x4-pixel cells: sum over blocks of max(visits in which the left half blends, ... the right half)
//   Hsum   sum over blocks of both halves' blending visits
//
// One warp per 8x4 block, the same pixel <-> lane mapping as the render kernels.  Not on the timed path: bench.py and the
// tests call it once per view for the roofline's flop model.
#include "kernels.cuh"

namespace lsx {

namespace {

__global__ void __launch_bounds__(32) render_stats_kernel(const RenderParams p, unsigned long long* __restrict__ out) {
    constexpr unsigned kFull = 0xffffffffu;
    const int tile = blockIdx.x >> 3, warp = blockIdx.x & 7;
    const int tile_x = tile % p.grid_x, tile_y = tile / p.grid_x;
    const unsigned lane = threadIdx.x;
    const int px = tile_x * TILE_X + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * TILE_Y + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t pix = (size_t)py * p.W + px;

    const uint2 range = p.ranges[tile];
    const int n = (int)(range.y - range.x);
    const int last_k = inside ? (int)p.k_contrib[pix] : 0;
    unsigned long long S = inside ? (unsigned long long)p.n_contrib[pix] : 0ull;
    const int cnt = (n > 0) ? (int)p.blk_cnt[8 * tile + warp] : 0;
    const int n_eff = min(__reduce_max_sync(kFull, last_k), cnt);
    const uint32_t* list = p.blk_list + (size_t)warp * p.list_stride + range.x;

    unsigned long long B = 0ull, Vb = 0ull;
    unsigned c0 = 0u, c1 = 0u;  // visits in which the left / right 4x4 half of the block blends (what-if: 4x4-pixel cells)
    for (int e = 0; e < n_eff; ++e) {
        const uint32_t id = __ldg(p.point_list + range.x + __ldg(list + e));
        const float* rec = p.records + (size_t)id * p.rec_stride;
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(rec));
        const float2 h1 = __ldg(reinterpret_cast<const float2*>(rec + 4));
        const float dx = __fadd_rn(h0.x, -pxf), dy = __fadd_rn(h0.y, -pyf);
        const float power = splat_power(h0.z, h0.w, h1.x, dx, dy);
        const float alpha = splat_alpha(h1.y, expf(power));
        const bool blend = (e < last_k) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
        const unsigned m = __ballot_sync(kFull, blend);
        B += blend ? 1ull : 0ull;
        Vb += (m != 0u && lane == 0u) ? 1ull : 0ull;
        c0 += (m & 0x0f0f0f0fu) ? 1u : 0u;  // lanes with (lane & 7) < 4
        c1 += (m & 0xf0f0f0f0u) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        S += __shfl_xor_sync(kFull, S, o);
        B += __shfl_xor_sync(kFull, B, o);
    }
    if (lane == 0) {
        if (S) atomicAdd(out + 0, S);
        if (B) atomicAdd(out + 1, B);
        if (n_eff) atomicAdd(out + 2, (unsigned long long)n_eff);
        if (Vb) atomicAdd(out + 3, Vb);
        if (cnt) atomicAdd(out + 4, (unsigned long long)cnt);
        if (c0 | c1) {
            atomicAdd(out + 5, (unsigned long long)(c0 > c1 ? c0 : c1));  // iterations of a warp walking both halves' lists in step
            atomicAdd(out + 6, (unsigned long long)(c0 + c1));            // (4x4 cell, entry) visits in which the cell blends
        }
    }
}

}  // namespace

int launch_render_stats(const RenderParams& p, unsigned long long* out, cudaStream_t stream) {
    LSX_CUDA_OK(cudaMemsetAsync(out, 0, 8 * sizeof(unsigned long long), stream));
    const long long blocks = (long long)p.grid_x * p.grid_y * 8;
    if (blocks <= 0) return 0;
    render_stats_kernel<<<(unsigned)blocks, 32, 0, stream>>>(p, out);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

}  // namespace lsx

// cull.cu — per-(tile, list entry) sub-tile footprint masks and the per-block compacted lists built from them.
//
// The reference tests every list entry of a tile against every pixel of the tile
// (forward.cu:335-407, backward.cu:524-676): an entry is blended at a pixel only if
//     power = -0.5 (a dx^2 + c dy^2) - b dx dy <= 0   and   opacity * exp(power) >= 1/255 .
// The tile list itself is part of the bit-exact contract (radius-square rectangles of 16x16 tiles), but it
// is very conservative: at the headline configuration ~70 % of the (8x4 pixel block, entry) pairs can never
// pass the alpha test.  This kernel computes, once per forward call, an 8-bit mask per list entry — bit w set
// iff the splat MAY reach alpha >= 1/255 somewhere inside block w of its tile (block w covers pixels
// x in [8 (w&1), +8), y in [4 (w>>1), +4), the pixels owned by warp w of the render CTAs).  The render
// kernels (forward and backward) let each warp visit only the entries whose bit is set.  Skipped pairs would
// have executed `continue` in the reference, so every output (including n_contrib / final_T / out_observe)
// is unchanged.
//
// The test is exact up to a safety margin: the maximum of `power` over the block's continuous cell (widened by
// half a pixel; a convex quadratic minimised over a box: centre inside -> 0, otherwise the minimum over the four
// edges) is compared with ln(1 / (255 opacity)); the margin (1e-3 + 1e-4 * magnitude of the terms) is >= 100x
// the fp32 rounding error of either evaluation, and every comparison is written so that NaN / non-convex /
// non-finite inputs keep the bit SET (the render kernel then applies the reference's own tests).
#include "kernels.cuh"

namespace lsx {

namespace {

// One list entry: footprint mask over the 2 x 4 grid of 8x4-pixel blocks of its tile.
//
// The blocks are widened by half a pixel to the continuous cells [x0 - .5 + 8 i, +8] x [y0 - .5 + 4 j, +4], so that
// neighbouring blocks share their edges: 3 vertical lines x 4 intervals + 5 horizontal lines x 2 intervals = 22
// one-dimensional minimisations of  q(dx, dy) = a dx^2 + 2 b dx dy + c dy^2  (4 instructions each once the line's
// coefficients are set up) give the exact minimum of q over every cell whose interior does not contain the centre.
__device__ __forceinline__ unsigned footprint_mask(const float mx, const float my, const float a, const float b, const float c,
                                                   const float o, const float tx0, const float ty0) {
    const float o255 = o * 255.0f;
    // every comparison is written so that NaN / non-finite / non-convex inputs fall through to "visit everywhere"
    const bool convex = (a > 0.f) && (c > 0.f) && (a * c - b * b > 0.f) && (a < 1.0e6f) && (c < 1.0e6f) &&
                        (fabsf(mx) < 1.0e7f) && (fabsf(my) < 1.0e7f);  // => every term below is finite
    if (!convex || !(o255 < 3.0e38f)) return 0xffu;
    if (o255 < 0.999f) return 0u;  // power <= 0 for a convex form, so alpha <= opacity < 1/255 everywhere
    const float thr = -logf(o255);   // blend needs  -0.5 q >= thr
    const float inv_a = 1.0f / a, inv_c = 1.0f / c;
    // d = mean - pixel; cell boundaries in d-space (descending in i because d = mean - coordinate)
    float dxl[3], dyl[5];
#pragma unroll
    for (int i = 0; i < 3; ++i) dxl[i] = mx - (tx0 - 0.5f + 8.0f * (float)i);
#pragma unroll
    for (int j = 0; j < 5; ++j) dyl[j] = my - (ty0 - 0.5f + 4.0f * (float)j);
    // safety margin: >= 100x the fp32 rounding error of either evaluation of the form anywhere in the tile
    const float DX = fmaxf(fabsf(dxl[0]), fabsf(dxl[2])), DY = fmaxf(fabsf(dyl[0]), fabsf(dyl[4]));
    const float qscale = a * DX * DX + 2.0f * fabsf(b) * DX * DY + c * DY * DY;
    const float qmax = -2.0f * (thr - (1.0e-3f + 1.0e-4f * (qscale + fabsf(thr))));  // keep iff q_min <= qmax
    // vertical lines dx = dxl[i]: q = A + dy (B + c dy), minimised at dy = t
    float qv[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float A = a * dxl[i] * dxl[i], B = 2.0f * b * dxl[i], t = -b * dxl[i] * inv_c;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float dy = fminf(dyl[j], fmaxf(dyl[j + 1], t));  // interval [dyl[j+1], dyl[j]]
            qv[i][j] = A + dy * (B + c * dy);
        }
    }
    // horizontal lines dy = dyl[j]
    float qh[5][2];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float A = c * dyl[j] * dyl[j], B = 2.0f * b * dyl[j], t = -b * dyl[j] * inv_a;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float dx = fminf(dxl[i], fmaxf(dxl[i + 1], t));  // interval [dxl[i+1], dxl[i]]
            qh[j][i] = A + dx * (B + a * dx);
        }
    }
    unsigned m = 0u;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const int i = w & 1, j = w >> 1;
        const bool centre_inside = (dxl[i + 1] <= 0.f) && (dxl[i] >= 0.f) && (dyl[j + 1] <= 0.f) && (dyl[j] >= 0.f);
        const float qmin = fminf(fminf(qv[i][j], qv[i + 1][j]), fminf(qh[j][i], qh[j + 1][i]));
        const bool never = !centre_inside && (qmin > qmax);
        m |= never ? 0u : (1u << w);
    }
    return m;
}

// One CTA per tile.  Besides the mask byte per list entry (kept: tests and the layout query read it) the kernel writes,
// for each of the tile's eight 8x4 blocks, the COMPACTED list of the entries whose bit is set — their positions inside the
// tile's range, in list (= depth) order — at blk_list[w * list_stride + range.x ...] and its length at blk_cnt[8 tile + w].
// The render kernels walk these lists (tile_stage.cuh: ListStage): full 16-entry rounds, no per-round mask / ballot work.
// Ordered compaction per 256-entry chunk: the threads leave their mask bytes in shared memory and warp w compacts bit w of
// all 256 of them (8 ballots): it alone owns block w's list, so no prefix has to be exchanged between warps (the first version
// let every lane place its own 8 bits — 8 ballots, a 64-word exchange and 8 shuffled bases per lane: 40 % of the kernel's
// instructions and two thirds of its stall samples, profiles/r6q_ncu_helpers_c5.md).
__global__ void __launch_bounds__(256) footprint_masks_kernel(const uint2* __restrict__ ranges,
                                                              const uint32_t* __restrict__ point_list,
                                                              const float* __restrict__ records, int rec_stride,
                                                              uint32_t grid_x, uint8_t* __restrict__ masks,
                                                              uint32_t* __restrict__ blk_list, size_t list_stride,
                                                              uint32_t* __restrict__ blk_cnt, int chk_points) {
    constexpr unsigned kFull = 0xffffffffu;
    __shared__ uint8_t s_m[2][256];  // [chunk parity][entry of the chunk]: double-buffered, one barrier per chunk
    const int tile = blockIdx.x;
    const uint2 r = ranges[tile];
    const uint32_t n = r.y - r.x;
    const float tx0 = (float)((tile % grid_x) * TILE_X), ty0 = (float)((tile / grid_x) * TILE_Y);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t run = 0;  // entries of block `warp` emitted by earlier chunks (warp-uniform)
    uint32_t* my_list = blk_list + (size_t)warp * list_stride + r.x;
    // software pipeline over the 256-entry chunks: the record head of chunk c + 1 and the Gaussian index of chunk c + 2
    // are in flight while chunk c is classified and compacted (the two loads are dependent: index -> record)
    auto load_id = [&](uint32_t i) -> uint32_t {
        if (i < n) LSX_CHECK_INDEX((long long)r.x + i, list_stride, "tile list slot");
        const uint32_t id = i < n ? __ldg(point_list + r.x + i) : 0u;
        LSX_CHECK_INDEX(id, chk_points, "Gaussian id of a list entry");
        return id;
    };
    float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 h1 = make_float2(0.f, 0.f);
    {
        const uint32_t id0 = load_id(threadIdx.x);
        if (threadIdx.x < n) {
            const float* rec = records + (size_t)id0 * rec_stride;
            h0 = __ldg(reinterpret_cast<const float4*>(rec));
            h1 = __ldg(reinterpret_cast<const float2*>(rec + 4));
        }
    }
    uint32_t id_next = load_id(256u + threadIdx.x);
    unsigned parity = 0;
    for (uint32_t base = 0; base < n; base += 256, parity ^= 1u) {
        const uint32_t i = base + threadIdx.x;
        float4 nh0 = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 nh1 = make_float2(0.f, 0.f);
        if (i + 256u < n) {
            const float* rec = records + (size_t)id_next * rec_stride;
            nh0 = __ldg(reinterpret_cast<const float4*>(rec));
            nh1 = __ldg(reinterpret_cast<const float2*>(rec + 4));
        }
        id_next = load_id(i + 512u);
        unsigned m = 0u;
        if (i < n) {
            m = footprint_mask(h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, tx0, ty0);
            masks[r.x + i] = (uint8_t)m;
        }
        s_m[parity][threadIdx.x] = (uint8_t)m;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool set = (s_m[parity][j * 32 + lane] >> warp) & 1u;
            const unsigned bal = __ballot_sync(kFull, set);
            if (set) {
                LSX_CHECK_INDEX(run + __popc(bal & lt), n, "block list write");
                my_list[run + __popc(bal & lt)] = base + (uint32_t)(j * 32) + lane;
            }
            run += __popc(bal);
        }
        h0 = nh0;
        h1 = nh1;
    }
    if (lane == 0) blk_cnt[8 * tile + warp] = run;
}

}  // namespace

int launch_footprint_masks(int num_tiles, const uint2* ranges, const uint32_t* point_list, const float* records,
                           int rec_stride, uint32_t grid_x, uint8_t* masks, uint32_t* blk_list, size_t list_stride,
                           uint32_t* blk_cnt, int num_points, cudaStream_t stream, bool debug) {
    if (num_tiles <= 0) return 0;
    footprint_masks_kernel<<<num_tiles, 256, 0, stream>>>(ranges, point_list, records, rec_stride, grid_x, masks, blk_list,
                                                         list_stride, blk_cnt, num_points);
    LSX_KERNEL_OK(stream, debug);
    return 0;
}

}  // namespace lsx

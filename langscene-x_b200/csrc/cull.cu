// cull.cu — per-(tile, list entry) sub-tile footprint masks.
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
// The test is exact up to a safety margin: the maximum of `power` over the block's continuous rectangle
// (a convex quadratic minimised over a box: centre inside -> 0, otherwise the minimum over the four edges)
// is compared with ln(1 / (255 opacity)); the margin (1e-3 + 1e-4 * magnitude of the terms) is >= 100x the
// fp32 rounding error of either evaluation, and every comparison is written so that NaN / non-convex /
// non-finite inputs keep the bit SET (the render kernel then applies the reference's own tests).
#include "kernels.cuh"

namespace lsx {

namespace {

struct SplatHead {
    float mx, my, a, b, c, o;
};

// true if the splat can pass the alpha test somewhere in pixels [x0, x0+nx) x [y0, y0+ny)
__device__ __forceinline__ bool block_may_blend(const SplatHead& s, const float thr, const float inv_a, const float inv_c,
                                                const float x0, const float y0, const float nx,
                                                const float ny) {
    // offsets d = mean - pixel over the block: dx in [dxl, dxh], dy in [dyl, dyh]
    const float dxh = s.mx - x0, dxl = s.mx - (x0 + nx - 1.0f);
    const float dyh = s.my - y0, dyl = s.my - (y0 + ny - 1.0f);
    if (dxl <= 0.f && dxh >= 0.f && dyl <= 0.f && dyh >= 0.f) return true;  // centre inside the block
    const float b2 = 2.0f * s.b;
    auto q = [&](float dx, float dy) { return s.a * dx * dx + b2 * dx * dy + s.c * dy * dy; };
    // edges dx = const: minimise over dy;  edges dy = const: minimise over dx
    const float dy1 = fminf(dyh, fmaxf(dyl, -s.b * dxl * inv_c));
    const float dy2 = fminf(dyh, fmaxf(dyl, -s.b * dxh * inv_c));
    const float dx3 = fminf(dxh, fmaxf(dxl, -s.b * dyl * inv_a));
    const float dx4 = fminf(dxh, fmaxf(dxl, -s.b * dyh * inv_a));
    const float qmin = fminf(fminf(q(dxl, dy1), q(dxh, dy2)), fminf(q(dx3, dyl), q(dx4, dyh)));
    const float DX = fmaxf(fabsf(dxl), fabsf(dxh)), DY = fmaxf(fabsf(dyl), fabsf(dyh));
    const float qscale = s.a * DX * DX + fabsf(b2) * DX * DY + s.c * DY * DY;
    // max power = -0.5 qmin ; blend needs power >= thr
    const bool never = (-0.5f * qmin) < (thr - (1.0e-3f + 1.0e-4f * (qscale + fabsf(thr))));
    return !never;
}

__global__ void __launch_bounds__(256) footprint_masks_kernel(const uint2* __restrict__ ranges,
                                                              const uint32_t* __restrict__ point_list,
                                                              const float* __restrict__ records, int rec_stride,
                                                              uint32_t grid_x, uint8_t* __restrict__ masks) {
    const int tile = blockIdx.x;
    const uint2 r = ranges[tile];
    const float tx0 = (float)((tile % grid_x) * TILE_X), ty0 = (float)((tile / grid_x) * TILE_Y);
    for (uint32_t i = r.x + threadIdx.x; i < r.y; i += blockDim.x) {
        const float* rec = records + (size_t)point_list[i] * rec_stride;
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(rec));
        const float2 h1 = __ldg(reinterpret_cast<const float2*>(rec + 4));
        SplatHead s{h0.x, h0.y, h0.z, h0.w, h1.x, h1.y};
        unsigned m = 0xffu;  // default (NaN / non-finite / non-convex inputs): visit everywhere
        const float o255 = s.o * 255.0f;
        const bool convex = (s.a > 0.f) && (s.c > 0.f) && (s.a * s.c - s.b * s.b > 0.f) && (s.a < 1.0e6f) &&
                            (s.c < 1.0e6f) && (fabsf(s.mx) < 1.0e7f) && (fabsf(s.my) < 1.0e7f);  // => every term finite
        if (convex && o255 < 0.999f) {
            m = 0u;  // power <= 0 for a convex form, so alpha <= opacity < 1/255 everywhere
        } else if (convex && o255 < 3.0e38f) {
            const float thr = -logf(o255);
            const float inv_a = 1.0f / s.a, inv_c = 1.0f / s.c;
            m = 0u;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const bool keep = block_may_blend(s, thr, inv_a, inv_c, tx0 + (float)((w & 1) * 8),
                                                  ty0 + (float)((w >> 1) * 4), 8.0f, 4.0f);
                m |= keep ? (1u << w) : 0u;
            }
        }
        masks[i] = (uint8_t)m;
    }
}

}  // namespace

int launch_footprint_masks(int num_tiles, const uint2* ranges, const uint32_t* point_list, const float* records,
                           int rec_stride, uint32_t grid_x, uint8_t* masks, cudaStream_t stream, bool debug) {
    if (num_tiles <= 0) return 0;
    footprint_masks_kernel<<<num_tiles, 256, 0, stream>>>(ranges, point_list, records, rec_stride, grid_x, masks);
    LSX_KERNEL_OK(stream, debug);
    return 0;
}

}  // namespace lsx

// render_fwd.cu — per-tile front-to-back alpha compositing of colour, language feature, instance feature
// and the 5-channel normal/alpha/distance map, plus plane depth.
//
// Reference behaviour restated: diff-langsurf-rasterizer/cuda_rasterizer/forward.cu:273-431
//   power = -0.5(a dx^2 + c dy^2) - b dx dy ; skip if power > 0 ; alpha = min(0.99, o*exp(power)) ;
//   skip if alpha < 1/255 ; stop (entry NOT blended) when T(1-alpha) < 1e-4 ; X += feat*alpha*T ;
//   out_observe[id] += 1 while T > 0.5 ; colour gets T*bg, features/maps do not ;
//   plane_depth = A4 / -(A0*ray.x + A1*ray.y + A2 + 1e-8) evaluated in double.
//
// B200 design
//   * one CTA per 16x16 tile, one thread per pixel; a warp owns a compact 8x4 pixel block (32-B row
//     segments -> whole-sector image stores, better splat/warp coherence than 16x2 rows);
//   * the tile's list is staged by the TMA engine (tile_stage.cuh): one contiguous, sector-aligned
//     record per entry holding xy/conic/opacity AND all blended channels, double-buffered so staging
//     overlaps blending; the blend loop reads them as warp-broadcast LDS.128;
//   * the loop is kept warp-converged: per entry one ballot decides whether any lane blends, the
//     out_observe count is warp-aggregated (one integer atomic per warp instead of one per pixel),
//     and warps/CTAs whose pixels are all saturated leave via warp/CTA votes;
//   * thresholds (alpha < 1/255, T < 1e-4, T > 0.5) use the same fp32 expressions and full-precision
//     expf as the reference so that n_contrib / final_T / out_observe are reproduced exactly.
#include "kernels.cuh"
#include "tile_stage.cuh"

namespace lsx {

namespace {

constexpr unsigned kFull = 0xffffffffu;

template <int CT4>
__global__ void __launch_bounds__(TILE_PIXELS) render_fwd_kernel(const RenderParams p) {
    constexpr int RS = (REC_HEAD + CT4 + 7) & ~7;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileStage<RS> stage;
    stage.init(smem_raw);

    const int tile = blockIdx.x;
    const int tile_x = tile % p.grid_x, tile_y = tile / p.grid_x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = tile_x * TILE_X + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * TILE_Y + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;

    const uint2 range = p.ranges[tile];
    const int n = (int)(range.y - range.x);
    const int nbatch = (n + STAGE_BATCH - 1) / STAGE_BATCH;

    float T = 1.0f;
    uint32_t last_contributor = 0;
    bool done = !inside;
    float acc[CT4];
#pragma unroll
    for (int c = 0; c < CT4; ++c) acc[c] = 0.f;

    auto entry_of = [&](int b) -> long long {
        const int e = b * STAGE_BATCH + (int)threadIdx.x;
        return e < n ? (long long)range.x + e : -1;
    };

    int issued = 0, consumed = 0;
    if (nbatch > 0) {
        stage.issue(0, entry_of(0), p.point_list, p.records);
        issued = 1;
    }
    for (int b = 0; b < nbatch; ++b) {
        if (b + 1 < nbatch) {
            stage.issue(b + 1, entry_of(b + 1), p.point_list, p.records);
            issued = b + 2;
        }
        stage.wait(b);
        consumed = b + 1;

        const float* rb = stage.rec_buf(b);
        const int* ib = stage.id_buf(b);
        const int cnt = min(STAGE_BATCH, n - b * STAGE_BATCH);
        if (!__all_sync(kFull, done)) {
            for (int j = 0; j < cnt; ++j) {
                const float4 h0 = *reinterpret_cast<const float4*>(rb + j * RS);      // x, y, conic.x, conic.y
                const float2 h1 = *reinterpret_cast<const float2*>(rb + j * RS + 4);  // conic.z, opacity
                bool blend = false;
                float alpha = 0.f, test_T = 0.f;
                if (!done) {
                    const float dx = __fadd_rn(h0.x, -pxf), dy = __fadd_rn(h0.y, -pyf);
                    const float power = splat_power(h0.z, h0.w, h1.x, dx, dy);
                    if (!(power > 0.0f)) {
                        alpha = splat_alpha(h1.y, expf(power));
                        if (!(alpha < 1.0f / 255.0f)) {
                            test_T = __fmul_rn(T, __fadd_rn(1.0f, -alpha));
                            if (test_T < 0.0001f)
                                done = true;
                            else
                                blend = true;
                        }
                    }
                }
                const unsigned bm = __ballot_sync(kFull, blend);
                if (bm == 0) {
                    if (__all_sync(kFull, done)) break;
                    continue;
                }
                if (blend) {
                    const float w = alpha * T;
                    const float4* ch = reinterpret_cast<const float4*>(rb + j * RS + REC_HEAD);
#pragma unroll
                    for (int q = 0; q < CT4 / 4; ++q) {
                        const float4 f = ch[q];
                        acc[4 * q + 0] += f.x * w;
                        acc[4 * q + 1] += f.y * w;
                        acc[4 * q + 2] += f.z * w;
                        acc[4 * q + 3] += f.w * w;
                    }
                }
                const unsigned om = __ballot_sync(kFull, blend && (T > 0.5f));
                if (om != 0 && lane == 0) atomicAdd(&p.out_observe[ib[j]], __popc(om));
                if (blend) {
                    T = test_T;
                    last_contributor = (uint32_t)(b * STAGE_BATCH + j + 1);
                }
            }
        }
        // CTA-wide vote; doubles as the barrier that frees buffer (b & 1) for batch b + 2
        if (__syncthreads_and(done)) break;
    }
    // never leave with bulk copies still in flight into this CTA's shared memory
    for (int b = consumed; b < issued; ++b) stage.wait(b);

    if (inside) {
        const size_t HW = (size_t)p.H * p.W;
        const size_t pix = (size_t)py * p.W + px;
        p.final_T[pix] = T;
        p.n_contrib[pix] = last_contributor;
#pragma unroll
        for (int c = 0; c < 3; ++c) p.out_color[c * HW + pix] = acc[c] + T * p.bg[c];
        int base = 3;
        if (p.include_feature) {
#pragma unroll
            for (int c = 0; c < CT4; ++c) {  // unrolled with compile-time register indices
                const int f = c - 3;
                if (f >= 0 && f < p.F) p.out_language_feature[f * HW + pix] = acc[c];
                const int fi = c - 3 - p.F;
                if (fi >= 0 && fi < p.Fi) p.out_language_feature_instance[fi * HW + pix] = acc[c];
            }
            base += p.F + p.Fi;
        }
        if (p.render_geo) {
            float am[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < CT4; ++c) {
                const int k = c - base;
                if (k >= 0 && k < 5) {
                    p.out_all_map[k * HW + pix] = acc[c];
#pragma unroll
                    for (int m = 0; m < 5; ++m)
                        if (m == k) am[m] = acc[c];
                }
            }
            const float rayx = (pxf - p.W * 0.5f) / p.focal_x;
            const float rayy = (pyf - p.H * 0.5f) / p.focal_y;
            p.out_plane_depth[pix] = am[4] / -(am[0] * rayx + am[1] * rayy + am[2] + 1.0e-8);
        }
    }
}

template <int CT4>
int launch_fwd_t(const RenderParams& p, cudaStream_t stream, bool debug) {
    constexpr int RS = (REC_HEAD + CT4 + 7) & ~7;
    const size_t smem = TileStage<RS>::kSmemBytes;
    static bool configured = false;  // per-instantiation; benign race (same value)
    if (!configured) {
        LSX_CUDA_OK(cudaFuncSetAttribute(render_fwd_kernel<CT4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int tiles = (int)(p.grid_x * p.grid_y);
    render_fwd_kernel<CT4><<<tiles, TILE_PIXELS, smem, stream>>>(p);
    LSX_KERNEL_OK(stream, debug);
    return 0;
}

}  // namespace

int launch_render_fwd(const RenderParams& p, cudaStream_t stream, bool debug) {
    switch (round_up4(p.n_channels)) {
        case 4: return launch_fwd_t<4>(p, stream, debug);
        case 8: return launch_fwd_t<8>(p, stream, debug);
        case 12: return launch_fwd_t<12>(p, stream, debug);
        case 16: return launch_fwd_t<16>(p, stream, debug);
        case 20: return launch_fwd_t<20>(p, stream, debug);
        case 24: return launch_fwd_t<24>(p, stream, debug);
        case 28: return launch_fwd_t<28>(p, stream, debug);
        case 32: return launch_fwd_t<32>(p, stream, debug);
        case 36: return launch_fwd_t<36>(p, stream, debug);
        case 40: return launch_fwd_t<40>(p, stream, debug);
        default:
            set_error("unsupported number of blended channels: %d (max 40)", p.n_channels);
            return -1;
    }
}

}  // namespace lsx

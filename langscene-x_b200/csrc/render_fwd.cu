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
//   * ONE WARP PER CTA: warp = one 8x4 pixel block of a 16x16 tile (32-B row segments -> whole-sector image
//     stores); 8 consecutive CTAs share a tile list.  No CTA-wide barrier exists, finished blocks free their
//     slot immediately, and up to 32 blocks are resident per SM;
//   * each warp walks its block's COMPACTED list — the entries whose footprint reaches the block (cull.cu,
//     tile_stage.cuh: ListStage): ~70 % of the (block, entry) pairs of the reference's loop are never touched and
//     every 16-entry round is full.  Staging: per-lane 16-B asynchronous copies (two lanes per record) of one
//     contiguous sector-aligned record per entry holding xy/conic/opacity AND all blended channels,
//     double-buffered so the copies overlap blending; the blend loop reads them as warp-broadcast LDS.128;
//   * no per-pixel `done` flag: a terminated pixel continues with T = 0, for which the reference's own
//     test (T (1 - alpha) < 1e-4) keeps failing, and its final transmittance is parked in a second
//     register; the out_observe count is warp-aggregated (one integer atomic per warp and entry);
//   * thresholds (alpha < 1/255, T < 1e-4, T > 0.5) use the same fp32 expressions and full-precision
//     expf as the reference so that n_contrib / final_T / out_observe are reproduced exactly.
#include "kernels.cuh"
#include "tile_stage.cuh"


namespace lsx {

namespace {

constexpr unsigned kFull = 0xffffffffu;

// At 24 / 28 blended channels the accumulation runs on packed pairs (FFMA2: two fp32 FMAs per issue slot, each rounded like
// fmaf, so the image does not change): the kernel is bound by instruction issue and the FMAs are 40 % of a blended visit.
// C3 (27 channels): 0.909 -> 0.851 ms; at 16 channels the packed form is 3 % slower (profiles/r6p_ffma2_ab.log), so the
// narrower instantiations keep the scalar form (and the wider ones, which would spill at 64 registers).  32 CTAs / SM are
// requested for the packed form: ptxas otherwise settles on 78 registers instead of the 64 that fit.
template <int CT4>
__global__ void __launch_bounds__(32, (CT4 == 24 || CT4 == 28) ? 32 : 0) render_fwd_kernel(const RenderParams p) {
    constexpr bool kPacked = (CT4 == 24 || CT4 == 28);
    constexpr int RS = (REC_HEAD + CT4 + 7) & ~7;
    using Stage = ListStage<RS>;
    constexpr int CHUNK = Stage::CHUNK;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int tile = blockIdx.x >> 3, warp = blockIdx.x & 7;
    const int tile_x = tile % p.grid_x, tile_y = tile / p.grid_x;
    const int lane = threadIdx.x;
    const int px = tile_x * TILE_X + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * TILE_Y + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;

    // this block's compacted list: the entries of the tile's range whose footprint reaches the block (cull.cu)
    const uint2 range = p.ranges[tile];
    const int cnt = (range.y > range.x) ? (int)p.blk_cnt[8 * tile + warp] : 0;
    const uint32_t* list = p.blk_list + (size_t)warp * p.list_stride + range.x;
    const int nrounds = (cnt + CHUNK - 1) / CHUNK;

    // A live pixel carries its transmittance in T.  When the reference would set `done`, T is parked in
    // T_final and T becomes 0: from then on T * (1 - alpha) < 1e-4 holds for every candidate, i.e. the
    // pixel keeps "terminating" without any state change — exactly the reference's behaviour of ignoring it.
    float T = inside ? 1.0f : 0.0f;
    float T_final = 1.0f;
    uint32_t last_k = 0;  // elements of the compacted list up to and including this pixel's last contributor
    f32x2 acc2[kPacked ? CT4 / 2 : 1];  // kPacked: channels (2 j, 2 j + 1)
    float acc[CT4];
#pragma unroll
    for (int j = 0; j < (kPacked ? CT4 / 2 : 1); ++j) acc2[j] = 0ull;
#pragma unroll
    for (int c = 0; c < CT4; ++c) acc[c] = 0.f;

    Stage stage;
    const bool any_live = !__all_sync(kFull, T == 0.0f);
    if (nrounds > 0 && any_live) {
        LSX_CHECK_INDEX(cnt, (long long)(range.y - range.x) + 1, "block list length");
        LSX_CHECK_INDEX((long long)range.y - 1, p.R, "tile range end");
        stage.start(smem_raw, list, p.point_list + range.x, p.records, 0, 1, cnt, (int)(range.y - range.x), p.P);
        for (int r = 0; r < nrounds; ++r) {
            stage.advance(r);
            const int m = stage.round_size(r);
            uint32_t ra = stage.rec_addr(r & 1);
            uint32_t ia = stage.ids_addr(r & 1);
            for (int s_ = 0; s_ < m; ++s_) {
                const float4 h0 = lds128(ra);      // x, y, conic.x, conic.y
                const float2 h1 = lds64(ra + 16);  // conic.z, opacity
                const float dx = __fadd_rn(h0.x, -pxf), dy = __fadd_rn(h0.y, -pyf);
                const float power = splat_power(h0.z, h0.w, h1.x, dx, dy);
                const float alpha = splat_alpha(h1.y, expf(power));
                const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -alpha));
                const bool cand = !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                const bool blend = cand && !(test_T < 0.0001f);
                if (cand && !blend && T != 0.0f) {  // the reference's `done = true` (entry NOT blended)
                    T_final = T;
                    T = 0.0f;
                }
                const unsigned om = __ballot_sync(kFull, blend && (T > 0.5f));
                if (om != 0 && lane == 0) atomicAdd(&p.out_observe[lds32i(ia)], __popc(om));
                if (blend) {
                    const float w = alpha * T;
                    if constexpr (kPacked) {
                        const f32x2 w2 = pack2(w, w);
                        f32x2 f[CT4 / 2];
                        lds_row_pairs<CT4 / 4>(ra + REC_HEAD * 4, f);
#pragma unroll
                        for (int j = 0; j < CT4 / 2; ++j) acc2[j] = fma2(f[j], w2, acc2[j]);
                    } else {
                        float4 f[CT4 / 4];
                        lds_row<CT4 / 4>(ra + REC_HEAD * 4, f);
#pragma unroll
                        for (int q = 0; q < CT4 / 4; ++q) {
                            acc[4 * q + 0] += f[q].x * w;
                            acc[4 * q + 1] += f[q].y * w;
                            acc[4 * q + 2] += f[q].z * w;
                            acc[4 * q + 3] += f[q].w * w;
                        }
                    }
                    T = test_T;
                    last_k = (uint32_t)(r * CHUNK + s_ + 1);
                }
                ra += Stage::kRecBytes;
                ia += 4;
            }
            if (__all_sync(kFull, T == 0.0f)) break;  // every pixel of the block is saturated
        }
        stage.drain();  // never leave with copies still in flight into this CTA's shared memory
    }
    if (T != 0.0f) T_final = T;  // never terminated
    if constexpr (kPacked) {
#pragma unroll
        for (int j = 0; j < CT4 / 2; ++j) {
            const float2 a = unpack2(acc2[j]);
            acc[2 * j] = a.x;
            acc[2 * j + 1] = a.y;
        }
    }

    if (inside) {
        const size_t HW = (size_t)p.H * p.W;
        const size_t pix = (size_t)py * p.W + px;
        T = T_final;
        p.final_T[pix] = T;
        // the reference's n_contrib: position of the last contributor in the TILE list + 1
        p.n_contrib[pix] = last_k ? __ldg(list + (last_k - 1)) + 1u : 0u;
        p.k_contrib[pix] = last_k;
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
    const size_t smem = ListStage<RS>::kSmemBytes;  // two 16-record buffers: ~5 KB, so 32 blocks fit on an SM
    const long long blocks = (long long)p.grid_x * p.grid_y * 8;
    render_fwd_kernel<CT4><<<(unsigned)blocks, 32, smem, stream>>>(p);
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

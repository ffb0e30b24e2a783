// render_bwd.cu — per-tile back-to-front gradient pass of the alpha compositing.
//
// Reference behaviour restated: diff-langsurf-rasterizer/cuda_rasterizer/backward.cu:399-677
//   per pixel, traverse the tile list backwards from the last contributor; recompute G = exp(power),
//   alpha; T <- T/(1-alpha); for every blended channel
//       dL/dfeat[g][ch] += alpha*T * dL/dpix[ch]
//       dL/dalpha       += (feat[g][ch] - accum_behind[ch]) * dL/dpix[ch]
//   then dL/dalpha *= T, background term, and the chain into mean2D (+abs), conic {x,y,w}, opacity.
//   plane-depth upstream gradient is folded into the all_map upstream gradient first (:497-503).
//   The reference issues one global fp32 atomicAdd per (pixel, Gaussian, value): Ct + 8 of them.
//
// B200 design
//   * the per-channel "accumulated colour behind" recurrence is linear, so its dot product with the
//     pixel's upstream gradient is carried as ONE scalar:  A <- a_prev * s_prev + (1 - a_prev) * A  with
//     s = <feat[g], dL/dpix>.  That removes 2*Ct registers and ~3*Ct flops per blend compared with the
//     per-channel form, and leaves the pixel's upstream gradient vector as the only wide state.
//   * all 32 lanes of a warp walk the list in lock-step, so the Ct+8 per-lane contributions of one
//     Gaussian are summed ACROSS the warp before touching memory: a transposing butterfly (each step
//     exchanges half of the remaining values) reduces N values in N-1 shuffles and leaves value k in
//     lane k, which then issues a single RED.ADD.F32 — <= Ct+8 atomics per (warp, Gaussian) instead of
//     32*(Ct+8), i.e. up to 32x fewer L2 atomics than the reference;
//   * CTA-level skip of the list tail beyond the tile's deepest contributor, TMA-staged records
//     (tile_stage.cuh), warp-ballot skip of Gaussians no lane blends.
#include "kernels.cuh"
#include "tile_stage.cuh"

namespace lsx {

namespace {

constexpr unsigned kFull = 0xffffffffu;

// Transposing butterfly: on entry every lane holds N partial values v[0..N); on exit v[0] of lane L holds the
// warp-wide sum of value (L >> (5 - log2 N)).  N-1 + (5 - log2 N) shuffles in total.
template <int N, int OFF>
struct WarpTransposeReduce {
    static __device__ __forceinline__ void run(float* v, const unsigned lane) {
        if constexpr (N > 1) {
            const bool upper = (lane & OFF) != 0;
#pragma unroll
            for (int i = 0; i < N / 2; ++i) {
                const float send = upper ? v[i] : v[i + N / 2];
                const float keep = upper ? v[i + N / 2] : v[i];
                v[i] = keep + __shfl_xor_sync(kFull, send, OFF);
            }
            if constexpr (OFF > 1) WarpTransposeReduce<N / 2, OFF / 2>::run(v, lane);
        } else {
            v[0] += __shfl_xor_sync(kFull, v[0], OFF);
            if constexpr (OFF > 1) WarpTransposeReduce<1, OFF / 2>::run(v, lane);
        }
    }
};

template <int N>
struct Log2 {
    static constexpr int value = 1 + Log2<N / 2>::value;
};
template <>
struct Log2<1> {
    static constexpr int value = 0;
};

struct LaneTarget {  // where a lane's reduced value goes: base + id*stride
    float* base;
    int stride;
};

// slot s of the per-Gaussian value vector: [0,CT4) blended channels, [CT4, CT4+8) geometry terms
template <int CT4>
__device__ __forceinline__ LaneTarget slot_target(const RenderParams& p, int s) {
    LaneTarget t{nullptr, 0};
    if (s < 0) return t;
    if (s < CT4) {
        int c = s;
        if (c < 3) return LaneTarget{p.dL_dcolors + c, 3};
        c -= 3;
        if (p.include_feature) {
            if (c < p.F) return LaneTarget{p.dL_dlanguage_feature + c, p.F};
            c -= p.F;
            if (c < p.Fi) return LaneTarget{p.dL_dlanguage_feature_instance + c, p.Fi};
            c -= p.Fi;
        }
        if (p.render_geo && c < 5) return LaneTarget{p.dL_dall_map + c, 5};
        return t;
    }
    switch (s - CT4) {
        case 0: return LaneTarget{p.dL_dmean2D + 0, 3};
        case 1: return LaneTarget{p.dL_dmean2D + 1, 3};
        case 2: return LaneTarget{p.dL_dmean2D_abs + 0, 3};
        case 3: return LaneTarget{p.dL_dmean2D_abs + 1, 3};
        case 4: return LaneTarget{p.dL_dconic + 0, 4};
        case 5: return LaneTarget{p.dL_dconic + 1, 4};
        case 6: return LaneTarget{p.dL_dconic + 3, 4};
        case 7: return LaneTarget{p.dL_dopacity, 1};
        default: return t;
    }
}

template <int CT4>
__global__ void __launch_bounds__(TILE_PIXELS) render_bwd_kernel(const RenderParams p) {
    constexpr int RS = (REC_HEAD + CT4 + 7) & ~7;
    constexpr int NV = CT4 + 8;                          // values reduced per Gaussian
    constexpr int N1 = NV <= 16 ? 16 : 32;               // first butterfly group (zero padded)
    constexpr int REM = NV > 32 ? NV - 32 : 0;
    constexpr int N2 = REM == 0 ? 0 : (REM <= 4 ? 4 : (REM <= 8 ? 8 : 16));
    constexpr int NVP = N1 + N2;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_tile_max;
    TileStage<RS> stage;
    if (threadIdx.x == 0) s_tile_max = 0;
    stage.init(smem_raw);  // contains a __syncthreads

    const int tile = blockIdx.x;
    const int tile_x = tile % p.grid_x, tile_y = tile / p.grid_x;
    const unsigned lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int px = tile_x * TILE_X + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * TILE_Y + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t HW = (size_t)p.H * p.W;
    const size_t pix = (size_t)py * p.W + px;

    const uint2 range = p.ranges[tile];
    const int n = (int)(range.y - range.x);

    // ---- per-pixel state -------------------------------------------------------------------------
    const float T_final = inside ? p.final_T[pix] : 0.f;
    float T = T_final;
    const int last_contributor = inside ? (int)p.n_contrib[pix] : 0;

    float g[CT4];  // upstream gradient of every blended channel of this pixel
#pragma unroll
    for (int c = 0; c < CT4; ++c) g[c] = 0.f;
    float bg_dot = 0.f;
    if (inside) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            g[c] = p.dL_dout_color[c * HW + pix];
            bg_dot += p.bg[c] * g[c];
        }
        int base = 3;
        if (p.include_feature) {
#pragma unroll
            for (int c = 3; c < CT4; ++c) {
                const int f = c - 3;
                if (f < p.F) g[c] = p.dL_dout_language_feature[f * HW + pix];
                const int fi = c - 3 - p.F;
                if (fi >= 0 && fi < p.Fi) g[c] = p.dL_dout_language_feature_instance[fi * HW + pix];
            }
            base += p.F + p.Fi;
        }
        if (p.render_geo) {
            float am[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) am[k] = p.dL_dout_all_map[k * HW + pix];
            // fold the plane-depth gradient into the map gradient (ray in double like the reference)
            const float rayx = (pxf - p.W * 0.5) / p.focal_x;
            const float rayy = (pyf - p.H * 0.5) / p.focal_y;
            const float nx = p.all_map_pixels[pix], ny = p.all_map_pixels[HW + pix], nz = p.all_map_pixels[2 * HW + pix];
            const float distance = p.all_map_pixels[4 * HW + pix];
            const float tmp = (nx * rayx + ny * rayy + nz + 1.0e-8);
            const float gd = p.dL_dout_plane_depth[pix];
            am[4] += (-gd / tmp);
            am[0] += gd * (distance / (tmp * tmp) * rayx);
            am[1] += gd * (distance / (tmp * tmp) * rayy);
            am[2] += gd * (distance / (tmp * tmp));
#pragma unroll
            for (int c = 3; c < CT4; ++c) {
                const int k = c - base;
#pragma unroll
                for (int m = 0; m < 5; ++m)
                    if (m == k) g[c] = am[m];
            }
        }
    }

    // deepest contributor of the tile: list entries at or beyond it are never blended by any pixel
    {
        const int wmax = __reduce_max_sync(kFull, last_contributor);
        if (lane == 0 && wmax > 0) atomicMax(&s_tile_max, wmax);
    }
    __syncthreads();
    const int n_eff = min(s_tile_max, n);
    const int nbatch = (n_eff + STAGE_BATCH - 1) / STAGE_BATCH;

    // per-lane atomic targets after the butterfly
    const LaneTarget tgt1 = slot_target<CT4>(p, (int)(lane >> (5 - Log2<N1>::value)));
    const bool own1 = (lane & ((1u << (5 - Log2<N1>::value)) - 1u)) == 0 && tgt1.base != nullptr;
    LaneTarget tgt2{nullptr, 0};
    bool own2 = false;
    if constexpr (N2 > 0) {
        const int s2 = (int)(lane >> (5 - Log2<(N2 > 0 ? N2 : 1)>::value));
        tgt2 = slot_target<CT4>(p, s2 < REM ? 32 + s2 : -1);
        own2 = (lane & ((1u << (5 - Log2<(N2 > 0 ? N2 : 1)>::value)) - 1u)) == 0 && tgt2.base != nullptr;
    }

    float A = 0.f;           // <accumulated colour behind, upstream gradient>
    float last_alpha = 0.f;
    float last_s = 0.f;
    const float ddelx_dx = 0.5 * p.W;
    const float ddely_dy = 0.5 * p.H;

    auto entry_of = [&](int b) -> long long {
        const int e = n_eff - 1 - (b * STAGE_BATCH + (int)threadIdx.x);
        return e >= 0 ? (long long)range.x + e : -1;
    };

    if (nbatch > 0) stage.issue(0, entry_of(0), p.point_list, p.records);
    for (int b = 0; b < nbatch; ++b) {
        if (b + 1 < nbatch) stage.issue(b + 1, entry_of(b + 1), p.point_list, p.records);
        stage.wait(b);
        const float* rb = stage.rec_buf(b);
        const int* ib = stage.id_buf(b);
        const int cnt = min(STAGE_BATCH, n_eff - b * STAGE_BATCH);
        const int e0 = n_eff - 1 - b * STAGE_BATCH;  // list index of slot 0

        for (int j = 0; j < cnt; ++j) {
            const int e = e0 - j;
            const float4 h0 = *reinterpret_cast<const float4*>(rb + j * RS);
            const float2 h1 = *reinterpret_cast<const float2*>(rb + j * RS + 4);
            bool blend = false;
            float G = 0.f, alpha = 0.f, dx = 0.f, dy = 0.f;
            if (e < last_contributor) {
                dx = __fadd_rn(h0.x, -pxf);
                dy = __fadd_rn(h0.y, -pyf);
                const float power = splat_power(h0.z, h0.w, h1.x, dx, dy);
                if (!(power > 0.0f)) {
                    G = expf(power);
                    alpha = splat_alpha(h1.y, G);
                    blend = !(alpha < 1.0f / 255.0f);
                }
            }
            if (__ballot_sync(kFull, blend) == 0) continue;

            float v[NVP];
#pragma unroll
            for (int k = 0; k < NVP; ++k) v[k] = 0.f;
            if (blend) {
                T = __fdiv_rn(T, __fadd_rn(1.f, -alpha));
                const float w = alpha * T;
                const float4* ch = reinterpret_cast<const float4*>(rb + j * RS + REC_HEAD);
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < CT4 / 4; ++q) {
                    const float4 f = ch[q];
                    s += f.x * g[4 * q + 0];
                    s += f.y * g[4 * q + 1];
                    s += f.z * g[4 * q + 2];
                    s += f.w * g[4 * q + 3];
                    v[4 * q + 0] = w * g[4 * q + 0];
                    v[4 * q + 1] = w * g[4 * q + 1];
                    v[4 * q + 2] = w * g[4 * q + 2];
                    v[4 * q + 3] = w * g[4 * q + 3];
                }
                A = last_alpha * last_s + (1.f - last_alpha) * A;
                last_s = s;
                float dL_dalpha = (s - A) * T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;

                const float dL_dG = h1.y * dL_dalpha;
                const float gdx = G * dx;
                const float gdy = G * dy;
                const float dG_ddelx = -gdx * h0.z - gdy * h0.w;
                const float dG_ddely = -gdy * h1.x - gdx * h0.w;
                const float mx = dL_dG * dG_ddelx * ddelx_dx;
                const float my = dL_dG * dG_ddely * ddely_dy;
                v[CT4 + 0] = mx;
                v[CT4 + 1] = my;
                v[CT4 + 2] = fabsf(mx);
                v[CT4 + 3] = fabsf(my);
                v[CT4 + 4] = -0.5f * gdx * dx * dL_dG;
                v[CT4 + 5] = -0.5f * gdx * dy * dL_dG;
                v[CT4 + 6] = -0.5f * gdy * dy * dL_dG;
                v[CT4 + 7] = G * dL_dalpha;
            }

            const int id = ib[j];
            WarpTransposeReduce<N1, 16>::run(v, lane);
            if (own1 && v[0] != 0.f) atomicAdd(tgt1.base + (size_t)id * tgt1.stride, v[0]);
            if constexpr (N2 > 0) {
                WarpTransposeReduce<N2, 16>::run(v + N1, lane);
                if (own2 && v[N1] != 0.f) atomicAdd(tgt2.base + (size_t)id * tgt2.stride, v[N1]);
            }
        }
        __syncthreads();  // frees buffer (b & 1) for batch b + 2
    }
}

template <int CT4>
int launch_bwd_t(const RenderParams& p, cudaStream_t stream, bool debug) {
    constexpr int RS = (REC_HEAD + CT4 + 7) & ~7;
    const size_t smem = TileStage<RS>::kSmemBytes;
    static bool configured = false;
    if (!configured) {
        LSX_CUDA_OK(cudaFuncSetAttribute(render_bwd_kernel<CT4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int tiles = (int)(p.grid_x * p.grid_y);
    render_bwd_kernel<CT4><<<tiles, TILE_PIXELS, smem, stream>>>(p);
    LSX_KERNEL_OK(stream, debug);
    return 0;
}

}  // namespace

int launch_render_bwd(const RenderParams& p, cudaStream_t stream, bool debug) {
    switch (round_up4(p.n_channels)) {
        case 4: return launch_bwd_t<4>(p, stream, debug);
        case 8: return launch_bwd_t<8>(p, stream, debug);
        case 12: return launch_bwd_t<12>(p, stream, debug);
        case 16: return launch_bwd_t<16>(p, stream, debug);
        case 20: return launch_bwd_t<20>(p, stream, debug);
        case 24: return launch_bwd_t<24>(p, stream, debug);
        case 28: return launch_bwd_t<28>(p, stream, debug);
        case 32: return launch_bwd_t<32>(p, stream, debug);
        case 36: return launch_bwd_t<36>(p, stream, debug);
        case 40: return launch_bwd_t<40>(p, stream, debug);
        default:
            set_error("unsupported number of blended channels: %d (max 40)", p.n_channels);
            return -1;
    }
}

}  // namespace lsx

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
//     pixel's upstream gradient is carried as ONE scalar:  B <- alpha * s + (1 - alpha) * B  with
//     s = <feat[g], dL/dpix>, and dL/dalpha = (s - B_behind) * T.  That removes 2*Ct registers and ~3*Ct flops
//     per blend compared with the per-channel form.  (Its rounding differs from the reference's per-channel
//     differences by O(1e-7 |s| / |s - B|); see the tolerance note in tests/test_parity_gpu.py.)
//   * ONE WARP PER CTA: warp = one 8x4 pixel block of a tile; it walks its block's compacted list (cull.cu; staged by
//     per-lane asynchronous copies, double-buffered, tile_stage.cuh: ListStage) back to front, starting below the
//     block's deepest contributor (render_fwd.cu leaves every pixel's depth in list elements).  No CTA-wide barrier.
//   * the channel gradients of one entry, dL/dfeat[c] = sum_pix (alpha T)_pix * dL/dpix[c], are an outer
//     product over the warp's 32 pixels.  Each lane keeps TWO views of the block's upstream gradient in
//     registers: its own pixel's channel vector (for s) and, transposed, channel `lane` of all 32 pixels.
//     Per entry the 32 weights alpha*T are exchanged through 128 B of shared memory and lane c evaluates
//     its channel's sum with 8 broadcast LDS.128 + 32 FFMA — no shuffles, no selects — then issues ONE
//     RED.ADD.F32 into the packed per-Gaussian gradient record (the Ct lanes hit consecutive addresses).
//     The 8 geometry terms (mean2D, |mean2D|, conic, opacity) are exchanged through shared memory the same way: when
//     at most 24 channels are blended (the reference's own 3-d feature configuration) the idle lanes 24..31 sum one
//     term each inside the outer-product phase; otherwise every lane sums 8 pixels of one term (2 LDS.128 + 7 FADD)
//     and two shuffles, issued one entry late so that they overlap the next entry's rcp -> dot-product chain, finish it.
//     The constant factors of the geometry terms (0.5 W, 0.5 H, -0.5) are left to preprocess_bwd_kernel (once per
//     Gaussian instead of once per visit), the blend section is straight-line code with selects (the warp issues it
//     anyway once one lane blends), 1 / (1 - alpha) is the raw rcp.approx (<= 1 ulp; T is only used at 1e-4): together
//     212 -> 193 SASS instructions per blended visit, 2.17 -> 2.01 ms at C3.
//     The reference issues Ct + 8 global atomics per (pixel, Gaussian); this kernel issues Ct + 8 per
//     (32-pixel block, Gaussian), all into one contiguous 144-B record.
//   * warp-ballot skip of entries no lane blends.
//
// Tried and dropped in round 2: packed FFMA2 for the dot product and the outer product (30 fewer issue slots of 193 per blended
// visit, 2.013 -> 1.987 ms: within noise — shared-memory wavefronts, 80 % of the L1 peak, and latency bound this kernel, not
// issue slots alone); contiguous-piece and bulk-copy (one 160-B cp.async.bulk per record + mbarrier) staging instead of the
// two-lanes-per-record LDGSTS, whose 16-B pieces each cost a wavefront: 2.23 / 2.08 ms (profiles/r6p_*_ab.log).
// Also (commit 35841d2, profiles/r6c_bwd_transposed_ab.jsonl): the TRANSPOSED formulation — lane =
// list entry with its record in registers, the 32 pixels walked in a loop, channel and geometry gradients accumulated in
// registers (no exchange, 16 instead of 53 shared-memory wavefronts per visit), the per-pixel recurrences turned into two
// warp scans.  Bit-for-bit the same parity tier (34 / 34 tests), the same ~165 instructions per 32 (pixel, entry) pairs, but
// 2.81 vs 2.01 ms at C3 and 0.63 vs 0.39 ms at C4: eleven dependent shuffles per pixel iteration at 16 warps / SM (128
// registers) leave the schedulers idle — this kernel is bound by issue slots AND latency, not by shared memory alone.
#include "kernels.cuh"
#include "tile_stage.cuh"

namespace lsx {

namespace {

constexpr unsigned kFull = 0xffffffffu;

// Upstream gradient of every blended channel at one pixel (planar inputs), with the plane-depth gradient folded
// into the map channels (backward.cu:479-503); bg_dot = <background, dL/dcolor>.
template <int CT4>
__device__ __forceinline__ void load_pixel_gradients(const RenderParams& p, const bool inside, const size_t pix, const size_t HW,
                                                     const float pxf, const float pyf, float (&g)[CT4], float& bg_dot) {
#pragma unroll
    for (int c = 0; c < CT4; ++c) g[c] = 0.f;
    bg_dot = 0.f;
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

}

template <int CT4, int MB>
__global__ void __launch_bounds__(32, MB) render_bwd_kernel(const RenderParams p) {
    constexpr int RS = (REC_HEAD + CT4 + 7) & ~7;
    constexpr int GS = CT4 + 8;                // floats per packed gradient record
    constexpr int NPASS = (CT4 + 31) / 32;     // channel passes of the outer-product accumulation
    constexpr int TS = CT4 + 1;                // row stride of the one-time transposition scratch (odd -> conflict-free)
    // The 8 geometry terms (mean2D.xy, |mean2D|.xy, conic xyw, opacity) of an entry are exchanged through shared memory
    // like the weights.  With <= 24 channels the idle lanes CT4..31 sum one term each in the outer-product phase itself
    // (their "gradient view" is all ones; C4: 0.51 -> 0.44 ms).  Otherwise lane L sums term L >> 2 over pixels
    // 8 (L & 3) .. + 7 with two LDS.128 + 7 FADD and two shuffles, issued one entry late, finish the sum: ~24
    // instructions and one carried register instead of the ~43 and eight of a transposing shuffle butterfly
    // (C3: 2.56 -> 2.47 ms together with 16-entry rounds, which give the 2.6 KB of exchange space back to occupancy).
    // A mixed split (4 idle lanes at 27 channels + 4 terms elsewhere) costs more registers than it saves: 2.56 -> 2.66 ms.
    constexpr int NGL = (NPASS == 1 && 32 - CT4 >= 8) ? 8 : 0;
    constexpr bool kGeoSmem = (NGL == 0);
    constexpr int NX = 8;                      // geometry arrays exchanged through smem
    constexpr int XROW = 36;                   // floats between exchanged arrays: 16-B aligned, bank offset 4 per array
    using Stage = ListStage<RS>;
    constexpr int CHUNK = Stage::CHUNK;
    static_assert(Stage::kIdsOff >= 32 * TS * 4, "transposition scratch must fit in the record buffers");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(16) float s_w[2][(1 + NX) * XROW];  // weights + NX geometry terms, double buffered

    const int tile = blockIdx.x >> 3, warp = blockIdx.x & 7;
    const int tile_x = tile % p.grid_x, tile_y = tile / p.grid_x;
    const unsigned lane = threadIdx.x;
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
    // this pixel's contributors are the first last_k elements of its block's compacted list (render_fwd.cu)
    const int last_k = inside ? (int)p.k_contrib[pix] : 0;

    float g[CT4];  // upstream gradient of every blended channel of this pixel
    float bg_dot = 0.f;
    load_pixel_gradients<CT4>(p, inside, pix, HW, pxf, pyf, g, bg_dot);

    // deepest contributor of the block: list entries at or beyond it are never blended by any of its pixels
    const int cnt = (n > 0) ? (int)p.blk_cnt[8 * tile + warp] : 0;
    const int n_eff = min(__reduce_max_sync(kFull, last_k), cnt);
    if (n_eff == 0) return;

    // ---- transposed view: gT[ps][q] = upstream gradient of channel (32 ps + lane) at pixel q of this block ----
    float gT[NPASS][32];
    {
        float* ts = reinterpret_cast<float*>(smem_raw);  // aliases the record buffers (not yet live)
#pragma unroll
        for (int c = 0; c < CT4; ++c) ts[lane * TS + c] = g[c];
        __syncwarp();
#pragma unroll
        for (int ps = 0; ps < NPASS; ++ps) {
            const int c = ps * 32 + (int)lane;
#pragma unroll
            for (int q = 0; q < 32; ++q) gT[ps][q] = (c < CT4) ? ts[q * TS + c] : ((c < CT4 + NGL) ? 1.0f : 0.f);
        }
        __syncwarp();
    }
    // which exchanged array this lane sums in the outer-product phase: 0 = weights, 1 + k = geometry term k
    const uint32_t my_array = ((int)lane >= CT4 && (int)lane < CT4 + NGL) ? (1u + lane - CT4) : 0u;
    Stage stage;
    const int nrounds = (n_eff + CHUNK - 1) / CHUNK;

    const float Tf_bg = T_final * bg_dot;  // background term of dL/dalpha: -T_final / (1 - alpha) * <bg, dL/dcolor>
    float Bacc = 0.f;  // <colour accumulated behind the current entry (inclusive), upstream gradient>
    const uint32_t w_addr = smem_u32(&s_w[0][0]);
    unsigned parity = 0;
    float psum = 0.f;           // kGeoSmem: this lane's 8-pixel partial sum of term lane >> 2 of the last blended entry
    float* pgrec = p.grad_records;

    // back to front over the compacted list: round r covers elements n_eff-1 - (16 r + slot)
    stage.start(smem_raw, p.blk_list + (size_t)warp * p.list_stride + range.x, p.point_list + range.x, p.records,
                n_eff - 1, -1, n_eff, n, p.P);
    LSX_CHECK_INDEX(cnt, (long long)n + 1, "block list length");
    LSX_CHECK_INDEX((long long)range.y - 1, p.R, "tile range end");
    for (int r = 0; r < nrounds; ++r) {
        stage.advance(r);
        const int m = stage.round_size(r);
        uint32_t ra = stage.rec_addr(r & 1) - Stage::kRecBytes;
        uint32_t ia = stage.ids_addr(r & 1) - 4;
        const int k0 = n_eff - 1 - r * CHUNK;  // list element of slot 0

        for (int s_ = 0; s_ < m; ++s_) {
            const int e = k0 - s_;
            ra += Stage::kRecBytes;
            ia += 4;
            const float4 h0 = lds128(ra);
            const float2 h1 = lds64(ra + 16);
            const float dx = __fadd_rn(h0.x, -pxf), dy = __fadd_rn(h0.y, -pyf);
            const float power = splat_power(h0.z, h0.w, h1.x, dx, dy);
            const float G = expf(power);
            const float alpha = splat_alpha(h1.y, G);
            const bool blend = (e < last_k) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
            if (__ballot_sync(kFull, blend) == 0) continue;
            // Geometry terms of the PREVIOUS blended entry: the last two reduction levels are issued here, one entry
            // late, so that their latency overlaps this entry's rcp -> dot product chain.  No "pending" flag: before the
            // first blended entry psum is 0 and pgrec points at the first record, so the guarded atomic does nothing, and
            // entries that no lane blends never reach this point.
            if constexpr (kGeoSmem) {
                psum += __shfl_xor_sync(kFull, psum, 1);
                psum += __shfl_xor_sync(kFull, psum, 2);
                if ((lane & 3u) == 0u && psum != 0.f) atomicAdd(pgrec + CT4 + (lane >> 2), psum);
            }

            // per-lane terms; lanes that do not blend carry u = w = 0 so that every product below vanishes
            float w = 0.f, u = 0.f;  // u = G * dL/dalpha
            {
                // straight-line form: some lane blends (ballot above), so the warp issues these instructions anyway;
                // lanes that do not blend compute on finite values (alpha <= 0.99) and discard the results
                const float om = __fadd_rn(1.f, -alpha);
                float rcp;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(om));
                const float Tn = T * rcp;
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int q = 0; q < CT4 / 4; ++q) {
                    const float4 f = lds128(ra + REC_HEAD * 4 + q * 16);
                    s0 += f.x * g[4 * q + 0];
                    s1 += f.y * g[4 * q + 1];
                    s2 += f.z * g[4 * q + 2];
                    s3 += f.w * g[4 * q + 3];
                }
                const float s = (s0 + s1) + (s2 + s3);
                const float dL_dalpha = (s - Bacc) * Tn - Tf_bg * rcp;
                w = blend ? alpha * Tn : 0.f;
                u = blend ? G * dL_dalpha : 0.f;
                T = blend ? Tn : T;
                Bacc = blend ? alpha * s + om * Bacc : Bacc;
            }
            const float ku = h1.y * u;  // opacity * G * dL/dalpha = G * dL/dG
            const float kdx = ku * dx, kdy = ku * dy;
            // the constant factors (0.5 W, 0.5 H on the mean2D terms, -0.5 on the conic terms) are applied once per
            // Gaussian by preprocess_bwd_kernel instead of once per (block, entry) visit
            const float mx = -kdx * h0.z - kdy * h0.w;
            const float my = -kdy * h1.x - kdx * h0.w;
            const float gv[8] = {mx, my, fabsf(mx), fabsf(my), kdx * dx, kdx * dy, kdy * dy, u};

            // ---- channel gradients: lane c sums w[q] * gT[c][q] over the block's 32 pixels; lanes CT4.. sum the
            //      first NGL geometry terms the same way ----
            const uint32_t wa = w_addr + parity * (uint32_t)((1 + NX) * XROW * 4);
            parity ^= 1u;
            sts32(wa + lane * 4u, w);
#pragma unroll
            for (int k = 0; k < NX; ++k) sts32(wa + (uint32_t)((1 + k) * XROW * 4) + lane * 4u, gv[k]);
            __syncwarp();
            if constexpr (kGeoSmem) {
                float4 ga[2];
                lds128xN<2>(wa + (1u + (lane >> 2)) * (uint32_t)(XROW * 4) + (lane & 3u) * 32u, ga);
                psum = ((ga[0].x + ga[0].y) + (ga[0].z + ga[0].w)) + ((ga[1].x + ga[1].y) + (ga[1].z + ga[1].w));
            }
            const uint32_t wl = wa + my_array * (uint32_t)(XROW * 4);
            float cg[NPASS][4];
#pragma unroll
            for (int ps = 0; ps < NPASS; ++ps) cg[ps][0] = cg[ps][1] = cg[ps][2] = cg[ps][3] = 0.f;
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
                const float4 wq = lds128(wl + q4 * 16);
#pragma unroll
                for (int ps = 0; ps < NPASS; ++ps) {
                    cg[ps][0] += wq.x * gT[ps][4 * q4 + 0];
                    cg[ps][1] += wq.y * gT[ps][4 * q4 + 1];
                    cg[ps][2] += wq.z * gT[ps][4 * q4 + 2];
                    cg[ps][3] += wq.w * gT[ps][4 * q4 + 3];
                }
            }
            LSX_CHECK_INDEX(lds32i(ia), p.P, "gradient record");
            float* grec = p.grad_records + (size_t)lds32i(ia) * GS;
#pragma unroll
            for (int ps = 0; ps < NPASS; ++ps) {
                const int c = ps * 32 + (int)lane;
                const float sum = (cg[ps][0] + cg[ps][1]) + (cg[ps][2] + cg[ps][3]);
                // record layout: channel c at c, geometry term k at CT4 + k — i.e. at this lane's own index
                if ((c < p.n_channels || (c >= CT4 && c < CT4 + NGL)) && sum != 0.f) atomicAdd(grec + c, sum);
            }
            pgrec = grec;  // the geometry terms (pv) are reduced at the top of the next iteration
        }
    }
    stage.drain();
    if constexpr (kGeoSmem) {  // geometry terms of the last blended entry
        psum += __shfl_xor_sync(kFull, psum, 1);
        psum += __shfl_xor_sync(kFull, psum, 2);
        if ((lane & 3u) == 0u && psum != 0.f) atomicAdd(pgrec + CT4 + (lane >> 2), psum);
    }
}

// At 28 channels 20 CTAs / SM are requested: ptxas otherwise settles on more registers than the 96 that fit.
template <int CT4>
int launch_bwd_t(const RenderParams& p, cudaStream_t stream, bool debug) {
    constexpr int RS = (REC_HEAD + CT4 + 7) & ~7;
    // (16 channels: asking for 24 / 28 / 32 CTAs per SM — 80 / 72 / 64 registers — is 5 / 13 / 49 % slower than the 96 registers
    // ptxas picks on its own, profiles/r7d_bwd16_ab.log)
    constexpr int MB = (CT4 == 28) ? 20 : 0;
    const size_t smem = ListStage<RS>::kSmemBytes;
    const long long blocks = (long long)p.grid_x * p.grid_y * 8;
    render_bwd_kernel<CT4, MB><<<(unsigned)blocks, 32, smem, stream>>>(p);
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

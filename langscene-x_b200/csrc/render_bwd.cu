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
#include <cstdlib>

#include "kernels.cuh"
#include "tile_stage.cuh"

#ifndef LSX_BWD_DEFAULT_VARIANT
#define LSX_BWD_DEFAULT_VARIANT 0
#endif
#ifndef LSX_BWD_T_MIN_BLOCKS
#define LSX_BWD_T_MIN_BLOCKS 0
#endif

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
                n_eff - 1, -1, n_eff);
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

// ---------------------------------------------------------------------------------------------------------------------
// Transposed formulation (LSX_BWD_VARIANT=1): lane = list ENTRY, the block's 32 pixels are walked in a loop.
//
// A round is 32 consecutive elements of the block's compacted list, back to front (lane 0 = deepest).  Every lane keeps
// ITS entry's record (head + Ct channels) in registers for the whole round — fetched once, straight from global/L2,
// no shared-memory staging — and accumulates that entry's channel gradients and its 8 geometry terms over the
// pixels IN REGISTERS: no cross-lane reduction, no weight exchange, no per-visit barrier.  The upstream gradients of
// the 32 pixels sit in shared memory (3.5 KB at 28 channels) and are read as warp-broadcast LDS.128: the same Ct/4
// loads feed the dot product s = <feat, dL/dpix> AND the channel-gradient FMAs (the pixel-per-lane kernel above pays
// Ct/4 + 8 loads and an exchange of 9 values through shared memory per visit: ~53 shared-memory wavefronts per visit
// against ~16 per pixel iteration here).
// What was a per-pixel serial recurrence becomes two warp scans per pixel iteration:
//     T_before(e) = T_p * prod_{i <= lane, blended} 1 / (1 - alpha_i)                 (inclusive prefix product)
//     dL/dalpha_e = s_e T_before(e) - (B_p + sum_{i < lane} w_i s_i) / (1 - alpha_e)  (exclusive prefix sum)
// with the per-pixel carries (T_p, B_p = sum over deeper entries of w s, started at T_final <bg, dL/dcolor>) living
// in the registers of lane p.  The closed form equals the reference's "colour behind" recurrence (backward.cu:581-641):
// accumulated colour behind e = (sum_{d deeper} w_d c_d) / (T_before(e) (1 - alpha_e)).
// Pixels whose last contributor lies in front of the whole round are skipped (uniform branch), which the pixel-per-lane
// form cannot do.  At the end of a round the 32 lanes' results go through a padded shared-memory tile so that the global
// reductions stay coalesced: lane c issues RED.ADD.F32 for float c of one entry's contiguous gradient record at a time.
template <int CT4, int MB>
__global__ void __launch_bounds__(32, MB) render_bwd_t_kernel(const RenderParams p) {
    constexpr int RS = (REC_HEAD + CT4 + 7) & ~7;
    constexpr int GS = CT4 + 8;          // floats per packed gradient record
    constexpr int FROW = GS | 1;         // odd row stride of the flush tile: conflict-free writes (lane = row) and reads (lane = column)
    __shared__ __align__(16) float s_g[32 * CT4];   // upstream gradient of every blended channel at the block's 32 pixels
    __shared__ float s_f[32 * FROW];                 // flush tile: [entry][channel gradients | geometry terms]
    __shared__ int s_id[32];

    const int tile = blockIdx.x >> 3, warp = blockIdx.x & 7;
    const int tile_x = tile % p.grid_x, tile_y = tile / p.grid_x;
    const unsigned lane = threadIdx.x;
    const int bx = tile_x * TILE_X + (warp & 1) * 8, by = tile_y * TILE_Y + (warp >> 1) * 4;
    const int px = bx + (int)(lane & 7), py = by + (int)(lane >> 3);
    const bool inside = px < p.W && py < p.H;
    const size_t HW = (size_t)p.H * p.W;
    const size_t pix = (size_t)py * p.W + px;

    const uint2 range = p.ranges[tile];
    const int n = (int)(range.y - range.x);

    // ---- state of pixel `lane` ----------------------------------------------------------------------
    float T = inside ? p.final_T[pix] : 0.f;
    const int last_k = inside ? (int)p.k_contrib[pix] : 0;
    const int cnt = (n > 0) ? (int)p.blk_cnt[8 * tile + warp] : 0;
    const int n_eff = min(__reduce_max_sync(kFull, last_k), cnt);
    if (n_eff == 0) return;
    float Bs;  // sum over deeper blended entries of w * s, plus the background term T_final * <bg, dL/dcolor>
    {
        float g[CT4];
        float bg_dot = 0.f;
        load_pixel_gradients<CT4>(p, inside, pix, HW, (float)px, (float)py, g, bg_dot);
        Bs = T * bg_dot;
#pragma unroll
        for (int q = 0; q < CT4 / 4; ++q)
            *reinterpret_cast<float4*>(&s_g[lane * CT4 + 4 * q]) = make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]);
    }
    __syncwarp();
    const uint32_t g_addr = smem_u32(s_g), f_addr = smem_u32(s_f), id_addr = smem_u32(s_id);

    const uint32_t* list = p.blk_list + (size_t)warp * p.list_stride + range.x;
    const uint32_t* plist = p.point_list + range.x;
    const int nrounds = (n_eff + 31) >> 5;
    int k = n_eff - 1 - (int)lane;  // this lane's list element in the current round (back to front)
    uint32_t id = (k >= 0) ? __ldg(plist + __ldg(list + k)) : 0u;

    for (int r = 0; r < nrounds; ++r) {
        const bool valid = k >= 0;
        const int kfirst = max(n_eff - 32 - 32 * r, 0);  // smallest list element of this round
        // ---- this lane's entry: record -> registers; next round's id in flight ----
        float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 h1 = make_float2(0.f, 0.f);
        float f[CT4];
#pragma unroll
        for (int c = 0; c < CT4; ++c) f[c] = 0.f;
        if (valid) {
            const float* rec = p.records + (size_t)id * RS;
            h0 = __ldg(reinterpret_cast<const float4*>(rec));
            h1 = __ldg(reinterpret_cast<const float2*>(rec + 4));
#pragma unroll
            for (int q = 0; q < CT4 / 4; ++q) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(rec + REC_HEAD + 4 * q));
                f[4 * q] = v.x; f[4 * q + 1] = v.y; f[4 * q + 2] = v.z; f[4 * q + 3] = v.w;
            }
        }
        const int kn = k - 32;
        const uint32_t id_next = (kn >= 0) ? __ldg(plist + __ldg(list + kn)) : 0u;

        float df[CT4], ge[8];
#pragma unroll
        for (int c = 0; c < CT4; ++c) df[c] = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) ge[c] = 0.f;

#pragma unroll 1
        for (int q_ = 0; q_ < 32; ++q_) {
            const int lk = __shfl_sync(kFull, last_k, q_);
            if (lk <= kfirst) continue;  // every entry of this round lies behind this pixel's last contributor
            const float pxf = (float)(bx + (q_ & 7)), pyf = (float)(by + (q_ >> 3));
            const float dx = __fadd_rn(h0.x, -pxf), dy = __fadd_rn(h0.y, -pyf);
            const float power = splat_power(h0.z, h0.w, h1.x, dx, dy);
            const float G = expf(power);
            const float alpha = splat_alpha(h1.y, G);
            const bool blend = valid && (k < lk) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
            if (__ballot_sync(kFull, blend) == 0) continue;
            const float om = blend ? __fadd_rn(1.f, -alpha) : 1.f;
            float rc;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(om));
            // inclusive prefix product of 1 / (1 - alpha) over the lanes (= entries, deepest first)
            float pr = rc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float t = __shfl_up_sync(kFull, pr, d);
                pr = (lane >= (unsigned)d) ? pr * t : pr;
            }
            const float Tp = __shfl_sync(kFull, T, q_);
            const float Tn = Tp * pr;  // transmittance in front of this entry
            const float w = blend ? alpha * Tn : 0.f;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            const uint32_t ga = g_addr + (uint32_t)(q_ * CT4 * 4);
#pragma unroll
            for (int q = 0; q < CT4 / 4; ++q) {
                const float4 gq = lds128(ga + q * 16);
                s0 += f[4 * q + 0] * gq.x;
                s1 += f[4 * q + 1] * gq.y;
                s2 += f[4 * q + 2] * gq.z;
                s3 += f[4 * q + 3] * gq.w;
                df[4 * q + 0] += w * gq.x;
                df[4 * q + 1] += w * gq.y;
                df[4 * q + 2] += w * gq.z;
                df[4 * q + 3] += w * gq.w;
            }
            const float s = (s0 + s1) + (s2 + s3);
            // prefix sum of w * s over the lanes: inclusive in ps, exclusive (deeper entries of this round) in ex
            float ps = w * s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float t = __shfl_up_sync(kFull, ps, d);
                ps = (lane >= (unsigned)d) ? ps + t : ps;
            }
            float ex = __shfl_up_sync(kFull, ps, 1);
            ex = (lane == 0u) ? 0.f : ex;
            const float Bp = __shfl_sync(kFull, Bs, q_);
            const float dL_dalpha = s * Tn - (Bp + ex) * rc;
            const float u = blend ? G * dL_dalpha : 0.f;  // G * dL/dalpha
            const float ku = h1.y * u;                   // opacity * G * dL/dalpha = G * dL/dG
            const float kdx = ku * dx, kdy = ku * dy;
            const float mx = -kdx * h0.z - kdy * h0.w;
            const float my = -kdy * h1.x - kdx * h0.w;
            ge[0] += mx;
            ge[1] += my;
            ge[2] += fabsf(mx);
            ge[3] += fabsf(my);
            ge[4] += kdx * dx;
            ge[5] += kdx * dy;
            ge[6] += kdy * dy;
            ge[7] += u;
            // carries of pixel q_ past this round
            const float prl = __shfl_sync(kFull, pr, 31), psl = __shfl_sync(kFull, ps, 31);
            if (lane == (unsigned)q_) {
                T = Tp * prl;
                Bs = Bp + psl;
            }
        }

        // ---- flush: lane = entry  ->  tile  ->  lane = float of one entry's contiguous gradient record ----
#pragma unroll
        for (int c = 0; c < CT4; ++c) sts32(f_addr + (uint32_t)((lane * FROW + c) * 4), df[c]);
#pragma unroll
        for (int c = 0; c < 8; ++c) sts32(f_addr + (uint32_t)((lane * FROW + CT4 + c) * 4), ge[c]);
        sts32i(id_addr + lane * 4u, (int)id);
        __syncwarp();
        const int m = min(32, n_eff - 32 * r);
#pragma unroll 1
        for (int e = 0; e < m; ++e) {
            float* grec = p.grad_records + (size_t)lds32i(id_addr + (uint32_t)(e * 4)) * GS;
#pragma unroll
            for (int c0 = 0; c0 < GS; c0 += 32) {
                const int c = c0 + (int)lane;
                if (c < GS) {
                    float v;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(f_addr + (uint32_t)((e * FROW + c) * 4)) : "memory");
                    if (v != 0.f) atomicAdd(grec + c, v);
                }
            }
        }
        __syncwarp();
        id = id_next;
        k = kn;
    }
}

// At 28 channels 20 CTAs / SM are requested: ptxas otherwise settles on more registers than the 96 that fit.
// LSX_BWD_VARIANT: 0 = pixel-per-lane kernel, 1 = entry-per-lane (transposed) kernel.  Read once.
int bwd_variant() {
    static const int v = [] {
        const char* e = getenv("LSX_BWD_VARIANT");
        return e ? atoi(e) : LSX_BWD_DEFAULT_VARIANT;
    }();
    return v;
}

template <int CT4>
int launch_bwd_t(const RenderParams& p, cudaStream_t stream, bool debug) {
    constexpr int RS = (REC_HEAD + CT4 + 7) & ~7;
    const long long blocks = (long long)p.grid_x * p.grid_y * 8;
    if (bwd_variant() == 1) {
        render_bwd_t_kernel<CT4, LSX_BWD_T_MIN_BLOCKS><<<(unsigned)blocks, 32, 0, stream>>>(p);
    } else {
        constexpr int MB = (CT4 == 28) ? 20 : 0;
        const size_t smem = ListStage<RS>::kSmemBytes;
        render_bwd_kernel<CT4, MB><<<(unsigned)blocks, 32, smem, stream>>>(p);
    }
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

// preprocess.cu — per-Gaussian stage of the rasterizer (forward K1, backward K8+K9 fused, markVisible).
//
// Reference behaviour restated here (paths under field_construction/submodules/diff-langsurf-rasterizer/):
//   forward : cuda_rasterizer/forward.cu:21-268  (SH->RGB, cov3D, EWA cov2D, conic, radius, tile rect)
//             cuda_rasterizer/auxiliary.h:41-56,139-164 (ndc2Pix in double, getRect, near cull z<=0.2)
//   backward: cuda_rasterizer/backward.cu:20-139 (SH), :144-274 (cov2D), :278-341 (cov3D), :346-396
//
// B200 design notes
//   * one thread per Gaussian and EVERY tensor row is read and written by its owner with 16-B accesses where its alignment
//     allows: the rows are contiguous per Gaussian (192 B of SH, 64 B of language feature, a 160-B record), so the owner's
//     accesses use whole sectors and consecutive lanes walk consecutive rows.  (Rounds 1-2 moved the block's slabs through
//     shared memory with coalesced copies: the cooperative index arithmetic was three quarters of the kernel's instructions
//     and the 47 KB tile held occupancy at 25 %; per-owner rows are 29 % / 11 % faster, profiles/r6r_*.)  Culled splats read
//     nothing but their position / covariance inputs;
//   * the forward kernel is a single streaming pass (HBM-bound: ~332 B in / ~250 B out per visible Gaussian at the
//     headline config) that ALSO
//       - zero-fills out_observe (no separate memset),
//       - emits the 32-bit depth sort key (0xFFFFFFFF for culled splats) for the depth-major presort,
//       - packs every attribute the tile renderers need into ONE sector-aligned record per Gaussian
//         ({xy, conic, opacity} + all blended channels), so that the per-tile staging is a single
//         contiguous copy per list entry instead of 5 scattered gathers.
//   * the backward kernel fuses the reference's computeCov2DCUDA + preprocessCUDA launches and writes
//     every gradient row exactly once (zeros for culled splats), which removes ~400 MB of separate
//     zero-fill traffic at 1M Gaussians.
//   * arithmetic that feeds radii / tile rects / depth keys keeps the reference's scalar expression
//     order (including the fp64 island in ndc_to_pix) because those outputs are compared bit-exactly.
#include "head_math.cuh"
#include "kernels.cuh"

namespace lsx {

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------

// world covariance from (scale, un-normalised quaternion); forward.cu:119-152
__device__ __forceinline__ void world_covariance(const float3 scale, const float mod, const float4 q, float* cov6) {
    Mat3 S;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) S.m[c][r] = 0.0f;
    S.m[0][0] = mod * scale.x;
    S.m[1][1] = mod * scale.y;
    S.m[2][2] = mod * scale.z;

    const float r = q.x, x = q.y, y = q.z, z = q.w;
    Mat3 R;
    R.m[0][0] = 1.f - 2.f * (y * y + z * z);
    R.m[0][1] = 2.f * (x * y - r * z);
    R.m[0][2] = 2.f * (x * z + r * y);
    R.m[1][0] = 2.f * (x * y + r * z);
    R.m[1][1] = 1.f - 2.f * (x * x + z * z);
    R.m[1][2] = 2.f * (y * z - r * x);
    R.m[2][0] = 2.f * (x * z - r * y);
    R.m[2][1] = 2.f * (y * z + r * x);
    R.m[2][2] = 1.f - 2.f * (x * x + y * y);

    const Mat3 Mx = mat3_mul(S, R);
    const Mat3 Sigma = mat3_mul(mat3_transpose(Mx), Mx);
    cov6[0] = Sigma.m[0][0];
    cov6[1] = Sigma.m[0][1];
    cov6[2] = Sigma.m[0][2];
    cov6[3] = Sigma.m[1][1];
    cov6[4] = Sigma.m[1][2];
    cov6[5] = Sigma.m[2][2];
}

struct Cov2DFrame {  // intermediates shared by forward and backward EWA projection
    Mat3 T;          // W * J
    Mat3 V;          // symmetric world covariance
    float3 t;        // clamped view-space mean
    float txtz, tytz, limx, limy;
};

__device__ __forceinline__ void ewa_frame(const float3 mean, const float fx, const float fy, const float tan_fovx,
                                          const float tan_fovy, const float* cov6, const float* __restrict__ view,
                                          Cov2DFrame& f) {
    float3 t = xform_point_4x3(mean, view);
    f.limx = 1.3f * tan_fovx;
    f.limy = 1.3f * tan_fovy;
    f.txtz = t.x / t.z;
    f.tytz = t.y / t.z;
    t.x = fminf(f.limx, fmaxf(-f.limx, f.txtz)) * t.z;
    t.y = fminf(f.limy, fmaxf(-f.limy, f.tytz)) * t.z;
    f.t = t;

    Mat3 J;
    J.m[0][0] = fx / t.z;
    J.m[0][1] = 0.0f;
    J.m[0][2] = -(fx * t.x) / (t.z * t.z);
    J.m[1][0] = 0.0f;
    J.m[1][1] = fy / t.z;
    J.m[1][2] = -(fy * t.y) / (t.z * t.z);
    J.m[2][0] = 0.0f;
    J.m[2][1] = 0.0f;
    J.m[2][2] = 0.0f;

    Mat3 Wm;
    Wm.m[0][0] = view[0]; Wm.m[0][1] = view[4]; Wm.m[0][2] = view[8];
    Wm.m[1][0] = view[1]; Wm.m[1][1] = view[5]; Wm.m[1][2] = view[9];
    Wm.m[2][0] = view[2]; Wm.m[2][1] = view[6]; Wm.m[2][2] = view[10];

    f.T = mat3_mul(Wm, J);

    f.V.m[0][0] = cov6[0]; f.V.m[0][1] = cov6[1]; f.V.m[0][2] = cov6[2];
    f.V.m[1][0] = cov6[1]; f.V.m[1][1] = cov6[3]; f.V.m[1][2] = cov6[4];
    f.V.m[2][0] = cov6[2]; f.V.m[2][1] = cov6[4]; f.V.m[2][2] = cov6[5];
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
#ifndef LSX_PRE_FWD_THREADS
#define LSX_PRE_FWD_THREADS 128
#endif
#ifndef LSX_PRE_BWD_THREADS
#define LSX_PRE_BWD_THREADS 128
#endif
constexpr int kPreFwdThreads = LSX_PRE_FWD_THREADS;
// smem stride of the threads' record rows: 16-B aligned rows, an odd number of 16-B units => the owners' float4 accesses
// are conflict free
__host__ __device__ static inline int record_tile_row(int rec_stride) { return rec_stride + 4; }

// One Gaussian per thread.  `rec`: the thread's own record row in shared memory (head, colour and pad columns are written here,
// the caller fills the feature / map columns of visible splats).
// RAW = the fused render-wrapper mode: `means3D / scales / rotations / opacities` are the reference's raw
// nn.Parameters (positions, log-scales, un-normalised quaternions, opacity logits), `all_map` is absent, and the wrapper's
// per-Gaussian work — optional camera-pose transform, exp / normalize / sigmoid, plane normal, all_map (head_math.cuh) — is
// done here instead of in separate passes over P.  It is a separate instantiation so that the code generated for the
// reference-shaped mode (whose radii / rects / keys are a bit-exact contract) is untouched.
// `sh_row` is the Gaussian's row of the (P, M, 3) tensor in global memory, fetched with 16-B loads right before the polynomial
// (only for splats that survive the culling tests).  Returns whether the splat is visible.
template <bool RAW>
__device__ __forceinline__ bool preprocess_fwd_row(const PreprocessFwdParams& p, const int idx, const float* __restrict__ sh_row,
                                                   float* __restrict__ rec) {
    // defaults for a culled splat (its record is never read; keep it finite)
    reinterpret_cast<float4*>(rec)[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    reinterpret_cast<float4*>(rec)[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    rec[REC_HEAD + 0] = rec[REC_HEAD + 1] = rec[REC_HEAD + 2] = 0.f;
    for (int c = p.n_channels; c < p.rec_stride - REC_HEAD; ++c) rec[REC_HEAD + c] = 0.f;
    if constexpr (RAW) {
        if (p.render_geo)  // the map columns are produced below; a culled splat keeps zeros
            for (int c = 0; c < 5; ++c) rec[REC_HEAD + 3 + (p.include_feature ? p.F + p.Fi : 0) + c] = 0.f;
    }
    p.radii[idx] = 0;
    p.tiles_touched[idx] = 0;
    p.out_observe[idx] = 0;
    p.depth_keys[idx] = 0xFFFFFFFFu;

    float3 pw = make_float3(p.means3D[3 * idx], p.means3D[3 * idx + 1], p.means3D[3 * idx + 2]);
    float4 q_in = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (RAW) {
        q_in = reinterpret_cast<const float4*>(p.rotations)[idx];
        if (p.pose != nullptr) {  // render(..., camera_pose=pose): means3D = R xyz + T, rotations = pose_q (x) rotation
            const Rot3 Rp = rotation_of(p.pose);
            float xw[3];
            pose_apply_point(Rp, p.pose, pw.x, pw.y, pw.z, xw);
            pw = make_float3(xw[0], xw[1], xw[2]);
            q_in = pose_apply_quat(p.pose, q_in);
        }
    }
    const float3 pv = xform_point_4x3(pw, p.view);
    if (pv.z <= 0.2f) {
        if (p.prefiltered) {
            printf("Point is filtered although prefiltered is set. This shouldn't happen!");
            __trap();
        }
        return false;
    }

    const float4 ph = xform_point_4x4(pw, p.proj);
    const float pw_inv = 1.0f / (ph.w + 0.0000001f);
    const float3 pn = make_float3(ph.x * pw_inv, ph.y * pw_inv, ph.z * pw_inv);

    float cov6[6];
    float raw_opacity = 0.f;  // RAW: sigmoid of the logit
    if constexpr (RAW) {
        const HeadCam hc = load_cam(p.view, p.campos);
        const float xw[3] = {pw.x, pw.y, pw.z};
        const float sraw[3] = {p.scales[3 * idx], p.scales[3 * idx + 1], p.scales[3 * idx + 2]};
        const HeadFrame hf = head_frame(hc, xw, sraw, q_in, p.opacities[idx]);
        raw_opacity = hf.sig;
        if (p.render_geo) {  // all_map = [local plane normal, 1, |<normal, camera-space position>|]
            float* am = rec + REC_HEAD + 3 + (p.include_feature ? p.F + p.Fi : 0);
            am[0] = hf.nl[0]; am[1] = hf.nl[1]; am[2] = hf.nl[2]; am[3] = 1.0f; am[4] = fabsf(hf.d);
        }
        world_covariance(make_float3(hf.sc[0], hf.sc[1], hf.sc[2]), p.scale_modifier, make_float4(hf.q[0], hf.q[1], hf.q[2], hf.q[3]),
                         cov6);
#pragma unroll
        for (int i = 0; i < 6; ++i) p.cov3D[6 * idx + i] = cov6[i];
    } else if (p.cov3D_precomp != nullptr) {
#pragma unroll
        for (int i = 0; i < 6; ++i) cov6[i] = p.cov3D_precomp[6 * idx + i];
    } else {
        const float3 sc = make_float3(p.scales[3 * idx], p.scales[3 * idx + 1], p.scales[3 * idx + 2]);
        const float4 q = reinterpret_cast<const float4*>(p.rotations)[idx];
        world_covariance(sc, p.scale_modifier, q, cov6);
#pragma unroll
        for (int i = 0; i < 6; ++i) p.cov3D[6 * idx + i] = cov6[i];
    }

    Cov2DFrame fr;
    ewa_frame(pw, p.focal_x, p.focal_y, p.tan_fovx, p.tan_fovy, cov6, p.view, fr);
    const Mat3 c2 = mat3_mul(mat3_mul(mat3_transpose(fr.T), mat3_transpose(fr.V)), fr.T);
    const float cxx = c2.m[0][0] + 0.3f;
    const float cxy = c2.m[0][1];
    const float cyy = c2.m[1][1] + 0.3f;

    const float det = (cxx * cyy - cxy * cxy);
    if (det == 0.0f) return false;
    const float det_inv = 1.f / det;
    const float3 conic = make_float3(cyy * det_inv, -cxy * det_inv, cxx * det_inv);

    const float mid = 0.5f * (cxx + cyy);
    const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
    const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
    const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
    const float2 pix = make_float2(ndc_to_pix(pn.x, p.W), ndc_to_pix(pn.y, p.H));
    uint2 rmin, rmax;
    tile_rect(pix, (int)my_radius, rmin, rmax, p.grid_x, p.grid_y);
    const uint32_t ntiles = (rmax.x - rmin.x) * (rmax.y - rmin.y);
    if (ntiles == 0) return false;

    // colour: SH evaluation (degrees 0..3) or the caller's precomputed RGB
    float rgb[3];
    if (p.colors_precomp == nullptr) {
        const float dx0 = pw.x - p.campos[0], dy0 = pw.y - p.campos[1], dz0 = pw.z - p.campos[2];
        // Operation order pinned with explicit intrinsics to the sequence nvcc 12.9 emits for the reference's
        // computeColorFromSH on sm_100a (which products are fused into FMAs is otherwise the compiler's
        // choice), so the packed colour is reproduced bit for bit.
        const float len = sqrtf(__fmaf_rn(dz0, dz0, __fmaf_rn(dx0, dx0, __fmul_rn(dy0, dy0))));
        const float x = dx0 / len, y = dy0 / len, z = dz0 / len;
        float shv[48];  // the coefficients in registers
        {
            const int n_sh = p.M * 3;
#pragma unroll
            for (int k = 0; k < 48; ++k) shv[k] = 0.f;
            if ((n_sh & 3) == 0 && (reinterpret_cast<uintptr_t>(sh_row) & 15u) == 0) {
                const float4* s4 = reinterpret_cast<const float4*>(sh_row);
#pragma unroll
                for (int j = 0; j < 12; ++j)
                    if (4 * j < n_sh) {
                        const float4 v = __ldg(s4 + j);
                        shv[4 * j] = v.x;
                        shv[4 * j + 1] = v.y;
                        shv[4 * j + 2] = v.z;
                        shv[4 * j + 3] = v.w;
                    }
            } else {
#pragma unroll
                for (int k = 0; k < 48; ++k)
                    if (k < n_sh) shv[k] = __ldg(sh_row + k);
            }
        }
        const float* sh = shv;
        unsigned clamp_bits = 0;
        // direction polynomials shared by the three colour channels
        float b1y = 0.f, b1z = 0.f, b1x = 0.f, b4 = 0.f, b5 = 0.f, b6 = 0.f, b7 = 0.f, b8 = 0.f;
        float b9 = 0.f, b10 = 0.f, b11 = 0.f, b12 = 0.f, b13 = 0.f, b14 = 0.f, b15 = 0.f;
        if (p.D > 0) {
            b1y = __fmul_rn(y, kSH1);
            b1z = __fmul_rn(z, kSH1);
            b1x = __fmul_rn(x, kSH1);
            if (p.D > 1) {
                const float xy = __fmul_rn(y, x), yz = __fmul_rn(z, y), xz = __fmul_rn(z, x);
                const float zz = __fmul_rn(z, z), xx = __fmul_rn(x, x), yy = __fmul_rn(y, y);
                const float zz2 = __fadd_rn(zz, zz);
                const float xx_yy = __fadd_rn(xx, -yy);
                b4 = __fmul_rn(xy, kSH2[0]);
                b5 = __fmul_rn(yz, kSH2[1]);
                b6 = __fmul_rn(__fadd_rn(-yy, __fadd_rn(-xx, zz2)), kSH2[2]);
                b7 = __fmul_rn(xz, kSH2[3]);
                b8 = __fmul_rn(xx_yy, kSH2[4]);
                if (p.D > 2) {
                    const float q4 = __fadd_rn(-yy, __fmaf_rn(zz, 4.0f, -xx));  // 4zz - xx - yy
                    b9 = __fmul_rn(__fmul_rn(y, kSH3[0]), __fmaf_rn(xx, 3.0f, -yy));
                    b10 = __fmul_rn(__fmul_rn(xy, kSH3[1]), z);
                    b11 = __fmul_rn(__fmul_rn(y, kSH3[2]), q4);
                    b12 = __fmul_rn(__fmul_rn(z, kSH3[3]), __fmaf_rn(yy, -3.0f, __fmaf_rn(xx, -3.0f, zz2)));
                    b13 = __fmul_rn(q4, __fmul_rn(x, kSH3[4]));
                    b14 = __fmul_rn(xx_yy, __fmul_rn(z, kSH3[5]));
                    b15 = __fmul_rn(__fmul_rn(x, kSH3[6]), __fmaf_rn(yy, -3.0f, xx));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float res = __fmul_rn(sh[c], kSH0);
            if (p.D > 0) {
                res = __fmaf_rn(-b1y, sh[3 + c], res);
                res = __fmaf_rn(b1z, sh[6 + c], res);
                res = __fmaf_rn(-b1x, sh[9 + c], res);
                if (p.D > 1) {
                    res = __fmaf_rn(b4, sh[12 + c], res);
                    res = __fmaf_rn(b5, sh[15 + c], res);
                    res = __fmaf_rn(b6, sh[18 + c], res);
                    res = __fmaf_rn(b7, sh[21 + c], res);
                    res = __fmaf_rn(b8, sh[24 + c], res);
                    if (p.D > 2) {
                        res = __fmaf_rn(b9, sh[27 + c], res);
                        res = __fmaf_rn(b10, sh[30 + c], res);
                        res = __fmaf_rn(b11, sh[33 + c], res);
                        res = __fmaf_rn(b12, sh[36 + c], res);
                        res = __fmaf_rn(b13, sh[39 + c], res);
                        res = __fmaf_rn(b14, sh[42 + c], res);
                        res = __fmaf_rn(b15, sh[45 + c], res);
                    }
                }
            }
            res = __fadd_rn(res, 0.5f);
            if (res < 0) clamp_bits |= (1u << c);
            rgb[c] = fmaxf(res, 0.0f);
        }
        p.clamped[idx] = (uint8_t)clamp_bits;
#pragma unroll
        for (int c = 0; c < 3; ++c) p.rgb[3 * idx + c] = rgb[c];
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) rgb[c] = p.colors_precomp[3 * idx + c];
    }

    const float opacity = RAW ? raw_opacity : p.opacities[idx];
    p.depths[idx] = pv.z;
    p.depth_keys[idx] = __float_as_uint(pv.z);
    p.radii[idx] = (int)my_radius;
    p.means2D[idx] = pix;
    p.conic_opacity[idx] = make_float4(conic.x, conic.y, conic.z, opacity);
    p.tiles_touched[idx] = ntiles;

    // packed blend record (sector aligned): head + colour; the other channels are already in the tile
    reinterpret_cast<float4*>(rec)[0] = make_float4(pix.x, pix.y, conic.x, conic.y);
    reinterpret_cast<float4*>(rec)[1] = make_float4(conic.z, opacity, 0.f, 0.f);
    rec[REC_HEAD + 0] = rgb[0];
    rec[REC_HEAD + 1] = rgb[1];
    rec[REC_HEAD + 2] = rgb[2];
    return true;
}

// own-row copy global -> the thread's record row in shared memory
__device__ __forceinline__ void row_fetch(float* __restrict__ dst, const float* __restrict__ src, const int n) {
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
#pragma unroll 4
        for (int k = 0; k < n; k += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + k));
            dst[k] = v.x;
            dst[k + 1] = v.y;
            dst[k + 2] = v.z;
            dst[k + 3] = v.w;
        }
    } else {
#pragma unroll 5
        for (int k = 0; k < n; ++k) dst[k] = __ldg(src + k);
    }
}

// One Gaussian per thread, every tensor row read and written by its owner: the rows of the (P, ...) tensors are contiguous
// per Gaussian (192 B of SH, 64 B of language feature, a 160-B record), so the owner's 16-B accesses use whole sectors and
// consecutive lanes walk consecutive rows.  Shared memory only holds the thread's own record row while it is assembled (its
// column offsets depend on F / Fi at run time): no barrier, no cooperative index arithmetic — that arithmetic was three
// quarters of the staged version's instructions.  Culled splats read nothing but their position / covariance inputs.
// 6 blocks of 128 threads per SM: 80 registers (ptxas settles on 86 otherwise); C5 0.621 -> 0.595 ms
#ifndef LSX_PRE_FWD_MINB
#define LSX_PRE_FWD_MINB 6
#endif
template <bool RAW>
__global__ void __launch_bounds__(kPreFwdThreads, LSX_PRE_FWD_MINB) preprocess_fwd_kernel(const PreprocessFwdParams p) {
    extern __shared__ __align__(16) float s_pre[];
    const int idx = blockIdx.x * kPreFwdThreads + threadIdx.x;
    if (idx >= p.P) return;
    const int rec_row = record_tile_row(p.rec_stride);
    float* rec = s_pre + threadIdx.x * rec_row;
    const int n_sh = p.M * 3;
    const bool visible = preprocess_fwd_row<RAW>(p, idx, p.shs ? p.shs + (size_t)idx * n_sh : nullptr, rec);
    float4* out = reinterpret_cast<float4*>(p.records + (size_t)idx * p.rec_stride);
    const int n4 = p.rec_stride >> 2;
    // a culled splat has no list entry (tiles_touched = 0), so nothing ever reads its record: it is not written
    if (!visible) return;
    int c = REC_HEAD + 3;
    if (p.include_feature) {
        row_fetch(rec + c, p.language_feature + (size_t)idx * p.F, p.F);
        c += p.F;
        row_fetch(rec + c, p.language_feature_instance + (size_t)idx * p.Fi, p.Fi);
        c += p.Fi;
    }
    if (!RAW && p.render_geo) row_fetch(rec + c, p.all_map + (size_t)idx * 5, 5);
    for (int j = 0; j < n4; ++j) out[j] = reinterpret_cast<const float4*>(rec)[j];
}

// ------------------------------------------------------------------------------------------------
// backward (K8 + K9 fused): dL/dconic, dL/dmean2D, dL/dcolor  ->  dL/d{mean3D, cov3D, sh, scale, rot}
// ------------------------------------------------------------------------------------------------
// One Gaussian per thread.  `sh_row` is this Gaussian's row of the (P, M, 3) tensor in global memory (read with 16-B loads, by
// visible splats only; dL/dsh goes straight from registers to the row's place in p.dL_dsh); `grec` is its packed gradient
// record from the tile backward pass (render_bwd.cu):
// [rgb(3) | language(F) | instance(Fi) | all_map(5) | pad | mean2D.xy, |mean2D|.xy, conic.xyw, opacity].
// Culled splats have an all-zero record (memset, never accumulated into).
constexpr int kPreBwdThreads = LSX_PRE_BWD_THREADS;

// RAW (see preprocess_fwd_row): the inputs are the raw parameters; the gradients of the activated scale / rotation / opacity /
// position and of the all_map row are NOT written but pushed on through the wrapper's backward (head_math.cuh: activations,
// plane normal, all_map, optional camera pose) and arrive in dL_dmeans3D / dL_dscales / dL_drotations / dL_dopacity as
// gradients of the RAW parameters; `pose_acc` collects this row's 16 pose-gradient terms.
template <bool RAW>
__device__ __forceinline__ void preprocess_bwd_row(const PreprocessBwdParams& p, const int idx, const float* __restrict__ sh_row,
                                                   const float (&grec)[3], const float (&ggeo)[8], const float (&gmap)[5],
                                                   float (&pose_acc)[kPoseTerms]) {
    const int n_sh = p.M * 3;
    float* dsh_out = (p.dL_dsh && p.shs) ? p.dL_dsh + (size_t)idx * n_sh : nullptr;
    const bool dsh_vec = (n_sh & 3) == 0 && (reinterpret_cast<uintptr_t>(dsh_out) & 15u) == 0;
    p.dL_dmean2D[3 * idx + 0] = ggeo[0];
    p.dL_dmean2D[3 * idx + 1] = ggeo[1];
    p.dL_dmean2D[3 * idx + 2] = 0.f;
    p.dL_dmean2D_abs[3 * idx + 0] = ggeo[2];
    p.dL_dmean2D_abs[3 * idx + 1] = ggeo[3];
    p.dL_dmean2D_abs[3 * idx + 2] = 0.f;
    reinterpret_cast<float4*>(p.dL_dconic)[idx] = make_float4(ggeo[4], ggeo[5], 0.f, ggeo[6]);
    // p.accumulate: bit mask of the outputs that are ADDED to (multi-view accumulation) instead of written
    const unsigned am = (unsigned)p.accumulate;
    auto put = [am](float* dst, const float v, const unsigned bit) { *dst = (am & bit) ? *dst + v : v; };
    if constexpr (!RAW) put(p.dL_dopacity + idx, ggeo[7], LSX_ACC_OPACITY);

    if (!(p.radii[idx] > 0)) {
        if constexpr (RAW) {
            if (!(am & LSX_ACC_OPACITY)) p.dL_dopacity[idx] = 0.f;
        }
        // culled splat: every gradient row is zero (the reference relies on torch::zeros for this)
        if (!(am & LSX_ACC_MEANS3D)) {
#pragma unroll
            for (int i = 0; i < 3; ++i) p.dL_dmeans3D[3 * idx + i] = 0.f;
        }
        if (!(am & LSX_ACC_COV3D) && p.dL_dcov3D != nullptr) {
#pragma unroll
            for (int i = 0; i < 6; ++i) p.dL_dcov3D[6 * idx + i] = 0.f;
        }
        if (!(am & LSX_ACC_SCALES)) {
#pragma unroll
            for (int i = 0; i < 3; ++i) p.dL_dscales[3 * idx + i] = 0.f;
        }
        if (!(am & LSX_ACC_ROTATIONS)) reinterpret_cast<float4*>(p.dL_drotations)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dsh_out && !(p.accumulate & LSX_ACC_SH)) {
            if (dsh_vec) {
                for (int i = 0; i < n_sh; i += 4) *reinterpret_cast<float4*>(dsh_out + i) = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                for (int i = 0; i < n_sh; ++i) dsh_out[i] = 0.f;
            }
        }
        return;
    }

    float3 mean = make_float3(p.means3D[3 * idx], p.means3D[3 * idx + 1], p.means3D[3 * idx + 2]);
    // RAW: redo the wrapper's forward for this row (pose transform, activations, plane normal)
    const float3 xyz_raw = mean;
    float4 q_raw = make_float4(0.f, 0.f, 0.f, 0.f), q_in = q_raw;
    HeadCam hc;
    HeadFrame hf;
    Rot3 Rp;
    if constexpr (RAW) {
        q_raw = reinterpret_cast<const float4*>(p.rotations)[idx];
        q_in = q_raw;
        if (p.pose != nullptr) {
            Rp = rotation_of(p.pose);
            float xw[3];
            pose_apply_point(Rp, p.pose, mean.x, mean.y, mean.z, xw);
            mean = make_float3(xw[0], xw[1], xw[2]);
            q_in = pose_apply_quat(p.pose, q_raw);
        }
        hc = load_cam(p.view, p.campos);
        const float xw2[3] = {mean.x, mean.y, mean.z};
        const float sraw[3] = {p.scales[3 * idx], p.scales[3 * idx + 1], p.scales[3 * idx + 2]};
        hf = head_frame(hc, xw2, sraw, q_in, 0.f);
        hf.sig = p.conic_opacity[idx].w;  // the activated opacity the forward pass stored (visible splats only)
    }
    const float* cov6 = (p.cov3D_precomp ? p.cov3D_precomp : p.cov3D) + 6 * (size_t)idx;
    float c6[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) c6[i] = cov6[i];

    // ---- part 1: conic gradient through the 2D covariance (backward.cu:144-274) -----------------
    const float3 g_conic = make_float3(ggeo[4], ggeo[5], ggeo[6]);
    Cov2DFrame fr;
    ewa_frame(mean, p.focal_x, p.focal_y, p.tan_fovx, p.tan_fovy, c6, p.view, fr);
    const Mat3& T = fr.T;
    const Mat3& V = fr.V;
    const float gate_x = (fr.txtz < -fr.limx || fr.txtz > fr.limx) ? 0.f : 1.f;
    const float gate_y = (fr.tytz < -fr.limy || fr.tytz > fr.limy) ? 0.f : 1.f;

    const Mat3 c2 = mat3_mul(mat3_mul(mat3_transpose(T), mat3_transpose(V)), T);
    const float a = c2.m[0][0] + 0.3f;
    const float b = c2.m[0][1];
    const float c = c2.m[1][1] + 0.3f;
    const float denom = a * c - b * b;
    float g_a = 0.f, g_b = 0.f, g_c = 0.f;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);

    float g_cov[6];
    if (denom2inv != 0) {
        g_a = denom2inv * (-c * c * g_conic.x + 2 * b * c * g_conic.y + (denom - a * c) * g_conic.z);
        g_c = denom2inv * (-a * a * g_conic.z + 2 * a * b * g_conic.y + (denom - a * c) * g_conic.x);
        g_b = denom2inv * 2 * (b * c * g_conic.x - (denom + 2 * b * b) * g_conic.y + a * b * g_conic.z);

        // diagonal world-covariance entries
        g_cov[0] = (T.m[0][0] * T.m[0][0] * g_a + T.m[0][0] * T.m[1][0] * g_b + T.m[1][0] * T.m[1][0] * g_c);
        g_cov[3] = (T.m[0][1] * T.m[0][1] * g_a + T.m[0][1] * T.m[1][1] * g_b + T.m[1][1] * T.m[1][1] * g_c);
        g_cov[5] = (T.m[0][2] * T.m[0][2] * g_a + T.m[0][2] * T.m[1][2] * g_b + T.m[1][2] * T.m[1][2] * g_c);
        // off-diagonal entries appear twice in the symmetric matrix
        g_cov[1] = 2 * T.m[0][0] * T.m[0][1] * g_a + (T.m[0][0] * T.m[1][1] + T.m[0][1] * T.m[1][0]) * g_b +
                   2 * T.m[1][0] * T.m[1][1] * g_c;
        g_cov[2] = 2 * T.m[0][0] * T.m[0][2] * g_a + (T.m[0][0] * T.m[1][2] + T.m[0][2] * T.m[1][0]) * g_b +
                   2 * T.m[1][0] * T.m[1][2] * g_c;
        g_cov[4] = 2 * T.m[0][2] * T.m[0][1] * g_a + (T.m[0][1] * T.m[1][2] + T.m[0][2] * T.m[1][1]) * g_b +
                   2 * T.m[1][1] * T.m[1][2] * g_c;
    } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) g_cov[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (!RAW || p.dL_dcov3D != nullptr) put(p.dL_dcov3D + 6 * idx + i, g_cov[i], LSX_ACC_COV3D);

    // gradient w.r.t. the upper 2x3 block of T
    float tv0[3], tv1[3];  // (row of T) . V columns
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        tv0[k] = T.m[0][0] * V.m[k][0] + T.m[0][1] * V.m[k][1] + T.m[0][2] * V.m[k][2];
        tv1[k] = T.m[1][0] * V.m[k][0] + T.m[1][1] * V.m[k][1] + T.m[1][2] * V.m[k][2];
    }
    const float g_T00 = 2 * tv0[0] * g_a + tv1[0] * g_b;
    const float g_T01 = 2 * tv0[1] * g_a + tv1[1] * g_b;
    const float g_T02 = 2 * tv0[2] * g_a + tv1[2] * g_b;
    const float g_T10 = 2 * tv1[0] * g_c + tv0[0] * g_b;
    const float g_T11 = 2 * tv1[1] * g_c + tv0[1] * g_b;
    const float g_T12 = 2 * tv1[2] * g_c + tv0[2] * g_b;

    // T = W * J  ->  the four non-constant Jacobian entries
    const float* vm = p.view;
    const float g_J00 = vm[0] * g_T00 + vm[4] * g_T01 + vm[8] * g_T02;
    const float g_J02 = vm[2] * g_T00 + vm[6] * g_T01 + vm[10] * g_T02;
    const float g_J11 = vm[1] * g_T10 + vm[5] * g_T11 + vm[9] * g_T12;
    const float g_J12 = vm[2] * g_T10 + vm[6] * g_T11 + vm[10] * g_T12;

    const float tz = 1.f / fr.t.z;
    const float tz2 = tz * tz;
    const float tz3 = tz2 * tz;
    const float g_tx = gate_x * -p.focal_x * tz2 * g_J02;
    const float g_ty = gate_y * -p.focal_y * tz2 * g_J12;
    const float g_tz = -p.focal_x * tz2 * g_J00 - p.focal_y * tz2 * g_J11 + (2 * p.focal_x * fr.t.x) * tz3 * g_J02 +
                       (2 * p.focal_y * fr.t.y) * tz3 * g_J12;

    // t = view * mean  ->  multiply by the transposed rotation part
    float3 g_mean = make_float3(vm[0] * g_tx + vm[1] * g_ty + vm[2] * g_tz, vm[4] * g_tx + vm[5] * g_ty + vm[6] * g_tz,
                                vm[8] * g_tx + vm[9] * g_ty + vm[10] * g_tz);

    // ---- part 2: screen-space mean gradient through the projection (backward.cu:370-387) ---------
    {
        const float* pr = p.proj;
        const float4 mh = xform_point_4x4(mean, pr);
        const float mw = 1.0f / (mh.w + 0.0000001f);
        const float mul1 = (pr[0] * mean.x + pr[4] * mean.y + pr[8] * mean.z + pr[12]) * mw * mw;
        const float mul2 = (pr[1] * mean.x + pr[5] * mean.y + pr[9] * mean.z + pr[13]) * mw * mw;
        const float gx = ggeo[0], gy = ggeo[1];
        float3 d;
        d.x = (pr[0] * mw - pr[3] * mul1) * gx + (pr[1] * mw - pr[3] * mul2) * gy;
        d.y = (pr[4] * mw - pr[7] * mul1) * gx + (pr[5] * mw - pr[7] * mul2) * gy;
        d.z = (pr[8] * mw - pr[11] * mul1) * gx + (pr[9] * mw - pr[11] * mul2) * gy;
        g_mean.x += d.x;
        g_mean.y += d.y;
        g_mean.z += d.z;
    }

    // ---- part 3: SH colour gradient (backward.cu:20-139) ------------------------------------------
    if (p.shs != nullptr) {
        const float ox = mean.x - p.campos[0], oy = mean.y - p.campos[1], oz = mean.z - p.campos[2];
        const float len = sqrtf(ox * ox + oy * oy + oz * oz);
        const float x = ox / len, y = oy / len, z = oz / len;
        float shv[48];  // the coefficients in registers
        {
#pragma unroll
            for (int k = 0; k < 48; ++k) shv[k] = 0.f;
            if (p.D > 0) {  // only the direction gradient reads coefficients
                if ((n_sh & 3) == 0 && (reinterpret_cast<uintptr_t>(sh_row) & 15u) == 0) {
                    const float4* s4 = reinterpret_cast<const float4*>(sh_row);
#pragma unroll
                    for (int j = 0; j < 12; ++j)
                        if (4 * j < n_sh) {
                            const float4 v = __ldg(s4 + j);
                            shv[4 * j] = v.x;
                            shv[4 * j + 1] = v.y;
                            shv[4 * j + 2] = v.z;
                            shv[4 * j + 3] = v.w;
                        }
                } else {
#pragma unroll
                    for (int k = 0; k < 48; ++k)
                        if (k < n_sh) shv[k] = __ldg(sh_row + k);
                }
            }
        }
        const float* sh = shv;
        const unsigned cl = p.clamped[idx];
        float gc[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) gc[k] = grec[k] * (((cl >> k) & 1u) ? 0.f : 1.f);

        // basis values (the derivative of RGB w.r.t. each coefficient) and d(RGB)/d(dir)
        float basis[16];
        float ddx[3] = {0.f, 0.f, 0.f}, ddy[3] = {0.f, 0.f, 0.f}, ddz[3] = {0.f, 0.f, 0.f};
        int nb = 1;
        basis[0] = kSH0;
        if (p.D > 0) {
            nb = 4;
            basis[1] = -kSH1 * y;
            basis[2] = kSH1 * z;
            basis[3] = -kSH1 * x;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                ddx[k] = -kSH1 * sh[9 + k];
                ddy[k] = -kSH1 * sh[3 + k];
                ddz[k] = kSH1 * sh[6 + k];
            }
            if (p.D > 1) {
                nb = 9;
                const float xx = x * x, yy = y * y, zz = z * z;
                const float xy = x * y, yz = y * z, xz = x * z;
                basis[4] = kSH2[0] * xy;
                basis[5] = kSH2[1] * yz;
                basis[6] = kSH2[2] * (2.f * zz - xx - yy);
                basis[7] = kSH2[3] * xz;
                basis[8] = kSH2[4] * (xx - yy);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    ddx[k] += kSH2[0] * y * sh[12 + k] + kSH2[2] * 2.f * -x * sh[18 + k] + kSH2[3] * z * sh[21 + k] +
                              kSH2[4] * 2.f * x * sh[24 + k];
                    ddy[k] += kSH2[0] * x * sh[12 + k] + kSH2[1] * z * sh[15 + k] + kSH2[2] * 2.f * -y * sh[18 + k] +
                              kSH2[4] * 2.f * -y * sh[24 + k];
                    ddz[k] += kSH2[1] * y * sh[15 + k] + kSH2[2] * 2.f * 2.f * z * sh[18 + k] + kSH2[3] * x * sh[21 + k];
                }
                if (p.D > 2) {
                    nb = 16;
                    basis[9] = kSH3[0] * y * (3.f * xx - yy);
                    basis[10] = kSH3[1] * xy * z;
                    basis[11] = kSH3[2] * y * (4.f * zz - xx - yy);
                    basis[12] = kSH3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                    basis[13] = kSH3[4] * x * (4.f * zz - xx - yy);
                    basis[14] = kSH3[5] * z * (xx - yy);
                    basis[15] = kSH3[6] * x * (xx - 3.f * yy);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        ddx[k] += (kSH3[0] * sh[27 + k] * 3.f * 2.f * xy + kSH3[1] * sh[30 + k] * yz +
                                   kSH3[2] * sh[33 + k] * -2.f * xy + kSH3[3] * sh[36 + k] * -3.f * 2.f * xz +
                                   kSH3[4] * sh[39 + k] * (-3.f * xx + 4.f * zz - yy) + kSH3[5] * sh[42 + k] * 2.f * xz +
                                   kSH3[6] * sh[45 + k] * 3.f * (xx - yy));
                        ddy[k] += (kSH3[0] * sh[27 + k] * 3.f * (xx - yy) + kSH3[1] * sh[30 + k] * xz +
                                   kSH3[2] * sh[33 + k] * (-3.f * yy + 4.f * zz - xx) +
                                   kSH3[3] * sh[36 + k] * -3.f * 2.f * yz + kSH3[4] * sh[39 + k] * -2.f * xy +
                                   kSH3[5] * sh[42 + k] * -2.f * yz + kSH3[6] * sh[45 + k] * -3.f * 2.f * xy);
                        ddz[k] += (kSH3[1] * sh[30 + k] * xy + kSH3[2] * sh[33 + k] * 4.f * 2.f * yz +
                                   kSH3[3] * sh[36 + k] * 3.f * (2.f * zz - xx - yy) +
                                   kSH3[4] * sh[39 + k] * 4.f * 2.f * xz + kSH3[5] * sh[42 + k] * (xx - yy));
                    }
                }
            }
        }
        {
            if (dsh_out) {
                // element e = 3 j + k of the row is basis[j] * gc[k] (zero beyond the active degree)
                auto elem = [&](const int e) -> float {
                    const int j = e / 3, k = e - 3 * j;
                    return (j < nb) ? basis[j] * gc[k] : 0.f;
                };
                const bool acc = (p.accumulate & LSX_ACC_SH) != 0;
                if (dsh_vec) {
#pragma unroll
                    for (int i = 0; i < 12; ++i)
                        if (4 * i < n_sh) {
                            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (acc) o = *reinterpret_cast<const float4*>(dsh_out + 4 * i);
                            *reinterpret_cast<float4*>(dsh_out + 4 * i) =
                                make_float4(o.x + elem(4 * i), o.y + elem(4 * i + 1), o.z + elem(4 * i + 2), o.w + elem(4 * i + 3));
                        }
                } else {
#pragma unroll
                    for (int e = 0; e < 48; ++e)
                        if (e < n_sh) dsh_out[e] = acc ? dsh_out[e] + elem(e) : elem(e);
                }
            }
        }

        // direction gradient, then through the normalisation dir = v / |v|
        const float gdx = ddx[0] * gc[0] + ddx[1] * gc[1] + ddx[2] * gc[2];
        const float gdy = ddy[0] * gc[0] + ddy[1] * gc[1] + ddy[2] * gc[2];
        const float gdz = ddz[0] * gc[0] + ddz[1] * gc[1] + ddz[2] * gc[2];
        const float sum2 = ox * ox + oy * oy + oz * oz;
        const float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
        g_mean.x += ((+sum2 - ox * ox) * gdx - oy * ox * gdy - oz * ox * gdz) * inv32;
        g_mean.y += (-ox * oy * gdx + (sum2 - oy * oy) * gdy - oz * oy * gdz) * inv32;
        g_mean.z += (-ox * oz * gdx - oy * oz * gdy + (sum2 - oz * oz) * gdz) * inv32;
    }

    if constexpr (!RAW) {
        put(p.dL_dmeans3D + 3 * idx + 0, g_mean.x, LSX_ACC_MEANS3D);
        put(p.dL_dmeans3D + 3 * idx + 1, g_mean.y, LSX_ACC_MEANS3D);
        put(p.dL_dmeans3D + 3 * idx + 2, g_mean.z, LSX_ACC_MEANS3D);
    }

    // ---- part 4: world covariance -> scale / rotation (backward.cu:278-341) ------------------------
    float g_sc[3] = {0.f, 0.f, 0.f};                     // RAW: dL/d(activated scale), dL/d(normalised quaternion)
    float4 g_qn = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.scales != nullptr) {
        const float3 sc = RAW ? make_float3(hf.sc[0], hf.sc[1], hf.sc[2])
                              : make_float3(p.scales[3 * idx], p.scales[3 * idx + 1], p.scales[3 * idx + 2]);
        const float4 q = RAW ? make_float4(hf.q[0], hf.q[1], hf.q[2], hf.q[3]) : reinterpret_cast<const float4*>(p.rotations)[idx];
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        Mat3 R;
        R.m[0][0] = 1.f - 2.f * (y * y + z * z);
        R.m[0][1] = 2.f * (x * y - r * z);
        R.m[0][2] = 2.f * (x * z + r * y);
        R.m[1][0] = 2.f * (x * y + r * z);
        R.m[1][1] = 1.f - 2.f * (x * x + z * z);
        R.m[1][2] = 2.f * (y * z - r * x);
        R.m[2][0] = 2.f * (x * z - r * y);
        R.m[2][1] = 2.f * (y * z + r * x);
        R.m[2][2] = 1.f - 2.f * (x * x + y * y);
        const float3 s = make_float3(p.scale_modifier * sc.x, p.scale_modifier * sc.y, p.scale_modifier * sc.z);
        Mat3 S;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) S.m[cc][rr] = 0.f;
        S.m[0][0] = s.x;
        S.m[1][1] = s.y;
        S.m[2][2] = s.z;
        const Mat3 Mx = mat3_mul(S, R);

        Mat3 gS;  // symmetric dL/dSigma with halved off-diagonals
        gS.m[0][0] = g_cov[0];        gS.m[0][1] = 0.5f * g_cov[1]; gS.m[0][2] = 0.5f * g_cov[2];
        gS.m[1][0] = 0.5f * g_cov[1]; gS.m[1][1] = g_cov[3];        gS.m[1][2] = 0.5f * g_cov[4];
        gS.m[2][0] = 0.5f * g_cov[2]; gS.m[2][1] = 0.5f * g_cov[4]; gS.m[2][2] = g_cov[5];

        Mat3 M2;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) M2.m[cc][rr] = 2.0f * Mx.m[cc][rr];
        const Mat3 gM = mat3_mul(M2, gS);
        const Mat3 Rt = mat3_transpose(R);
        Mat3 gMt = mat3_transpose(gM);

        g_sc[0] = Rt.m[0][0] * gMt.m[0][0] + Rt.m[0][1] * gMt.m[0][1] + Rt.m[0][2] * gMt.m[0][2];
        if constexpr (!RAW) put(p.dL_dscales + 3 * idx + 0, g_sc[0], LSX_ACC_SCALES);
        g_sc[1] = Rt.m[1][0] * gMt.m[1][0] + Rt.m[1][1] * gMt.m[1][1] + Rt.m[1][2] * gMt.m[1][2];
        if constexpr (!RAW) put(p.dL_dscales + 3 * idx + 1, g_sc[1], LSX_ACC_SCALES);
        g_sc[2] = Rt.m[2][0] * gMt.m[2][0] + Rt.m[2][1] * gMt.m[2][1] + Rt.m[2][2] * gMt.m[2][2];
        if constexpr (!RAW) put(p.dL_dscales + 3 * idx + 2, g_sc[2], LSX_ACC_SCALES);

#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
            gMt.m[0][rr] *= s.x;
            gMt.m[1][rr] *= s.y;
            gMt.m[2][rr] *= s.z;
        }
        float4 gq;
        gq.x = 2 * z * (gMt.m[0][1] - gMt.m[1][0]) + 2 * y * (gMt.m[2][0] - gMt.m[0][2]) + 2 * x * (gMt.m[1][2] - gMt.m[2][1]);
        gq.y = 2 * y * (gMt.m[1][0] + gMt.m[0][1]) + 2 * z * (gMt.m[2][0] + gMt.m[0][2]) + 2 * r * (gMt.m[1][2] - gMt.m[2][1]) -
               4 * x * (gMt.m[2][2] + gMt.m[1][1]);
        gq.z = 2 * x * (gMt.m[1][0] + gMt.m[0][1]) + 2 * r * (gMt.m[2][0] - gMt.m[0][2]) + 2 * z * (gMt.m[1][2] + gMt.m[2][1]) -
               4 * y * (gMt.m[2][2] + gMt.m[0][0]);
        gq.w = 2 * r * (gMt.m[0][1] - gMt.m[1][0]) + 2 * x * (gMt.m[2][0] + gMt.m[0][2]) + 2 * y * (gMt.m[1][2] + gMt.m[2][1]) -
               4 * z * (gMt.m[1][1] + gMt.m[0][0]);
        g_qn = gq;
        if constexpr (!RAW) {
            if (am & LSX_ACC_ROTATIONS) {
                const float4 o = reinterpret_cast<const float4*>(p.dL_drotations)[idx];
                gq = make_float4(gq.x + o.x, gq.y + o.y, gq.z + o.z, gq.w + o.w);
            }
            reinterpret_cast<float4*>(p.dL_drotations)[idx] = gq;  // w.r.t. the raw (un-normalised) quaternion
        }
    } else if (!RAW) {
        if (!(am & LSX_ACC_SCALES)) {
#pragma unroll
            for (int i = 0; i < 3; ++i) p.dL_dscales[3 * idx + i] = 0.f;
        }
        if (!(am & LSX_ACC_ROTATIONS)) reinterpret_cast<float4*>(p.dL_drotations)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    if constexpr (RAW) {
        // ---- part 5: the render wrapper's backward (activations, plane normal, all_map, camera pose) ----
        const float gm[3] = {g_mean.x, g_mean.y, g_mean.z};
        const float gr[4] = {g_qn.x, g_qn.y, g_qn.z, g_qn.w};
        const HeadGrads hg = head_backward_row(hc, hf, q_in, g_sc, gr, ggeo[7], gmap, gm);
        float d_xyz[3] = {hg.xyz[0], hg.xyz[1], hg.xyz[2]};
        float4 d_q = hg.qraw;
        if (p.pose != nullptr) {
            pose_backward_point(Rp, xyz_raw.x, xyz_raw.y, xyz_raw.z, hg.xyz[0], hg.xyz[1], hg.xyz[2], d_xyz, pose_acc);
            d_q = pose_backward_quat(p.pose, q_raw, hg.qraw, pose_acc);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) put(p.dL_dmeans3D + 3 * idx + i, d_xyz[i], LSX_ACC_MEANS3D);
#pragma unroll
        for (int i = 0; i < 3; ++i) put(p.dL_dscales + 3 * idx + i, hg.sraw[i], LSX_ACC_SCALES);
        if (am & LSX_ACC_ROTATIONS) {
            const float4 o = reinterpret_cast<const float4*>(p.dL_drotations)[idx];
            d_q = make_float4(d_q.x + o.x, d_q.y + o.y, d_q.z + o.z, d_q.w + o.w);
        }
        reinterpret_cast<float4*>(p.dL_drotations)[idx] = d_q;
        put(p.dL_dopacity + idx, hg.oraw, LSX_ACC_OPACITY);
    }
}

// own-row copy of `n` record columns -> the row of a (P, n) gradient tensor (accumulate: +=)
__device__ __forceinline__ void row_put(float* __restrict__ dst, const float* __restrict__ src, const int n, const bool acc) {
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
#pragma unroll 4
        for (int k = 0; k < n; k += 4) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (acc) o = *reinterpret_cast<const float4*>(dst + k);
            *reinterpret_cast<float4*>(dst + k) =
                make_float4(o.x + __ldg(src + k), o.y + __ldg(src + k + 1), o.z + __ldg(src + k + 2), o.w + __ldg(src + k + 3));
        }
    } else {
#pragma unroll 5
        for (int k = 0; k < n; ++k) dst[k] = acc ? dst[k] + __ldg(src + k) : __ldg(src + k);
    }
}
#ifndef LSX_PRE_BWD_MINB
#define LSX_PRE_BWD_MINB 0
#endif
#ifndef LSX_PRE_BWD_MINB_RAW
#define LSX_PRE_BWD_MINB_RAW 0
#endif
// One Gaussian per thread, every row read and written by its owner (see the forward kernel): the gradient record's columns go
// to the rows of the reference's gradient tensors, dL/dsh goes from registers to its row.  No shared-memory slab: the block
// only meets for the pose-gradient partial sums of the fused-wrapper mode.
template <bool RAW>
__global__ void __launch_bounds__(kPreBwdThreads, RAW ? LSX_PRE_BWD_MINB_RAW : LSX_PRE_BWD_MINB) preprocess_bwd_kernel(const PreprocessBwdParams p) {
    const int n_sh = p.M * 3;
    const int idx = blockIdx.x * kPreBwdThreads + threadIdx.x;
    const unsigned am = (unsigned)p.accumulate;
    float pose_acc[kPoseTerms];
#pragma unroll
    for (int k = 0; k < kPoseTerms; ++k) pose_acc[k] = 0.f;
    if (idx < p.P) {
        const float* rec = p.grad_records + (size_t)idx * p.grad_stride;
        float gcol[3], ggeo[8], gmap[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (RAW && p.render_geo) {  // the all_map gradient goes on through the wrapper's backward instead of out
            const int cm = 3 + (p.include_feature ? p.F + p.Fi : 0);
#pragma unroll
            for (int k = 0; k < 5; ++k) gmap[k] = __ldg(rec + cm + k);
        }
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(rec));
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(rec + p.n_channels_pad));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(rec + p.n_channels_pad + 4));
        gcol[0] = c4.x; gcol[1] = c4.y; gcol[2] = c4.z;
        ggeo[0] = g0.x; ggeo[1] = g0.y; ggeo[2] = g0.z; ggeo[3] = g0.w;
        ggeo[4] = g1.x; ggeo[5] = g1.y; ggeo[6] = g1.z; ggeo[7] = g1.w;
        ggeo[0] *= p.geo_scale_x; ggeo[1] *= p.geo_scale_y; ggeo[2] *= p.geo_scale_x; ggeo[3] *= p.geo_scale_y;
        ggeo[4] *= -0.5f; ggeo[5] *= -0.5f; ggeo[6] *= -0.5f;
        row_put(p.dL_dcolor + (size_t)idx * 3, rec, 3, (am & LSX_ACC_COLORS) != 0);
        int c = 3;
        if (p.include_feature) {
            row_put(p.dL_dlanguage_feature + (size_t)idx * p.F, rec + c, p.F, (am & LSX_ACC_LANG) != 0);
            c += p.F;
            row_put(p.dL_dlanguage_feature_instance + (size_t)idx * p.Fi, rec + c, p.Fi, (am & LSX_ACC_INST) != 0);
            c += p.Fi;
        }
        if (RAW && p.dL_dall_map == nullptr) {
            // raw mode: all_map is not an input of the call, its gradient is consumed by preprocess_bwd_row
        } else if (p.render_geo) {
            row_put(p.dL_dall_map + (size_t)idx * 5, rec + c, 5, (am & LSX_ACC_ALL_MAP) != 0);
        } else if (!(am & LSX_ACC_ALL_MAP)) {
#pragma unroll
            for (int k = 0; k < 5; ++k) p.dL_dall_map[(size_t)idx * 5 + k] = 0.f;
        }
        preprocess_bwd_row<RAW>(p, idx, p.shs ? p.shs + (size_t)idx * n_sh : nullptr, gcol, ggeo, gmap, pose_acc);
    }
    if (RAW && p.pose != nullptr && p.pose_partials != nullptr) {
        // this block's row of pose-gradient partial sums (fixed order: deterministic); lsx::launch_pose_finish adds the rows
        __shared__ float s_pose[kPreBwdThreads / 32][kPoseTerms];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < kPoseTerms; ++k) {
            float v = pose_acc[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) s_pose[warp][k] = v;
        }
        __syncthreads();
        if (threadIdx.x < kPoseTerms) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < kPreBwdThreads / 32; ++w) v += s_pose[w][threadIdx.x];
            p.pose_partials[(size_t)blockIdx.x * kPoseTerms + threadIdx.x] = v;
        }
    }
}

__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* __restrict__ means3D,
                                                           const float* __restrict__ view, uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float3 pw = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    const float3 pv = xform_point_4x3(pw, view);
    present[idx] = pv.z > 0.2f ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
// The opt-in dynamic shared-memory limit is a per-device attribute of the function: remember what has been set for
// (kernel slot, device) so that a process driving several GPUs configures each of them (benign race: monotone maximum).
static cudaError_t ensure_dynamic_smem(const void* func, size_t bytes, int slot) {
    static size_t configured[4][64] = {};
    if (bytes <= 48 * 1024) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const int d = dev < 0 ? 0 : (dev > 63 ? 63 : dev);
    if (dev > 63 || bytes > configured[slot][d]) {
        e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e == cudaSuccess) configured[slot][d] = bytes;
    }
    return e;
}

int launch_preprocess_fwd(const PreprocessFwdParams& p, cudaStream_t stream, bool debug) {
    if (p.P <= 0) return 0;
    const size_t smem = (size_t)kPreFwdThreads * record_tile_row(p.rec_stride) * sizeof(float);  // the threads' record rows
    if (p.raw_params) {
        LSX_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(preprocess_fwd_kernel<true>), smem, 2));
        preprocess_fwd_kernel<true><<<ceil_div(p.P, kPreFwdThreads), kPreFwdThreads, smem, stream>>>(p);
    } else {
        LSX_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(preprocess_fwd_kernel<false>), smem, 0));
        preprocess_fwd_kernel<false><<<ceil_div(p.P, kPreFwdThreads), kPreFwdThreads, smem, stream>>>(p);
    }
    LSX_KERNEL_OK(stream, debug);
    return 0;
}

int launch_preprocess_bwd(const PreprocessBwdParams& p, cudaStream_t stream, bool debug) {
    if (p.P <= 0) return 0;
    const size_t smem = 0;
    const int blocks = ceil_div(p.P, kPreBwdThreads);
    if (p.raw_params) {
        LSX_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(preprocess_bwd_kernel<true>), smem, 3));
        preprocess_bwd_kernel<true><<<blocks, kPreBwdThreads, smem, stream>>>(p);
        LSX_KERNEL_OK(stream, debug);
        if (p.pose != nullptr && p.dL_dpose != nullptr) {
            const int rc = launch_pose_finish(blocks, p.pose, p.pose_partials, p.dL_dpose, p.accumulate_pose, stream);
            if (rc) return rc;
        }
    } else {
        LSX_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(preprocess_bwd_kernel<false>), smem, 1));
        preprocess_bwd_kernel<false><<<blocks, kPreBwdThreads, smem, stream>>>(p);
        LSX_KERNEL_OK(stream, debug);
    }
    return 0;
}

int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present, cudaStream_t stream) {
    if (P <= 0) return 0;
    mark_visible_kernel<<<ceil_div(P, 256), 256, 0, stream>>>(P, means3D, view, present);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

}  // namespace lsx

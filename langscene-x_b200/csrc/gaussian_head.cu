// gaussian_head.cu — the per-Gaussian half of LangScene-X's render wrapper, forward and backward.
//
// Reference behaviour restated:
//   activations   scales = exp(_scaling), rotations = F.normalize(_rotation), opacity = sigmoid(_opacity)
//                 (field_construction/scene/gaussian_model.py:53-61,193-213)
//   plane normal  column argmin(scales) of quaternion_to_matrix(rotations) (pytorch3d.transforms, dependency absent from the
//                 reference tree: R = I + two_s * [...], two_s = 2 / |q|^2, q = (r, i, j, k)), flipped towards the camera
//                 (gaussian_model.py:225-236 get_smallest_axis / get_normal)
//   all_map       [normal @ W2V[:3,:3], 1, |<local normal, xyz @ W2V[:3,:3] + W2V[3,:3]>|]
//                 (field_construction/gaussian_renderer/__init__.py:188-196)
// The reference spends ~15 torch kernels forward and ~30 backward on this, each a full pass over P; here one streaming
// kernel each way (forward 44 B read + 52 B written per Gaussian, backward 96 B read + 44 B written).
#include "../../include/lsx_rasterizer.h"
#include "kernels.cuh"

namespace lsx {
namespace {

struct HeadCam {
    float v[16];   // world_view_transform, row-major as torch stores it (row-vector convention)
    float cam[3];  // camera centre
};

struct HeadFrame {  // everything both passes need about one Gaussian
    float sc[3], q[4], nq, sig;
    int idx;
    float flip;
    float ng0[3];  // un-flipped normal = column idx of R
    float nl[3], pc[3], d;
};

__device__ __forceinline__ void rot_column(const float* q, const int idx, float* col) {
    const float r = q[0], i = q[1], j = q[2], k = q[3];
    const float s2 = 2.0f / (r * r + i * i + j * j + k * k);
    if (idx == 0) {
        col[0] = 1.f - s2 * (j * j + k * k); col[1] = s2 * (i * j + k * r); col[2] = s2 * (i * k - j * r);
    } else if (idx == 1) {
        col[0] = s2 * (i * j - k * r); col[1] = 1.f - s2 * (i * i + k * k); col[2] = s2 * (j * k + i * r);
    } else {
        col[0] = s2 * (i * k + j * r); col[1] = s2 * (j * k - i * r); col[2] = 1.f - s2 * (i * i + j * j);
    }
}

__device__ __forceinline__ HeadFrame head_frame(const HeadCam& c, const float* xyz, const float* sraw, const float4 qraw,
                                                const float oraw) {
    HeadFrame f;
#pragma unroll
    for (int a = 0; a < 3; ++a) f.sc[a] = expf(sraw[a]);
    f.nq = fmaxf(sqrtf(qraw.x * qraw.x + qraw.y * qraw.y + qraw.z * qraw.z + qraw.w * qraw.w), 1.0e-12f);
    f.q[0] = qraw.x / f.nq; f.q[1] = qraw.y / f.nq; f.q[2] = qraw.z / f.nq; f.q[3] = qraw.w / f.nq;
    f.sig = 1.0f / (1.0f + expf(-oraw));
    f.idx = 0;  // torch.min: first index of the minimum
    if (f.sc[1] < f.sc[f.idx]) f.idx = 1;
    if (f.sc[2] < f.sc[f.idx]) f.idx = 2;
    rot_column(f.q, f.idx, f.ng0);
    const float dot = f.ng0[0] * (c.cam[0] - xyz[0]) + f.ng0[1] * (c.cam[1] - xyz[1]) + f.ng0[2] * (c.cam[2] - xyz[2]);
    f.flip = dot < 0.0f ? -1.0f : 1.0f;
#pragma unroll
    for (int col = 0; col < 3; ++col) {
        f.nl[col] = f.flip * (f.ng0[0] * c.v[col] + f.ng0[1] * c.v[4 + col] + f.ng0[2] * c.v[8 + col]);
        f.pc[col] = xyz[0] * c.v[col] + xyz[1] * c.v[4 + col] + xyz[2] * c.v[8 + col] + c.v[12 + col];
    }
    f.d = f.nl[0] * f.pc[0] + f.nl[1] * f.pc[1] + f.nl[2] * f.pc[2];
    return f;
}

__device__ __forceinline__ HeadCam load_cam(const float* __restrict__ view, const float* __restrict__ campos) {
    HeadCam c;
#pragma unroll
    for (int k = 0; k < 16; ++k) c.v[k] = __ldg(view + k);
#pragma unroll
    for (int k = 0; k < 3; ++k) c.cam[k] = __ldg(campos + k);
    return c;
}

__global__ void __launch_bounds__(256) gaussian_head_fwd_kernel(const int P, const float* __restrict__ view,
                                                                const float* __restrict__ campos, const float* __restrict__ xyz,
                                                                const float* __restrict__ scaling_raw,
                                                                const float* __restrict__ rotation_raw,
                                                                const float* __restrict__ opacity_raw, float* __restrict__ scales,
                                                                float* __restrict__ rotations, float* __restrict__ opacity,
                                                                float* __restrict__ all_map) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const HeadCam c = load_cam(view, campos);
    const float x[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    const float s[3] = {scaling_raw[3 * i], scaling_raw[3 * i + 1], scaling_raw[3 * i + 2]};
    const HeadFrame f = head_frame(c, x, s, reinterpret_cast<const float4*>(rotation_raw)[i], opacity_raw[i]);
#pragma unroll
    for (int a = 0; a < 3; ++a) scales[3 * i + a] = f.sc[a];
    reinterpret_cast<float4*>(rotations)[i] = make_float4(f.q[0], f.q[1], f.q[2], f.q[3]);
    opacity[i] = f.sig;
    all_map[5 * i + 0] = f.nl[0];
    all_map[5 * i + 1] = f.nl[1];
    all_map[5 * i + 2] = f.nl[2];
    all_map[5 * i + 3] = 1.0f;
    all_map[5 * i + 4] = fabsf(f.d);
}

__global__ void __launch_bounds__(256) gaussian_head_bwd_kernel(
    const int P, const float* __restrict__ view, const float* __restrict__ campos, const float* __restrict__ xyz, const float* __restrict__ scaling_raw,
    const float* __restrict__ rotation_raw, const float* __restrict__ opacity_raw, const float* __restrict__ g_scales,
    const float* __restrict__ g_rotations, const float* __restrict__ g_opacity, const float* __restrict__ g_all_map,
    const float* __restrict__ g_means3D, float* __restrict__ g_xyz, float* __restrict__ g_scaling_raw,
    float* __restrict__ g_rotation_raw, float* __restrict__ g_opacity_raw, const unsigned acc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    // acc: bit 0 xyz, 1 scaling, 2 rotation, 3 opacity — add to the buffer instead of writing it
    auto put = [acc](float* dst, const float v, const unsigned bit) { *dst = (acc & bit) ? *dst + v : v; };
    const HeadCam c = load_cam(view, campos);
    const float x[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    const float s[3] = {scaling_raw[3 * i], scaling_raw[3 * i + 1], scaling_raw[3 * i + 2]};
    const float4 qraw = reinterpret_cast<const float4*>(rotation_raw)[i];
    const HeadFrame f = head_frame(c, x, s, qraw, opacity_raw[i]);

    // activations
#pragma unroll
    for (int a = 0; a < 3; ++a) put(g_scaling_raw + 3 * i + a, (g_scales ? g_scales[3 * i + a] : 0.f) * f.sc[a], 2u);
    put(g_opacity_raw + i, (g_opacity ? g_opacity[i] : 0.f) * f.sig * (1.0f - f.sig), 8u);

    // all_map -> local normal, camera-space point
    float gam[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (g_all_map) {
#pragma unroll
        for (int a = 0; a < 5; ++a) gam[a] = g_all_map[5 * i + a];
    }
    const float sgn = f.d > 0.f ? 1.f : (f.d < 0.f ? -1.f : 0.f);
    float g_nl[3], g_pc[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        g_nl[a] = gam[a] + gam[4] * sgn * f.pc[a];
        g_pc[a] = gam[4] * sgn * f.nl[a];
    }
    // back through the view rotation (row r of W2V[:3,:3] dotted with the camera-space gradient)
    float g_col[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        g_col[r] = f.flip * (g_nl[0] * c.v[4 * r] + g_nl[1] * c.v[4 * r + 1] + g_nl[2] * c.v[4 * r + 2]);
        const float gx = g_pc[0] * c.v[4 * r] + g_pc[1] * c.v[4 * r + 1] + g_pc[2] * c.v[4 * r + 2];
        put(g_xyz + 3 * i + r, gx + (g_means3D ? g_means3D[3 * i + r] : 0.f), 1u);
    }
    // column idx of R(q): entry e = const + s2 * fe(q);  d e / d q_m = s2 * d fe / d q_m - s2^2 * q_m * fe   (|q| = 1 here,
    // but pytorch3d's two_s = 2 / |q|^2 is differentiated too)
    const float r = f.q[0], qi = f.q[1], qj = f.q[2], qk = f.q[3];
    const float s2 = 2.0f / (r * r + qi * qi + qj * qj + qk * qk);
    float fe[3], dfe[3][4];  // dfe[e][m], m over (r, i, j, k)
    if (f.idx == 0) {
        fe[0] = -(qj * qj + qk * qk); dfe[0][0] = 0.f;  dfe[0][1] = 0.f;  dfe[0][2] = -2.f * qj; dfe[0][3] = -2.f * qk;
        fe[1] = qi * qj + qk * r;     dfe[1][0] = qk;   dfe[1][1] = qj;   dfe[1][2] = qi;        dfe[1][3] = r;
        fe[2] = qi * qk - qj * r;     dfe[2][0] = -qj;  dfe[2][1] = qk;   dfe[2][2] = -r;        dfe[2][3] = qi;
    } else if (f.idx == 1) {
        fe[0] = qi * qj - qk * r;     dfe[0][0] = -qk;  dfe[0][1] = qj;   dfe[0][2] = qi;        dfe[0][3] = -r;
        fe[1] = -(qi * qi + qk * qk); dfe[1][0] = 0.f;  dfe[1][1] = -2.f * qi; dfe[1][2] = 0.f;  dfe[1][3] = -2.f * qk;
        fe[2] = qj * qk + qi * r;     dfe[2][0] = qi;   dfe[2][1] = r;    dfe[2][2] = qk;        dfe[2][3] = qj;
    } else {
        fe[0] = qi * qk + qj * r;     dfe[0][0] = qj;   dfe[0][1] = qk;   dfe[0][2] = r;         dfe[0][3] = qi;
        fe[1] = qj * qk - qi * r;     dfe[1][0] = -qi;  dfe[1][1] = -r;   dfe[1][2] = qk;        dfe[1][3] = qj;
        fe[2] = -(qi * qi + qj * qj); dfe[2][0] = 0.f;  dfe[2][1] = -2.f * qi; dfe[2][2] = -2.f * qj; dfe[2][3] = 0.f;
    }
    float gq[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        float acc = g_rotations ? g_rotations[4 * i + m] : 0.f;
#pragma unroll
        for (int e = 0; e < 3; ++e) acc += g_col[e] * (s2 * dfe[e][m] - s2 * s2 * f.q[m] * fe[e]);
        gq[m] = acc;
    }
    // F.normalize backward
    const float norm_raw = sqrtf(qraw.x * qraw.x + qraw.y * qraw.y + qraw.z * qraw.z + qraw.w * qraw.w);
    float4 out;
    if (norm_raw > 1.0e-12f) {
        const float qg = f.q[0] * gq[0] + f.q[1] * gq[1] + f.q[2] * gq[2] + f.q[3] * gq[3];
        out = make_float4((gq[0] - f.q[0] * qg) / f.nq, (gq[1] - f.q[1] * qg) / f.nq, (gq[2] - f.q[2] * qg) / f.nq,
                          (gq[3] - f.q[3] * qg) / f.nq);
    } else {
        out = make_float4(gq[0] / f.nq, gq[1] / f.nq, gq[2] / f.nq, gq[3] / f.nq);  // q = q_raw / eps
    }
    if (acc & 4u) {
        const float4 o = reinterpret_cast<const float4*>(g_rotation_raw)[i];
        out = make_float4(out.x + o.x, out.y + o.y, out.z + o.z, out.w + o.w);
    }
    reinterpret_cast<float4*>(g_rotation_raw)[i] = out;
}

}  // namespace
}  // namespace lsx

using namespace lsx;

extern "C" int lsx_gaussian_head_forward(int32_t P, const float* viewmatrix, const float* campos, const float* xyz,
                                         const float* scaling_raw, const float* rotation_raw, const float* opacity_raw,
                                         float* scales, float* rotations, float* opacity, float* all_map, void* stream_) {
    if (P < 0 || !viewmatrix || !campos ||
        (P > 0 && (!xyz || !scaling_raw || !rotation_raw || !opacity_raw || !scales || !rotations || !opacity || !all_map))) {
        set_error("lsx_gaussian_head_forward: bad arguments");
        return -1;
    }
    if (P == 0) return 0;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    gaussian_head_fwd_kernel<<<ceil_div(P, 256), 256, 0, stream>>>(P, viewmatrix, campos, xyz, scaling_raw, rotation_raw,
                                                                   opacity_raw, scales, rotations, opacity, all_map);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

extern "C" int lsx_gaussian_head_backward_acc(int32_t P, const float* viewmatrix, const float* campos, const float* xyz,
                                              const float* scaling_raw, const float* rotation_raw, const float* opacity_raw,
                                              const float* dL_dscales, const float* dL_drotations, const float* dL_dopacity,
                                              const float* dL_dall_map, const float* dL_dmeans3D, float* dL_dxyz,
                                              float* dL_dscaling_raw, float* dL_drotation_raw, float* dL_dopacity_raw,
                                              int32_t accumulate_mask, void* stream_) {
    if (P < 0 || !viewmatrix || !campos ||
        (P > 0 && (!xyz || !scaling_raw || !rotation_raw || !opacity_raw || !dL_dxyz || !dL_dscaling_raw || !dL_drotation_raw ||
                   !dL_dopacity_raw))) {
        set_error("lsx_gaussian_head_backward: bad arguments");
        return -1;
    }
    if (P == 0) return 0;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    gaussian_head_bwd_kernel<<<ceil_div(P, 256), 256, 0, stream>>>(P, viewmatrix, campos, xyz, scaling_raw, rotation_raw,
                                                                   opacity_raw, dL_dscales, dL_drotations,
                                                                   dL_dopacity, dL_dall_map, dL_dmeans3D, dL_dxyz, dL_dscaling_raw,
                                                                   dL_drotation_raw, dL_dopacity_raw, (unsigned)accumulate_mask);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

extern "C" int lsx_gaussian_head_backward(int32_t P, const float* viewmatrix, const float* campos, const float* xyz,
                                          const float* scaling_raw, const float* rotation_raw, const float* opacity_raw,
                                          const float* dL_dscales, const float* dL_drotations, const float* dL_dopacity,
                                          const float* dL_dall_map, const float* dL_dmeans3D, float* dL_dxyz,
                                          float* dL_dscaling_raw, float* dL_drotation_raw, float* dL_dopacity_raw, void* stream_) {
    return lsx_gaussian_head_backward_acc(P, viewmatrix, campos, xyz, scaling_raw, rotation_raw, opacity_raw, dL_dscales,
                                          dL_drotations, dL_dopacity, dL_dall_map, dL_dmeans3D, dL_dxyz, dL_dscaling_raw,
                                          dL_drotation_raw, dL_dopacity_raw, 0, stream_);
}

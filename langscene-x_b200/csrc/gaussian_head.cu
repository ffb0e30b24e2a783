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
#include "head_math.cuh"
#include "kernels.cuh"

namespace lsx {
namespace {

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

    float gs[3], gr[4], gam[5], gm[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) gs[a] = g_scales ? g_scales[3 * i + a] : 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) gr[a] = g_rotations ? g_rotations[4 * i + a] : 0.f;
#pragma unroll
    for (int a = 0; a < 5; ++a) gam[a] = g_all_map ? g_all_map[5 * i + a] : 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) gm[a] = g_means3D ? g_means3D[3 * i + a] : 0.f;
    const HeadGrads o = head_backward_row(c, f, qraw, gs, gr, g_opacity ? g_opacity[i] : 0.f, gam, gm);
#pragma unroll
    for (int a = 0; a < 3; ++a) put(g_scaling_raw + 3 * i + a, o.sraw[a], 2u);
    put(g_opacity_raw + i, o.oraw, 8u);
#pragma unroll
    for (int a = 0; a < 3; ++a) put(g_xyz + 3 * i + a, o.xyz[a], 1u);
    float4 out = o.qraw;
    if (acc & 4u) {
        const float4 prev = reinterpret_cast<const float4*>(g_rotation_raw)[i];
        out = make_float4(out.x + prev.x, out.y + prev.y, out.z + prev.z, out.w + prev.w);
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
    if ((reinterpret_cast<uintptr_t>(rotation_raw) | reinterpret_cast<uintptr_t>(rotations)) & 15u) {
        set_error("lsx_gaussian_head_forward: the (P, 4) rotation arrays must be 16-byte aligned");
        return -1;
    }
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
    if ((reinterpret_cast<uintptr_t>(rotation_raw) | reinterpret_cast<uintptr_t>(dL_drotations) |
         reinterpret_cast<uintptr_t>(dL_drotation_raw)) & 15u) {
        set_error("lsx_gaussian_head_backward: the (P, 4) rotation arrays must be 16-byte aligned");
        return -1;
    }
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

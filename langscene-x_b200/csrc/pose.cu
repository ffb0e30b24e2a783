// pose.cu — camera-pose transform of the Gaussians ("pose transform rel_w2c @ xyz + quadmultiply")
//           and the masked L1 of the language-feature loss (8(f) rank 2).
//
// Reference behaviour restated
//   render(..., camera_pose=pose)  field_construction/gaussian_renderer/__init__.py:79-87:
//       rel_w2c = get_camera_from_tensor(pose)            field_construction/utils/pose_utils.py:60-87 (quad2rotation :13-58)
//       means3D   = (rel_w2c @ [xyz 1]^T)^T[:, :3]        = R(q/|q|) xyz + T         pose = [q(4) | T(3)]
//       rotations = quadmultiply(pose[:4], _rotation)     pose_utils.py:89-107 — Hamilton product with the RAW pose quaternion
//   (~12 torch kernels forward, ~25 backward incl. two (P,4)x(4,4) matmuls; the pose gradient is a reduction over all P rows).
//   language loss  field_construction/gaussian_field.py:450-451 with l1_loss = mean |a - b| (utils/loss_utils.py:20-21):
//       l1_loss(language_feature * mask, gt * mask)       mask (1,H,W) or (F,H,W), broadcast over the feature channels
//
// B200 design: one streaming kernel each way (HBM-bound: 28 B in + 28 B out per Gaussian forward; 56 B in + 28 B out backward).
// The backward pass folds the whole pose gradient into the same pass: every thread keeps 16 partial sums (dL/dR 3x3 as the
// outer product g x^T, dL/dT, and the four quaternion-product terms), blocks reduce them with shuffles into a partials array,
// and a one-block epilogue adds the partials in a fixed order (deterministic) and applies the chain rule through R(q/|q|) in
// double precision.  The masked L1 does the same with per-block partial sums; no atomics anywhere.
#include "../../include/lsx_rasterizer.h"
#include "head_math.cuh"
#include "kernels.cuh"

namespace lsx {
namespace {

constexpr int kPoseBlocks = 148 * 4;   // partial rows: one per block of the backward grid (grid-stride over P)
constexpr int kL1Blocks = 148 * 8;

__global__ void __launch_bounds__(256) pose_fwd_kernel(const int P, const float* __restrict__ pose, const float* __restrict__ xyz,
                                                       const float* __restrict__ rot, float* __restrict__ out_xyz,
                                                       float* __restrict__ out_rot) {
    const Rot3 R = rotation_of(pose);
    const float w1 = pose[0], x1 = pose[1], y1 = pose[2], z1 = pose[3];
    const float tx = pose[4], ty = pose[5], tz = pose[6];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
        out_xyz[3 * i + 0] = R.m[0][0] * px + R.m[0][1] * py + R.m[0][2] * pz + tx;
        out_xyz[3 * i + 1] = R.m[1][0] * px + R.m[1][1] * py + R.m[1][2] * pz + ty;
        out_xyz[3 * i + 2] = R.m[2][0] * px + R.m[2][1] * py + R.m[2][2] * pz + tz;
        if (rot) {
            const float4 q = *reinterpret_cast<const float4*>(rot + 4ll * i);  // (w2, x2, y2, z2)
            float4 o;
            o.x = w1 * q.x - x1 * q.y - y1 * q.z - z1 * q.w;
            o.y = w1 * q.y + x1 * q.x + y1 * q.w - z1 * q.z;
            o.z = w1 * q.z - x1 * q.w + y1 * q.x + z1 * q.y;
            o.w = w1 * q.w + x1 * q.z - y1 * q.y + z1 * q.x;
            *reinterpret_cast<float4*>(out_rot + 4ll * i) = o;
        }
    }
}

__global__ void __launch_bounds__(256) pose_bwd_kernel(const int P, const float* __restrict__ pose, const float* __restrict__ xyz,
                                                       const float* __restrict__ rot, const float* __restrict__ g_xyz,
                                                       const float* __restrict__ g_rot, float* __restrict__ d_xyz,
                                                       float* __restrict__ d_rot, float* __restrict__ partials) {
    const Rot3 R = rotation_of(pose);
    const float w1 = pose[0], x1 = pose[1], y1 = pose[2], z1 = pose[3];
    float acc[kPoseTerms];
#pragma unroll
    for (int k = 0; k < kPoseTerms; ++k) acc[k] = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        if (g_xyz) {
            const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
            const float gx = g_xyz[3 * i], gy = g_xyz[3 * i + 1], gz = g_xyz[3 * i + 2];
            d_xyz[3 * i + 0] = R.m[0][0] * gx + R.m[1][0] * gy + R.m[2][0] * gz;  // R^T g
            d_xyz[3 * i + 1] = R.m[0][1] * gx + R.m[1][1] * gy + R.m[2][1] * gz;
            d_xyz[3 * i + 2] = R.m[0][2] * gx + R.m[1][2] * gy + R.m[2][2] * gz;
            acc[0] += gx * px; acc[1] += gx * py; acc[2] += gx * pz;             // dL/dR = sum g x^T
            acc[3] += gy * px; acc[4] += gy * py; acc[5] += gy * pz;
            acc[6] += gz * px; acc[7] += gz * py; acc[8] += gz * pz;
            acc[9] += gx; acc[10] += gy; acc[11] += gz;                           // dL/dT
        } else {
            d_xyz[3 * i] = d_xyz[3 * i + 1] = d_xyz[3 * i + 2] = 0.f;
        }
        if (d_rot) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g_rot) {
                const float4 q = *reinterpret_cast<const float4*>(rot + 4ll * i);
                const float4 g = *reinterpret_cast<const float4*>(g_rot + 4ll * i);  // (g_w, g_x, g_y, g_z)
                o.x = g.x * w1 + g.y * x1 + g.z * y1 + g.w * z1;
                o.y = -g.x * x1 + g.y * w1 + g.z * z1 - g.w * y1;
                o.z = -g.x * y1 - g.y * z1 + g.z * w1 + g.w * x1;
                o.w = -g.x * z1 + g.y * y1 - g.z * x1 + g.w * w1;
                acc[12] += g.x * q.x + g.y * q.y + g.z * q.z + g.w * q.w;   // d/dw1
                acc[13] += -g.x * q.y + g.y * q.x - g.z * q.w + g.w * q.z;  // d/dx1
                acc[14] += -g.x * q.z + g.y * q.w + g.z * q.x - g.w * q.y;  // d/dy1
                acc[15] += -g.x * q.w - g.y * q.z + g.z * q.y + g.w * q.x;  // d/dz1
            }
            *reinterpret_cast<float4*>(d_rot + 4ll * i) = o;
        }
    }
    __shared__ float s_part[8][kPoseTerms];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kPoseTerms; ++k) {
        float v = acc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) s_part[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < kPoseTerms) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += s_part[w][threadIdx.x];
        partials[blockIdx.x * kPoseTerms + threadIdx.x] = v;
    }
}

// one block of 16 warps: warp k adds term k of the partial rows (every lane a strided subset in double, then a fixed shuffle
// tree: deterministic), thread 0 applies the chain rule through R(q / |q|)
__global__ void __launch_bounds__(512) pose_bwd_finish_kernel(const int nblocks, const float* __restrict__ pose,
                                                              const float* __restrict__ partials, float* __restrict__ d_pose,
                                                              const int accumulate) {
    __shared__ double s[kPoseTerms];
    const int lane = threadIdx.x & 31, term = threadIdx.x >> 5;
    double v = 0.0;
    for (int b = lane; b < nblocks; b += 32) v += (double)partials[(size_t)b * kPoseTerms + term];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) s[term] = v;
    __syncthreads();
    if (threadIdx.x != 0) return;
    const double a = pose[0], b = pose[1], c = pose[2], d = pose[3];
    const double n = sqrt(a * a + b * b + c * c + d * d);
    const double r = a / n, x = b / n, y = c / n, z = d / n;
    const double D00 = s[0], D01 = s[1], D02 = s[2], D10 = s[3], D11 = s[4], D12 = s[5], D20 = s[6], D21 = s[7], D22 = s[8];
    const double gr = 2.0 * (-z * D01 + y * D02 + z * D10 - x * D12 - y * D20 + x * D21);
    const double gx = 2.0 * (y * D01 + z * D02 + y * D10 - 2.0 * x * D11 - r * D12 + z * D20 + r * D21 - 2.0 * x * D22);
    const double gy = 2.0 * (-2.0 * y * D00 + x * D01 + r * D02 + x * D10 + z * D12 - r * D20 + z * D21 - 2.0 * y * D22);
    const double gz = 2.0 * (-2.0 * z * D00 - r * D01 + x * D02 + r * D10 - 2.0 * z * D11 + y * D12 + x * D20 + y * D21);
    const double dot = gr * r + gx * x + gy * y + gz * z;  // d(q/|q|)/dq = (I - qn qn^T) / |q|
    const double out[7] = {(gr - dot * r) / n + s[12], (gx - dot * x) / n + s[13], (gy - dot * y) / n + s[14],
                           (gz - dot * z) / n + s[15], s[9], s[10], s[11]};
    for (int k = 0; k < 7; ++k) d_pose[k] = (float)(out[k] + (accumulate ? (double)d_pose[k] : 0.0));
}

// ---- masked L1 ------------------------------------------------------------------------------------------------------
// partial[b] = sum over this block's elements of |a*m - b*m| ; optional sign map for the backward pass is recomputed there.
__global__ void __launch_bounds__(256) masked_l1_fwd_kernel(const long long n, const long long hw, const int mask_channels,
                                                            const float* __restrict__ a, const float* __restrict__ b,
                                                            const float* __restrict__ mask, float* __restrict__ partial) {
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float m = mask ? mask[mask_channels == 1 ? i % hw : i] : 1.0f;
        acc += fabsf(__fadd_rn(__fmul_rn(a[i], m), -__fmul_rn(b[i], m)));
    }
    __shared__ float s_part[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += s_part[w];
        partial[blockIdx.x] = v;
    }
}

// dL/da = upstream / n * sign(a*m - b*m) * m      (torch: abs' = sign, with sign(0) = 0)
__global__ void __launch_bounds__(256) masked_l1_bwd_kernel(const long long n, const long long hw, const int mask_channels,
                                                            const float* __restrict__ a, const float* __restrict__ b,
                                                            const float* __restrict__ mask, const float* __restrict__ upstream,
                                                            float* __restrict__ d_a) {
    const float scale = (upstream ? *upstream : 1.0f) / (float)n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float m = mask ? mask[mask_channels == 1 ? i % hw : i] : 1.0f;
        const float d = __fadd_rn(__fmul_rn(a[i], m), -__fmul_rn(b[i], m));
        const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        d_a[i] = scale * sgn * m;
    }
}

}  // namespace

int launch_pose_finish(int rows, const float* pose, const float* partials, float* d_pose, int accumulate, cudaStream_t stream) {
    pose_bwd_finish_kernel<<<1, 512, 0, stream>>>(rows, pose, partials, d_pose, accumulate);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

}  // namespace lsx

using namespace lsx;

extern "C" int32_t lsx_pose_num_partials(void) { return kPoseBlocks * kPoseTerms; }

extern "C" int lsx_pose_transform_forward(int32_t P, const float* pose, const float* xyz, const float* rotation_raw,
                                          float* out_means3D, float* out_rotations, void* stream_) {
    if (P < 0 || (P > 0 && (!pose || !xyz || !out_means3D || ((rotation_raw == nullptr) != (out_rotations == nullptr))))) {
        set_error("lsx_pose_transform_forward: bad arguments");
        return -1;
    }
    if (P == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(rotation_raw) | reinterpret_cast<uintptr_t>(out_rotations)) & 15u) {
        set_error("lsx_pose_transform_forward: the (P, 4) rotation arrays must be 16-byte aligned");
        return -1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int blocks = (P + 255) / 256 < 148 * 8 ? (P + 255) / 256 : 148 * 8;
    pose_fwd_kernel<<<blocks, 256, 0, stream>>>(P, pose, xyz, rotation_raw, out_means3D, out_rotations);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

extern "C" int lsx_pose_transform_backward(int32_t P, const float* pose, const float* xyz, const float* rotation_raw,
                                           const float* dL_dmeans3D, const float* dL_drotations, float* dL_dxyz,
                                           float* dL_drotation_raw, float* dL_dpose, float* partials, void* stream_) {
    if (P < 0 || !pose || !dL_dpose || !partials || (P > 0 && (!xyz || !dL_dxyz || (dL_drotations && !rotation_raw)))) {
        set_error("lsx_pose_transform_backward: bad arguments");
        return -1;
    }
    if ((reinterpret_cast<uintptr_t>(rotation_raw) | reinterpret_cast<uintptr_t>(dL_drotations) |
         reinterpret_cast<uintptr_t>(dL_drotation_raw)) & 15u) {
        set_error("lsx_pose_transform_backward: the (P, 4) rotation arrays must be 16-byte aligned");
        return -1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int blocks = (P + 255) / 256 < kPoseBlocks ? (P + 255) / 256 : kPoseBlocks;
    if (blocks < 1) blocks = 1;
    pose_bwd_kernel<<<blocks, 256, 0, stream>>>(P, pose, xyz, rotation_raw, dL_dmeans3D, dL_drotations, dL_dxyz,
                                                dL_drotation_raw, partials);
    LSX_KERNEL_OK(stream, false);
    return launch_pose_finish(blocks, pose, partials, dL_dpose, 0, stream);
}

extern "C" int32_t lsx_masked_l1_num_blocks(int64_t n) {
    const long long want = (n + 1023) / 1024;
    return (int32_t)(want < 1 ? 1 : (want < kL1Blocks ? want : kL1Blocks));
}

extern "C" int lsx_masked_l1_forward(int32_t C, int32_t H, int32_t W, int32_t mask_channels, const float* a, const float* b,
                                     const float* mask, float* partial, void* stream_) {
    if (C <= 0 || H <= 0 || W <= 0 || !a || !b || !partial || (mask && mask_channels != 1 && mask_channels != C)) {
        set_error("lsx_masked_l1_forward: bad arguments (mask must have 1 or C channels)");
        return -1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long hw = (long long)H * W, n = hw * C;
    masked_l1_fwd_kernel<<<lsx_masked_l1_num_blocks(n), 256, 0, stream>>>(n, hw, mask_channels, a, b, mask, partial);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

extern "C" int lsx_masked_l1_backward(int32_t C, int32_t H, int32_t W, int32_t mask_channels, const float* a, const float* b,
                                      const float* mask, const float* upstream, float* dL_da, void* stream_) {
    if (C <= 0 || H <= 0 || W <= 0 || !a || !b || !dL_da || (mask && mask_channels != 1 && mask_channels != C)) {
        set_error("lsx_masked_l1_backward: bad arguments (mask must have 1 or C channels)");
        return -1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long hw = (long long)H * W, n = hw * C;
    masked_l1_bwd_kernel<<<lsx_masked_l1_num_blocks(n), 256, 0, stream>>>(n, hw, mask_channels, a, b, mask, upstream, dL_da);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

// ---- row pack / unpack for the on-disk formats ------------------------------------------------
// GaussianModel.save_ply / load_ply (field_construction/scene/gaussian_model.py:415-441,448-504) move between the optimiser's
// per-group tensors and ONE row of `ncols` floats per Gaussian (x y z nx ny nz f_dc_* f_rest_* opacity scale_* rot_*
// language_feature_* instance_feature_*; SH coefficients transposed to channel-major).  The reference does it on the host
// (numpy concatenate + a Python tuple per row); here it is one gather / scatter over the arena on the device:
//     rows[r * ncols + c] = arena[col_begin[c] + r * col_stride[c]]      (col_begin[c] < 0: constant 0, the normals)
namespace lsx {
namespace {
__global__ void __launch_bounds__(256) rows_pack_kernel(const long long n, const int ncols, const long long* __restrict__ col_begin,
                                                        const int* __restrict__ col_stride, const float* __restrict__ arena,
                                                        float* __restrict__ rows) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / ncols;
        const int c = (int)(e - r * ncols);
        const long long b = col_begin[c];
        rows[e] = b < 0 ? 0.f : arena[b + r * col_stride[c]];
    }
}
__global__ void __launch_bounds__(256) rows_unpack_kernel(const long long n, const int ncols, const long long* __restrict__ col_begin,
                                                          const int* __restrict__ col_stride, const float* __restrict__ rows,
                                                          float* __restrict__ arena) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / ncols;
        const int c = (int)(e - r * ncols);
        const long long b = col_begin[c];
        if (b >= 0) arena[b + r * col_stride[c]] = rows[e];
    }
}
}  // namespace
}  // namespace lsx

extern "C" int lsx_rows_pack(int64_t P, int32_t ncols, const int64_t* col_begin, const int32_t* col_stride, const float* arena,
                             float* rows, void* stream_) {
    if (P < 0 || ncols <= 0 || !col_begin || !col_stride || (P > 0 && (!arena || !rows))) {
        set_error("lsx_rows_pack: bad arguments");
        return -1;
    }
    if (P == 0) return 0;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long n = P * ncols, want = (n + 255) / 256;
    rows_pack_kernel<<<(int)(want < 148 * 32 ? want : 148 * 32), 256, 0, stream>>>(n, ncols, (const long long*)col_begin, col_stride,
                                                                                arena, rows);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

extern "C" int lsx_rows_unpack(int64_t P, int32_t ncols, const int64_t* col_begin, const int32_t* col_stride, const float* rows,
                               float* arena, void* stream_) {
    if (P < 0 || ncols <= 0 || !col_begin || !col_stride || (P > 0 && (!arena || !rows))) {
        set_error("lsx_rows_unpack: bad arguments");
        return -1;
    }
    if (P == 0) return 0;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long n = P * ncols, want = (n + 255) / 256;
    rows_unpack_kernel<<<(int)(want < 148 * 32 ? want : 148 * 32), 256, 0, stream>>>(n, ncols, (const long long*)col_begin,
                                                                                  col_stride, rows, arena);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

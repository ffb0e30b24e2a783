// head_math.cuh — per-Gaussian math of LangScene-X's render wrapper, shared by the stand-alone wrapper kernels
// (gaussian_head.cu, pose.cu) and by the preprocess kernels' fused "raw parameter" mode (preprocess.cu), so that the fused and
// the unfused routes execute the same expressions.
//
// Reference behaviour restated:
//   activations   scales = exp(_scaling), rotations = F.normalize(_rotation), opacity = sigmoid(_opacity)
//                 (field_construction/scene/gaussian_model.py:53-61,193-213)
//   plane normal  column argmin(scales) of quaternion_to_matrix(rotations) (pytorch3d.transforms, dependency absent from the
//                 reference tree: R = I + two_s * [...], two_s = 2 / |q|^2, q = (r, i, j, k)), flipped towards the camera
//                 (gaussian_model.py:225-236 get_smallest_axis / get_normal)
//   all_map       [normal @ W2V[:3,:3], 1, |<local normal, xyz @ W2V[:3,:3] + W2V[3,:3]>|]
//                 (field_construction/gaussian_renderer/__init__.py:188-196)
//   pose          means3D = R(q / |q|) xyz + T,  rotations = quadmultiply(pose[:4], _rotation)  (Hamilton product with the RAW
//                 pose quaternion)   (gaussian_renderer/__init__.py:79-87, utils/pose_utils.py:13-107)
#pragma once
#include "common.cuh"

namespace lsx {

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


// gradients w.r.t. the raw parameters of one Gaussian from the gradients the rasterizer returns for its activated inputs
struct HeadGrads {
    float xyz[3];    // includes g_means3D (the rasterizer's own position gradient)
    float sraw[3];
    float4 qraw;     // w.r.t. the quaternion handed to head_frame (raw, or pose-multiplied raw)
    float oraw;
};

__device__ __forceinline__ HeadGrads head_backward_row(const HeadCam& c, const HeadFrame& f, const float4 qraw,
                                                       const float (&g_scales)[3], const float (&g_rotations)[4],
                                                       const float g_opacity, const float (&gam)[5],
                                                       const float (&g_means3D)[3]) {
    HeadGrads o;
    // activations
#pragma unroll
    for (int a = 0; a < 3; ++a) o.sraw[a] = g_scales[a] * f.sc[a];
    o.oraw = g_opacity * f.sig * (1.0f - f.sig);

    // all_map -> local normal, camera-space point
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
        o.xyz[r] = gx + g_means3D[r];
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
        float acc = g_rotations[m];
#pragma unroll
        for (int e = 0; e < 3; ++e) acc += g_col[e] * (s2 * dfe[e][m] - s2 * s2 * f.q[m] * fe[e]);
        gq[m] = acc;
    }
    // F.normalize backward
    const float norm_raw = sqrtf(qraw.x * qraw.x + qraw.y * qraw.y + qraw.z * qraw.z + qraw.w * qraw.w);
    if (norm_raw > 1.0e-12f) {
        const float qg = f.q[0] * gq[0] + f.q[1] * gq[1] + f.q[2] * gq[2] + f.q[3] * gq[3];
        o.qraw = make_float4((gq[0] - f.q[0] * qg) / f.nq, (gq[1] - f.q[1] * qg) / f.nq, (gq[2] - f.q[2] * qg) / f.nq,
                             (gq[3] - f.q[3] * qg) / f.nq);
    } else {
        o.qraw = make_float4(gq[0] / f.nq, gq[1] / f.nq, gq[2] / f.nq, gq[3] / f.nq);  // q = q_raw / eps
    }
    return o;
}

// ---- camera-pose transform -------------------------------------------------------------------------------------------
struct Rot3 {
    float m[3][3];
};

__device__ __forceinline__ Rot3 rotation_of(const float* __restrict__ pose) {
    const float a = pose[0], b = pose[1], c = pose[2], d = pose[3];
    const float n = sqrtf(a * a + b * b + c * c + d * d);
    const float r = a / n, x = b / n, y = c / n, z = d / n;
    Rot3 R;
    R.m[0][0] = 1.f - 2.f * (y * y + z * z);
    R.m[0][1] = 2.f * (x * y - r * z);
    R.m[0][2] = 2.f * (x * z + r * y);
    R.m[1][0] = 2.f * (x * y + r * z);
    R.m[1][1] = 1.f - 2.f * (x * x + z * z);
    R.m[1][2] = 2.f * (y * z - r * x);
    R.m[2][0] = 2.f * (x * z - r * y);
    R.m[2][1] = 2.f * (y * z + r * x);
    R.m[2][2] = 1.f - 2.f * (x * x + y * y);
    return R;
}

constexpr int kPoseTerms = 16;  // per-row contributions to the pose gradient: dL/dR (9), dL/dT (3), quaternion product (4)

// means3D = R p + T
__device__ __forceinline__ void pose_apply_point(const Rot3& R, const float* __restrict__ pose, const float px, const float py,
                                                 const float pz, float (&out)[3]) {
    out[0] = R.m[0][0] * px + R.m[0][1] * py + R.m[0][2] * pz + pose[4];
    out[1] = R.m[1][0] * px + R.m[1][1] * py + R.m[1][2] * pz + pose[5];
    out[2] = R.m[2][0] * px + R.m[2][1] * py + R.m[2][2] * pz + pose[6];
}
// rotations = pose_q (x) q   (Hamilton product with the RAW pose quaternion)
__device__ __forceinline__ float4 pose_apply_quat(const float* __restrict__ pose, const float4 q) {
    const float w1 = pose[0], x1 = pose[1], y1 = pose[2], z1 = pose[3];
    float4 o;
    o.x = w1 * q.x - x1 * q.y - y1 * q.z - z1 * q.w;
    o.y = w1 * q.y + x1 * q.x + y1 * q.w - z1 * q.z;
    o.z = w1 * q.z - x1 * q.w + y1 * q.x + z1 * q.y;
    o.w = w1 * q.w + x1 * q.z - y1 * q.y + z1 * q.x;
    return o;
}
// backward of both for one row: d_xyz = R^T g, d_q = conj-product, and the row's 16 pose-gradient terms added to acc
__device__ __forceinline__ void pose_backward_point(const Rot3& R, const float px, const float py, const float pz, const float gx,
                                                    const float gy, const float gz, float (&d_xyz)[3], float (&acc)[kPoseTerms]) {
    d_xyz[0] = R.m[0][0] * gx + R.m[1][0] * gy + R.m[2][0] * gz;  // R^T g
    d_xyz[1] = R.m[0][1] * gx + R.m[1][1] * gy + R.m[2][1] * gz;
    d_xyz[2] = R.m[0][2] * gx + R.m[1][2] * gy + R.m[2][2] * gz;
    acc[0] += gx * px; acc[1] += gx * py; acc[2] += gx * pz;             // dL/dR = sum g x^T
    acc[3] += gy * px; acc[4] += gy * py; acc[5] += gy * pz;
    acc[6] += gz * px; acc[7] += gz * py; acc[8] += gz * pz;
    acc[9] += gx; acc[10] += gy; acc[11] += gz;                           // dL/dT
}
__device__ __forceinline__ float4 pose_backward_quat(const float* __restrict__ pose, const float4 q, const float4 g,
                                                     float (&acc)[kPoseTerms]) {
    const float w1 = pose[0], x1 = pose[1], y1 = pose[2], z1 = pose[3];
    float4 o;
    o.x = g.x * w1 + g.y * x1 + g.z * y1 + g.w * z1;
    o.y = -g.x * x1 + g.y * w1 + g.z * z1 - g.w * y1;
    o.z = -g.x * y1 - g.y * z1 + g.z * w1 + g.w * x1;
    o.w = -g.x * z1 + g.y * y1 - g.z * x1 + g.w * w1;
    acc[12] += g.x * q.x + g.y * q.y + g.z * q.z + g.w * q.w;   // d/dw1
    acc[13] += -g.x * q.y + g.y * q.x - g.z * q.w + g.w * q.z;  // d/dx1
    acc[14] += -g.x * q.z + g.y * q.w + g.z * q.x - g.w * q.y;  // d/dy1
    acc[15] += -g.x * q.w - g.y * q.z + g.z * q.y + g.w * q.x;  // d/dz1
    return o;
}

}  // namespace lsx

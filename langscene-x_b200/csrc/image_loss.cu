// image_loss.cu — fused L1 + SSIM image loss, forward and backward (the loss side of the loop).
//
// Reference behaviour restated (field_construction/utils/loss_utils.py):
//   l1_loss :20-21   mean |x - y|
//   ssim    :37-75   11x11 Gaussian window (sigma 1.5, :32-41), zero padding 5, depthwise per channel:
//                    mu1 = w*x, mu2 = w*y, s11 = w*(x x) - mu1^2, s22 = w*(y y) - mu2^2, s12 = w*(x y) - mu1 mu2,
//                    ssim_map = (2 mu1 mu2 + C1)(2 s12 + C2) / ((mu1^2 + mu2^2 + C1)(s11 + s22 + C2)), C1 = 1e-4, C2 = 9e-4,
//                    result = mean over all C*H*W pixels.
//   call site (field_construction/gaussian_field.py:238-246):  (1 - lambda) * l1 + lambda * (1 - ssim).
// The reference runs 5 depthwise 11x11 convolutions, ~15 element-wise kernels and their autograd graph (5 more transposed
// convolutions) per image.  Here: ONE kernel forward (separable 11-tap window on a 26x26 shared-memory tile, five moments
// at once, per-block partial sums of the SSIM map and of |x - y|, three derivative maps written for the backward pass)
// and ONE kernel backward (the same separable window over the three derivative maps, combined with x and y).
#include "../../include/lsx_rasterizer.h"
#include "kernels.cuh"

namespace lsx {
namespace {

constexpr int kWin = 11, kHalo = 5, kTileO = 16, kTileI = kTileO + 2 * kHalo;  // 26

struct Window {
    float g[kWin];
};

__global__ void __launch_bounds__(256) ssim_l1_fwd_kernel(const int C, const int H, const int W, const Window win,
                                                          const float* __restrict__ img1, const float* __restrict__ img2,
                                                          float* __restrict__ dmaps,     // 3 x C*H*W: d/dmu1, d/dE[x^2], d/dE[xy]
                                                          float* __restrict__ partial) {  // 2 x #blocks: ssim sum, l1 sum
    __shared__ float sx[kTileI][kTileI + 1], sy[kTileI][kTileI + 1];
    __shared__ float h[5][kTileI][kTileO + 1];
    __shared__ float red[2][8];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int x0 = blockIdx.x * kTileO, y0 = blockIdx.y * kTileO, c = blockIdx.z;
    const size_t plane = (size_t)H * W;
    const float* p1 = img1 + c * plane;
    const float* p2 = img2 + c * plane;
    for (int i = threadIdx.x; i < kTileI * kTileI; i += 256) {
        const int ly = i / kTileI, lx = i - ly * kTileI;
        const int gy = y0 + ly - kHalo, gx = x0 + lx - kHalo;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        sx[ly][lx] = in ? p1[(size_t)gy * W + gx] : 0.f;
        sy[ly][lx] = in ? p2[(size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    // horizontal pass: 26 rows x 16 columns, five moments
    for (int i = threadIdx.x; i < kTileI * kTileO; i += 256) {
        const int ly = i / kTileO, lx = i - ly * kTileO;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float wv = win.g[k], xv = sx[ly][lx + k], yv = sy[ly][lx + k];
            a0 += wv * xv;
            a1 += wv * yv;
            a2 += wv * xv * xv;
            a3 += wv * yv * yv;
            a4 += wv * xv * yv;
        }
        h[0][ly][lx] = a0; h[1][ly][lx] = a1; h[2][ly][lx] = a2; h[3][ly][lx] = a3; h[4][ly][lx] = a4;
    }
    __syncthreads();
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
        const float wv = win.g[k];
#pragma unroll
        for (int q = 0; q < 5; ++q) m[q] += wv * h[q][ty + k][tx];
    }
    const int gx = x0 + tx, gy = y0 + ty;
    float ssim_v = 0.f, l1_v = 0.f;
    if (gx < W && gy < H) {
        const float mu1 = m[0], mu2 = m[1];
        const float s11 = m[2] - mu1 * mu1, s22 = m[3] - mu2 * mu2, s12 = m[4] - mu1 * mu2;
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float A = 2.f * mu1 * mu2 + C1, B = 2.f * s12 + C2, Cc = mu1 * mu1 + mu2 * mu2 + C1, D = s11 + s22 + C2;
        const float inv_cd = 1.f / (Cc * D);
        ssim_v = A * B * inv_cd;
        // partial derivatives of the SSIM value at this pixel w.r.t. its own mu1, E[x^2], E[xy]
        const float d_mu1 = 2.f * mu2 * (B - A) * inv_cd - 2.f * mu1 * ssim_v / Cc + 2.f * mu1 * ssim_v / D;
        const float d_ex2 = -ssim_v / D;
        const float d_exy = 2.f * A * inv_cd;
        const size_t o = c * plane + (size_t)gy * W + gx, n = (size_t)C * plane;
        dmaps[o] = d_mu1;
        dmaps[n + o] = d_ex2;
        dmaps[2 * n + o] = d_exy;
        l1_v = fabsf(sx[ty + kHalo][tx + kHalo] - sy[ty + kHalo][tx + kHalo]);
    }
    // block partial sums (deterministic: fixed tree)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ssim_v += __shfl_xor_sync(0xffffffffu, ssim_v, o);
        l1_v += __shfl_xor_sync(0xffffffffu, l1_v, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = ssim_v;
        red[1][threadIdx.x >> 5] = l1_v;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        const size_t nblk = (size_t)gridDim.x * gridDim.y * gridDim.z;
        const size_t b = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        partial[threadIdx.x * nblk + b] = s;
    }
}

// dL/dimg1 = k_ssim * [ w*(d_mu1) + 2 x w*(d_ex2) + y w*(d_exy) ] + k_l1 * sign(x - y)
__global__ void __launch_bounds__(256) ssim_l1_bwd_kernel(const int C, const int H, const int W, const Window win,
                                                          const float* __restrict__ img1, const float* __restrict__ img2,
                                                          const float* __restrict__ dmaps, const float* __restrict__ g_ssim,
                                                          const float* __restrict__ g_l1, const float inv_n,
                                                          float* __restrict__ g_img1) {
    // upstream gradients of the two means are read from device memory: no host round trip in the autograd backward
    const float k_ssim = g_ssim ? __ldg(g_ssim) * inv_n : 0.f, k_l1 = g_l1 ? __ldg(g_l1) * inv_n : 0.f;
    __shared__ float s[3][kTileI][kTileI + 1];
    __shared__ float h[3][kTileI][kTileO + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int x0 = blockIdx.x * kTileO, y0 = blockIdx.y * kTileO, c = blockIdx.z;
    const size_t plane = (size_t)H * W, n = (size_t)C * plane;
    for (int i = threadIdx.x; i < kTileI * kTileI; i += 256) {
        const int ly = i / kTileI, lx = i - ly * kTileI;
        const int gy = y0 + ly - kHalo, gx = x0 + lx - kHalo;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const size_t o = c * plane + (size_t)gy * W + gx;
#pragma unroll
        for (int q = 0; q < 3; ++q) s[q][ly][lx] = in ? dmaps[q * n + o] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kTileI * kTileO; i += 256) {
        const int ly = i / kTileO, lx = i - ly * kTileO;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float wv = win.g[k];
            a0 += wv * s[0][ly][lx + k];
            a1 += wv * s[1][ly][lx + k];
            a2 += wv * s[2][ly][lx + k];
        }
        h[0][ly][lx] = a0; h[1][ly][lx] = a1; h[2][ly][lx] = a2;
    }
    __syncthreads();
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx >= W || gy >= H) return;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
        const float wv = win.g[k];
        m0 += wv * h[0][ty + k][tx];
        m1 += wv * h[1][ty + k][tx];
        m2 += wv * h[2][ty + k][tx];
    }
    const size_t o = c * plane + (size_t)gy * W + gx;
    const float x = img1[o], y = img2[o];
    const float d = x - y;
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    g_img1[o] = k_ssim * (m0 + 2.f * x * m1 + y * m2) + k_l1 * sgn;
}

Window make_window() {
    // exactly the reference's construction: float32 tensor of exp(...) values, divided by its float32 sum
    Window w;
    float sum = 0.f;
    for (int i = 0; i < kWin; ++i) {
        w.g[i] = (float)exp(-(double)((i - kWin / 2) * (i - kWin / 2)) / (2.0 * 1.5 * 1.5));
        sum += w.g[i];
    }
    for (int i = 0; i < kWin; ++i) w.g[i] = w.g[i] / sum;
    return w;
}

}  // namespace
}  // namespace lsx

using namespace lsx;

extern "C" int64_t lsx_image_loss_num_blocks(int32_t C, int32_t H, int32_t W) {
    return (int64_t)ceil_div(W, kTileO) * ceil_div(H, kTileO) * C;
}

extern "C" int lsx_image_loss_forward(int32_t C, int32_t H, int32_t W, const float* img1, const float* img2, float* dmaps,
                                      float* partial, void* stream_) {
    if (C <= 0 || H <= 0 || W <= 0 || C > 65535 || !img1 || !img2 || !dmaps || !partial) {
        set_error("lsx_image_loss_forward: bad arguments");
        return -1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const dim3 grid((unsigned)ceil_div(W, kTileO), (unsigned)ceil_div(H, kTileO), (unsigned)C);
    ssim_l1_fwd_kernel<<<grid, 256, 0, stream>>>(C, H, W, make_window(), img1, img2, dmaps, partial);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

extern "C" int lsx_image_loss_backward(int32_t C, int32_t H, int32_t W, const float* img1, const float* img2, const float* dmaps,
                                       const float* dL_dssim_mean, const float* dL_dl1_mean, float* dL_dimg1, void* stream_) {
    if (C <= 0 || H <= 0 || W <= 0 || C > 65535 || !img1 || !img2 || !dmaps || !dL_dimg1) {
        set_error("lsx_image_loss_backward: bad arguments");
        return -1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const dim3 grid((unsigned)ceil_div(W, kTileO), (unsigned)ceil_div(H, kTileO), (unsigned)C);
    ssim_l1_bwd_kernel<<<grid, 256, 0, stream>>>(C, H, W, make_window(), img1, img2, dmaps, dL_dssim_mean, dL_dl1_mean,
                                                 1.0f / ((float)C * (float)H * (float)W), dL_dimg1);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

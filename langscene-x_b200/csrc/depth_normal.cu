// depth_normal.cu — plane depth -> per-pixel normal ("depth normal"), forward and backward.
//
// First of the "next" rows of SURVEY.md 8(f): the step directly behind the rasterizer in LangScene-X's render wrapper,
//     depth_normal = render_normal(cam, plane_depth) * rendered_alpha.detach()
// (field_construction/gaussian_renderer/__init__.py:28-40,233-235).  Reference behaviour restated
// (field_construction/utils/graphics_utils.py:16-75, offset == None branch):
//     P(y,x)   = depth(y,x) * ((x - cx)/fx, (y - cy)/fy, 1)              back-projection with K = [[fx,0,cx],[0,fy,cy],[0,0,1]]
//     n(y,x)   = normalize( (P(y,x+1) - P(y,x-1)) x (P(y-1,x) - P(y+1,x)) )     for interior pixels, eps = 1e-12
//     n        = 0 on the one-pixel border (zero padding)
// The reference runs ~15 torch kernels (meshgrid, stack, 3x3 inverse, matmul, 4 slices, cross, normalize, pad, permute)
// plus their autograd graph over the W*H image; here it is one HBM-bound stencil kernel each way
// (forward: 4 B read (+4 alpha) and 12 B written per pixel; backward: 13-point stencil of depth, 12 B gradient read,
// 4 B written), with the alpha multiplication of the call site fused in.
#include "../../include/lsx_rasterizer.h"
#include "kernels.cuh"

namespace lsx {
namespace {

struct Intr {
    float fx, fy, cx, cy;
};

__device__ __forceinline__ float3 ray_of(const Intr k, const int x, const int y) {
    return make_float3(((float)x - k.cx) / k.fx, ((float)y - k.cy) / k.fy, 1.0f);
}
__device__ __forceinline__ float3 cross3(const float3 a, const float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float dot3(const float3 a, const float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// the two edge vectors of the interior pixel (x, y)
__device__ __forceinline__ void edges_at(const float* __restrict__ depth, const int W, const Intr k, const int x, const int y,
                                         float3& a, float3& b) {
    const float zr = depth[y * W + x + 1], zl = depth[y * W + x - 1];
    const float zt = depth[(y - 1) * W + x], zb = depth[(y + 1) * W + x];
    const float3 rr = ray_of(k, x + 1, y), rl = ray_of(k, x - 1, y), rt = ray_of(k, x, y - 1), rb = ray_of(k, x, y + 1);
    a = make_float3(zr * rr.x - zl * rl.x, zr * rr.y - zl * rl.y, zr - zl);  // left -> right
    b = make_float3(zt * rt.x - zb * rb.x, zt * rt.y - zb * rb.y, zt - zb);  // bottom -> top
}

__global__ void __launch_bounds__(256) depth_normal_fwd_kernel(const int W, const int H, const Intr k,
                                                               const float* __restrict__ depth,
                                                               const float* __restrict__ alpha, float* __restrict__ out) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)W * H, pix = (size_t)y * W + x;
    float3 n = make_float3(0.f, 0.f, 0.f);
    if (x > 0 && x < W - 1 && y > 0 && y < H - 1) {
        float3 a, b;
        edges_at(depth, W, k, x, y, a, b);
        const float3 c = cross3(a, b);
        const float inv = 1.0f / fmaxf(sqrtf(dot3(c, c)), 1.0e-12f);
        const float s = alpha ? alpha[pix] * inv : inv;
        n = make_float3(c.x * s, c.y * s, c.z * s);
    }
    out[pix] = n.x;
    out[HW + pix] = n.y;
    out[2 * HW + pix] = n.z;
}

// d(loss)/d(edge vectors) of the interior pixel (x, y) given the upstream gradient of its normal
__device__ __forceinline__ void edge_grads_at(const float* __restrict__ depth, const float* __restrict__ alpha,
                                              const float* __restrict__ g_out, const int W, const size_t HW, const Intr k,
                                              const int x, const int y, float3& g_a, float3& g_b) {
    float3 a, b;
    edges_at(depth, W, k, x, y, a, b);
    const float3 c = cross3(a, b);
    const float len = sqrtf(dot3(c, c));
    const size_t pix = (size_t)y * W + x;
    const float s = alpha ? alpha[pix] : 1.0f;
    const float3 g = make_float3(g_out[pix] * s, g_out[HW + pix] * s, g_out[2 * HW + pix] * s);
    float3 g_c;
    if (len > 1.0e-12f) {
        const float inv = 1.0f / len;
        const float3 n = make_float3(c.x * inv, c.y * inv, c.z * inv);
        const float ng = dot3(n, g);
        g_c = make_float3((g.x - n.x * ng) * inv, (g.y - n.y * ng) * inv, (g.z - n.z * ng) * inv);
    } else {
        g_c = make_float3(g.x * 1.0e12f, g.y * 1.0e12f, g.z * 1.0e12f);  // n = c / eps
    }
    g_a = cross3(b, g_c);
    g_b = cross3(g_c, a);
}

__global__ void __launch_bounds__(256) depth_normal_bwd_kernel(const int W, const int H, const Intr k,
                                                               const float* __restrict__ depth,
                                                               const float* __restrict__ alpha,
                                                               const float* __restrict__ g_out, float* __restrict__ g_depth) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)W * H;
    const float3 r = ray_of(k, x, y);
    auto interior = [&](int xx, int yy) { return xx > 0 && xx < W - 1 && yy > 0 && yy < H - 1; };
    float gz = 0.f;
    float3 ga, gb;
    if (interior(x - 1, y)) {  // this pixel is the RIGHT point of (x-1, y)
        edge_grads_at(depth, alpha, g_out, W, HW, k, x - 1, y, ga, gb);
        gz += dot3(ga, r);
    }
    if (interior(x + 1, y)) {  // LEFT point of (x+1, y)
        edge_grads_at(depth, alpha, g_out, W, HW, k, x + 1, y, ga, gb);
        gz -= dot3(ga, r);
    }
    if (interior(x, y + 1)) {  // TOP point of (x, y+1)
        edge_grads_at(depth, alpha, g_out, W, HW, k, x, y + 1, ga, gb);
        gz += dot3(gb, r);
    }
    if (interior(x, y - 1)) {  // BOTTOM point of (x, y-1)
        edge_grads_at(depth, alpha, g_out, W, HW, k, x, y - 1, ga, gb);
        gz -= dot3(gb, r);
    }
    g_depth[(size_t)y * W + x] = gz;
}

}  // namespace
}  // namespace lsx

using namespace lsx;

extern "C" int lsx_depth_normal_forward(int32_t W, int32_t H, float fx, float fy, float cx, float cy, const float* depth,
                                        const float* alpha, float* out_normal, void* stream_) {
    if (W <= 0 || H <= 0 || !depth || !out_normal || !(fx != 0.f) || !(fy != 0.f)) {
        set_error("lsx_depth_normal_forward: bad arguments");
        return -1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, 8));
    depth_normal_fwd_kernel<<<grid, 256, 0, stream>>>(W, H, Intr{fx, fy, cx, cy}, depth, alpha, out_normal);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

extern "C" int lsx_depth_normal_backward(int32_t W, int32_t H, float fx, float fy, float cx, float cy, const float* depth,
                                         const float* alpha, const float* dL_dnormal, float* dL_ddepth, void* stream_) {
    if (W <= 0 || H <= 0 || !depth || !dL_dnormal || !dL_ddepth || !(fx != 0.f) || !(fy != 0.f)) {
        set_error("lsx_depth_normal_backward: bad arguments");
        return -1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, 8));
    depth_normal_bwd_kernel<<<grid, 256, 0, stream>>>(W, H, Intr{fx, fy, cx, cy}, depth, alpha, dL_dnormal, dL_ddepth);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

// arena_adam.cu — one-kernel Adam step over the flat parameter arena (optimiser half; 8(e)).
//
// Reference behaviour: torch.optim.Adam(param_groups, lr=0.0, eps=1e-15) with one learning rate per group
// (field_construction/scene/gaussian_model.py:313-328; default betas (0.9, 0.999), no weight decay, no amsgrad), stepped once
// per iteration (field_construction/gaussian_field.py:537-543).  torch's update, restated per element:
//     m <- b1 m + (1 - b1) g ;  v <- b2 v + (1 - b2) g^2
//     p <- p - (lr / (1 - b1^t)) * m / ( sqrt(v) / sqrt(1 - b2^t) + eps )
// The view-sharded trainer keeps parameters, gradients (after the all-reduce) and both moments as flat fp32 arenas with the
// same group layout, so the whole optimiser step is ONE streaming kernel: 16 B read + 12 B written per element (HBM-bound),
// instead of ~10 element-wise passes per parameter group.
#include "../../include/lsx_rasterizer.h"
#include "kernels.cuh"

namespace lsx {
namespace {

constexpr int kMaxGroups = LSX_ADAM_MAX_GROUPS;

struct AdamGroups {
    long long begin[kMaxGroups + 1];  // element offsets, ascending; group i = [begin[i], begin[i+1])
    float step_size[kMaxGroups];      // lr_i / (1 - b1^t)
    int n;
};

__global__ void __launch_bounds__(256) arena_adam_kernel(const long long n, const AdamGroups grp, const float b1, const float b2,
                                                         const float sqrt_bc2, const float eps, float* __restrict__ p,
                                                         const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        // group of this 4-element packet (arena groups are 64-element aligned, so a packet never straddles two)
        int gi = 0;
#pragma unroll 1
        while (gi + 1 < grp.n && i >= grp.begin[gi + 1]) ++gi;
        const float ss = (i >= grp.begin[0] && i < grp.begin[grp.n]) ? grp.step_size[gi] : 0.f;
        if (i + 3 < n) {
            const float4 gg = *reinterpret_cast<const float4*>(g + i);
            float4 mm = *reinterpret_cast<const float4*>(m + i);
            float4 vv = *reinterpret_cast<const float4*>(v + i);
            float4 pp = *reinterpret_cast<const float4*>(p + i);
            const float ga[4] = {gg.x, gg.y, gg.z, gg.w};
            float ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w}, pa[4] = {pp.x, pp.y, pp.z, pp.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ma[k] = ma[k] + (1.f - b1) * (ga[k] - ma[k]);            // exp_avg.lerp_(grad, 1 - beta1)
                va[k] = b2 * va[k] + (1.f - b2) * ga[k] * ga[k];         // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
                pa[k] -= ss * (ma[k] / (sqrtf(va[k]) / sqrt_bc2 + eps));  // addcdiv_(exp_avg, sqrt/bc2_sqrt + eps, -step_size)
            }
            *reinterpret_cast<float4*>(m + i) = make_float4(ma[0], ma[1], ma[2], ma[3]);
            *reinterpret_cast<float4*>(v + i) = make_float4(va[0], va[1], va[2], va[3]);
            *reinterpret_cast<float4*>(p + i) = make_float4(pa[0], pa[1], pa[2], pa[3]);
        } else {
            for (long long j = i; j < n; ++j) {
                const float gj = g[j];
                const float mj = m[j] + (1.f - b1) * (gj - m[j]);
                const float vj = b2 * v[j] + (1.f - b2) * gj * gj;
                m[j] = mj;
                v[j] = vj;
                p[j] -= ss * (mj / (sqrtf(vj) / sqrt_bc2 + eps));
            }
        }
    }
}

}  // namespace
}  // namespace lsx

using namespace lsx;

extern "C" int lsx_arena_adam_step(int64_t n, int32_t n_groups, const int64_t* group_begin_host, const float* group_lr_host,
                                   int32_t step, float beta1, float beta2, float eps, float* params, const float* grads,
                                   float* exp_avg, float* exp_avg_sq, void* stream_) {
    if (n < 0 || n_groups <= 0 || n_groups > kMaxGroups || !group_begin_host || !group_lr_host || step < 1 ||
        (n > 0 && (!params || !grads || !exp_avg || !exp_avg_sq))) {
        set_error("lsx_arena_adam_step: bad arguments (at most %d groups, step >= 1)", kMaxGroups);
        return -1;
    }
    if (n == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u) {
        set_error("lsx_arena_adam_step: arenas must be 16-byte aligned");
        return -1;
    }
    AdamGroups grp{};
    grp.n = n_groups;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    for (int i = 0; i <= n_groups; ++i) {
        grp.begin[i] = group_begin_host[i];
        if (i > 0 && (grp.begin[i] < grp.begin[i - 1] || (grp.begin[i - 1] & 3))) {
            set_error("lsx_arena_adam_step: group offsets must ascend and be multiples of 4");
            return -1;
        }
    }
    for (int i = 0; i < n_groups; ++i) grp.step_size[i] = (float)((double)group_lr_host[i] / bc1);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long packets = (n + 3) / 4;
    const int blocks = (int)(packets / 256 + 1 < (long long)sms * 16 ? packets / 256 + 1 : (long long)sms * 16);
    arena_adam_kernel<<<blocks, 256, 0, stream>>>(n, grp, beta1, beta2, (float)sqrt(bc2), eps, params, grads, exp_avg,
                                                  exp_avg_sq);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

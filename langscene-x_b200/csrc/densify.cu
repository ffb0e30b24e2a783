// densify.cu — densification / pruning on the flat arenas.
//
// Reference behaviour restated (field_construction/scene/gaussian_model.py unless noted):
//   add_densification_stats :720-724 and the max_radii2D update of field_construction/gaussian_field.py:521-523;
//   densify_and_prune :700-718 = densify_and_clone :664-698, densify_and_split :612-662 (N = 2), final prune :709-716;
//   densification_postfix / cat_tensors_to_optimizer :561-610 (new rows get zero Adam moments, statistics reset);
//   prune_points / _prune_optimizer :520-559; reset_opacity :443-446 with replace_tensor_to_optimizer :506-518;
//   build_rotation: field_construction/utils/general_utils.py:80-101.
// The reference grows every parameter tensor twice (torch.cat per group, per Adam moment), then filters each with a boolean
// mask twice, re-creating nn.Parameters: ~150 kernels and 4 full copies of parameters + moments.
//
// B200 design — "plan, then one gather"
//   Every decision is a function of the ORIGINAL P rows (rows appended by the clone step carry a zero padded gradient and
//   their parent's scale, so the split step never selects them; its quantile sees them as n_clone leading zeros).  So:
//     1. classify_kernel      one pass over statistics + scaling + opacity -> flag byte, g, g_abs, candidate counts
//     2. (host)               the reference's budget branches on three counts; the rare capped branches take a quantile:
//                             radix sort of the masked values (sort.cu) + the interpolation ATen's quantile uses
//     3. plan_flags / 5 scans / plan_scatter   -> row_map[dst] = (kind << 30 | src), noise_index[dst]; totals to the host
//     4. apply_kernel         ONE streaming gather per arena group: parameters, exp_avg, exp_avg_sq (moments zero for new
//                             rows), sampled positions xyz + R(q) (exp(s) * z) and child scales log(exp(s) / 1.6) inline.
//   HBM traffic: 32 B/Gaussian (classify) + ~60 B/Gaussian (plan) + 24 B per arena element (apply: read + write of
//   parameter and both moments) — one copy instead of four.  Row order equals the reference's: surviving originals,
//   surviving clones, surviving first children, surviving second children.
#include <cmath>
#include "../../include/lsx_rasterizer.h"
#include "kernels.cuh"

namespace lsx {
namespace {

enum : uint32_t { F_CLONE = 1, F_SPLIT = 2, F_ABS = 4, F_BIG = 8, F_PRUNE_SELF = 16, F_PRUNE_CHILD = 32, F_ABS_ELIG = 64 };
enum { C_CLONE = 0, C_SPLIT = 1, C_ABS = 2, C_RESEL = 3, C_KO = 4, C_KC = 5, C_KS = 6, C_NC = 7, C_NS = 8, C_COUNT = 16 };

__global__ void __launch_bounds__(256) stats_update_kernel(const int P, const float* __restrict__ g2d,
                                                           const float* __restrict__ g2d_abs, const int* __restrict__ radii,
                                                           const int* __restrict__ observe, float* __restrict__ accum,
                                                           float* __restrict__ accum_abs, float* __restrict__ denom,
                                                           float* __restrict__ max_radii) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;  // update_filter = radii > 0
    const float x = g2d[3 * i], y = g2d[3 * i + 1], xa = g2d_abs[3 * i], ya = g2d_abs[3 * i + 1];
    accum[i] += sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
    accum_abs[i] += sqrtf(__fadd_rn(__fmul_rn(xa, xa), __fmul_rn(ya, ya)));
    denom[i] += 1.0f;
    if (!observe || observe[i] > 0) max_radii[i] = fmaxf(max_radii[i], (float)r);
}

struct ClassifyParams {
    int P;
    const float* accum;
    const float* accum_abs;
    const float* denom;
    const float* max_radii;
    const float* scaling;  // (P,3) raw (log) scales
    const float* opacity;  // (P) raw (logit) opacity
    float max_grad, abs_max_grad, dense_extent, abs_radii_thr, min_opacity, world_extent;  // world_extent < 0: test off
    uint8_t* flags;
    float* g;
    float* ga;
    uint32_t* counts;
};

__device__ __forceinline__ float sigmoidf_ref(const float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256) classify_kernel(const ClassifyParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t f = 0;
    if (i < p.P) {
        const float d = p.denom[i];
        float g = p.accum[i] / d, ga = p.accum_abs[i] / d;
        if (g != g) g = 0.f;  // grads[grads.isnan()] = 0.0  (:703-704)
        if (ga != ga) ga = 0.f;
        const float s0 = expf(p.scaling[3 * i]), s1 = expf(p.scaling[3 * i + 1]), s2 = expf(p.scaling[3 * i + 2]);
        const float smax = fmaxf(s0, fmaxf(s1, s2));
        const bool big = smax > p.dense_extent;
        const bool hot = g >= p.max_grad;
        if (big) f |= F_BIG;
        if (hot && !big) f |= F_CLONE;
        if (hot && big) f |= F_SPLIT;
        const bool elig = big && !(hot && big) && p.max_radii[i] > p.abs_radii_thr;  // :632-635
        if (elig) f |= F_ABS_ELIG;
        if (elig && ga >= p.abs_max_grad) f |= F_ABS;
        const bool low = sigmoidf_ref(p.opacity[i]) < p.min_opacity;  // :709
        const bool ws = p.world_extent >= 0.f;
        if (low || (ws && smax > p.world_extent)) f |= F_PRUNE_SELF;  // :714 (big_points_vs is always false, see header)
        // children: scale = exp(log(exp(s) / (0.8 * 2)))  (:650) evaluated the way get_scaling does at the prune
        const float c0 = expf(logf(s0 / 1.6f)), c1 = expf(logf(s1 / 1.6f)), c2 = expf(logf(s2 / 1.6f));
        if (low || (ws && fmaxf(c0, fmaxf(c1, c2)) > p.world_extent)) f |= F_PRUNE_CHILD;
        p.flags[i] = (uint8_t)f;
        p.g[i] = g;
        p.ga[i] = ga;
    }
    const int nc = __syncthreads_count(f & F_CLONE), ns = __syncthreads_count(f & F_SPLIT), na = __syncthreads_count(f & F_ABS);
    if (threadIdx.x == 0) {
        if (nc) atomicAdd(p.counts + C_CLONE, (uint32_t)nc);
        if (ns) atomicAdd(p.counts + C_SPLIT, (uint32_t)ns);
        if (na) atomicAdd(p.counts + C_ABS, (uint32_t)na);
    }
}

// keys[i] = bits of (flags[i] & bit ? values[i] : 0): non-negative floats order like their bit patterns
__global__ void __launch_bounds__(256) masked_keys_kernel(const int P, const float* __restrict__ values,
                                                          const uint8_t* __restrict__ flags, const uint32_t bit,
                                                          uint32_t* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) keys[i] = (flags[i] & bit) ? __float_as_uint(values[i]) : 0u;
}

// ATen quantile_compute (linear interpolation) on [n_pad zeros | sorted]: values_below.lerp_(values_above, weights)
__global__ void quantile_pick_kernel(const uint32_t* __restrict__ sorted, const int n_pad, const int lo, const int hi,
                                     const float w, float* __restrict__ thr) {
    const float a = lo < n_pad ? 0.f : __uint_as_float(sorted[lo - n_pad]);
    const float b = hi < n_pad ? 0.f : __uint_as_float(sorted[hi - n_pad]);
    const float diff = __fadd_rn(b, -a);
    *thr = (w < 0.5f) ? __fadd_rn(a, __fmul_rn(w, diff)) : __fadd_rn(b, -__fmul_rn(diff, __fadd_rn(1.0f, -w)));
}

// flag `set_bit` <- (masked value > threshold); other bits in `clear_bits` are dropped (the capped split branch has no abs pass)
__global__ void __launch_bounds__(256) reselect_kernel(const int P, const float* __restrict__ values, const uint32_t mask_bit,
                                                       const float* __restrict__ thr, const uint32_t set_bit,
                                                       const uint32_t clear_bits, uint8_t* __restrict__ flags,
                                                       uint32_t* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool sel = false;
    if (i < P) {
        uint32_t f = flags[i];
        const float v = (f & mask_bit) ? values[i] : 0.f;
        sel = v > *thr;
        f = (f & ~(set_bit | clear_bits)) | (sel ? set_bit : 0u);
        flags[i] = (uint8_t)f;
    }
    const int n = __syncthreads_count(sel);
    if (threadIdx.x == 0 && n) atomicAdd(count, (uint32_t)n);
}

struct PlanArrays {
    uint32_t* in[5];   // keep original, keep clone, keep split, clone (rank), split (rank)
    uint32_t* out[5];
};

__global__ void __launch_bounds__(256) plan_flags_kernel(const int P, const uint8_t* __restrict__ flags, const PlanArrays a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t f = flags[i];
    const bool clone = f & F_CLONE, split = (f & (F_SPLIT | F_ABS)) != 0;
    a.in[0][i] = (!split && !(f & F_PRUNE_SELF)) ? 1u : 0u;
    a.in[1][i] = (clone && !(f & F_PRUNE_SELF)) ? 1u : 0u;
    a.in[2][i] = (split && !(f & F_PRUNE_CHILD)) ? 1u : 0u;
    a.in[3][i] = clone ? 1u : 0u;
    a.in[4][i] = split ? 1u : 0u;
}

// totals: counts[C_KO..C_NS] = scan totals (device)
__global__ void __launch_bounds__(256) plan_scatter_kernel(const int P, const PlanArrays a, const uint32_t* __restrict__ counts,
                                                           uint32_t* __restrict__ row_map, int32_t* __restrict__ noise_index) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t n_ko = counts[C_KO], n_kc = counts[C_KC], n_ks = counts[C_KS], n_split = counts[C_NS];
    if (a.in[0][i]) {
        const uint32_t d = a.out[0][i];
        row_map[d] = (uint32_t)i;
        noise_index[d] = -1;
    }
    if (a.in[1][i]) {
        const uint32_t d = n_ko + a.out[1][i];
        row_map[d] = (uint32_t)i | (1u << 30);
        noise_index[d] = (int32_t)a.out[3][i];
    }
    if (a.in[2][i]) {
        const uint32_t d0 = n_ko + n_kc + a.out[2][i], d1 = d0 + n_ks;
        row_map[d0] = (uint32_t)i | (2u << 30);
        row_map[d1] = (uint32_t)i | (3u << 30);
        noise_index[d0] = (int32_t)a.out[4][i];
        noise_index[d1] = (int32_t)(a.out[4][i] + n_split);
    }
}

struct ApplyParams {
    long long n;      // P_new * width
    long long n_pad;  // n rounded up to the arena's group alignment: elements [n, n_pad) are zero-filled
    int width, role;
    const uint32_t* row_map;
    const int32_t* noise_index;
    const float* old_p;
    const float* old_m;
    const float* old_v;
    float* new_p;
    float* new_m;
    float* new_v;
    // sampling inputs (old arena): only read for role XYZ
    const float* scaling;
    const float* rotation;
    const float* z_clone;
    const float* z_split;
};

// W > 0: compile-time row width (constant division); W == 0: run-time width.  Four independent elements per thread and trip:
// 12 loads in flight per thread keep enough bytes outstanding to cover HBM latency with 4-byte accesses.
template <int W>
__global__ void __launch_bounds__(256) apply_kernel(const ApplyParams p) {
    constexpr int U = 4;
    const unsigned width = W > 0 ? (unsigned)W : (unsigned)p.width;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e0 = p.n + (long long)blockIdx.x * blockDim.x + threadIdx.x; e0 < p.n_pad; e0 += stride) {
        p.new_p[e0] = 0.f;  // alignment padding between groups (the Adam kernel streams over it)
        if (p.new_m) p.new_m[e0] = 0.f;
        if (p.new_v) p.new_v[e0] = 0.f;
    }
    for (long long e0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; e0 < p.n; e0 += U * stride) {
        float val[U], mv[U], vv[U];
        long long so[U];
        uint32_t kind[U];
        int col[U];
        long long row[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long e = e0 + u * stride;
            kind[u] = 0u;
            so[u] = -1;
            if (e < p.n) {
                const long long dst = p.n < 0x7fffffffll ? (long long)((unsigned)e / width) : e / width;
                col[u] = (int)(e - dst * width);
                row[u] = dst;
                const uint32_t rm = p.row_map[dst];
                kind[u] = rm >> 30;
                so[u] = (long long)(rm & 0x3fffffffu) * width + col[u];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (so[u] >= 0) {
                val[u] = p.old_p[so[u]];
                mv[u] = (p.old_m && kind[u] == 0u) ? p.old_m[so[u]] : 0.f;
                vv[u] = (p.old_v && kind[u] == 0u) ? p.old_v[so[u]] : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (so[u] < 0) continue;
            const long long e = e0 + u * stride;
            if (kind[u] != 0u) {
                if (p.role == LSX_DENSIFY_ROLE_XYZ) {
                    // new_xyz = R(q / |q|) (exp(s) * z) + xyz   (:686-689, :645-648; general_utils.py:80-101)
                    const long long src = so[u] / 3;
                    const int c = col[u];
                    const float* q = p.rotation + 4ll * src;
                    const float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
                    const float r = q[0] / nrm, x = q[1] / nrm, y = q[2] / nrm, zq = q[3] / nrm;
                    const int ni = p.noise_index[row[u]];
                    const float* z = (kind[u] == 1u ? p.z_clone : p.z_split) + 3ll * ni;
                    const float* s = p.scaling + 3ll * src;
                    const float v0 = expf(s[0]) * z[0], v1 = expf(s[1]) * z[1], v2 = expf(s[2]) * z[2];
                    float r0, r1, r2;
                    if (c == 0) {
                        r0 = 1.f - 2.f * (y * y + zq * zq), r1 = 2.f * (x * y - r * zq), r2 = 2.f * (x * zq + r * y);
                    } else if (c == 1) {
                        r0 = 2.f * (x * y + r * zq), r1 = 1.f - 2.f * (x * x + zq * zq), r2 = 2.f * (y * zq - r * x);
                    } else {
                        r0 = 2.f * (x * zq - r * y), r1 = 2.f * (y * zq + r * x), r2 = 1.f - 2.f * (x * x + y * y);
                    }
                    val[u] = (r0 * v0 + r1 * v1 + r2 * v2) + val[u];
                } else if (p.role == LSX_DENSIFY_ROLE_SCALING && kind[u] >= 2u) {
                    val[u] = logf(expf(val[u]) / 1.6f);  // scaling_inverse_activation(get_scaling / (0.8 * N))  (:650)
                }
            }
            p.new_p[e] = val[u];
            if (p.new_m) p.new_m[e] = mv[u];
            if (p.new_v) p.new_v[e] = vv[u];
        }
    }
}

void launch_apply(const ApplyParams& p, int blocks, cudaStream_t stream) {
    switch (p.width) {
        case 1: apply_kernel<1><<<blocks, 256, 0, stream>>>(p); break;
        case 3: apply_kernel<3><<<blocks, 256, 0, stream>>>(p); break;
        case 4: apply_kernel<4><<<blocks, 256, 0, stream>>>(p); break;
        case 16: apply_kernel<16><<<blocks, 256, 0, stream>>>(p); break;
        case 45: apply_kernel<45><<<blocks, 256, 0, stream>>>(p); break;
        case 48: apply_kernel<48><<<blocks, 256, 0, stream>>>(p); break;
        default: apply_kernel<0><<<blocks, 256, 0, stream>>>(p); break;
    }
}

__global__ void __launch_bounds__(256) reset_opacity_kernel(const int P, float* __restrict__ opacity, float* __restrict__ m,
                                                            float* __restrict__ v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float x = fminf(sigmoidf_ref(opacity[i]), 0.01f);  // inverse_sigmoid(min(get_opacity, 0.01))  (:444)
    opacity[i] = logf(x / (1.0f - x));
    if (m) m[i] = 0.f;
    if (v) v[i] = 0.f;
}

// ---- workspace carving ----------------------------------------------------------------------------------------------
struct Workspace {
    uint8_t* flags;
    float* g;
    float* ga;
    uint32_t* counts;  // C_COUNT
    float* thr;        // 4
    PlanArrays plan;   // 10 x u32[P]; the capped branches alias the first four as sort ping-pong buffers
    void* temp;        // scan / sort temporaries
    uint32_t* row_map;     // 2P
    int32_t* noise_index;  // 2P
    size_t bytes;
};

Workspace carve(char* base, int P) {
    Workspace w{};
    size_t off = 0;
    const size_t n = (size_t)(P > 0 ? P : 1);
    auto take = [&](size_t b) {
        char* ptr = base ? base + off : nullptr;
        off += align_up(b, 256);
        return ptr;
    };
    w.flags = reinterpret_cast<uint8_t*>(take(n));
    w.g = reinterpret_cast<float*>(take(n * 4));
    w.ga = reinterpret_cast<float*>(take(n * 4));
    w.counts = reinterpret_cast<uint32_t*>(take(C_COUNT * 4));
    w.thr = reinterpret_cast<float*>(take(16));
    for (int k = 0; k < 5; ++k) w.plan.in[k] = reinterpret_cast<uint32_t*>(take(n * 4));
    for (int k = 0; k < 5; ++k) w.plan.out[k] = reinterpret_cast<uint32_t*>(take(n * 4));
    const size_t t1 = scan_temp_bytes((int)n), t2 = radix_sort_temp_bytes((int)n);
    w.temp = take(t1 > t2 ? t1 : t2);
    w.row_map = reinterpret_cast<uint32_t*>(take(2 * n * 4));
    w.noise_index = reinterpret_cast<int32_t*>(take(2 * n * 4));
    w.bytes = off;
    return w;
}

inline int blocks_for(long long n) { return (int)((n + 255) / 256); }

// quantile of [n_pad zeros | masked values] -> *w.thr ; rank arithmetic in float32 like ATen (q and n - 1 as floats)
int quantile_to_threshold(const Workspace& w, int P, const float* values, uint32_t mask_bit, int n_pad, double q,
                          cudaStream_t stream) {
    if (!(q >= 0.0 && q <= 1.0)) {
        set_error("densify: quantile() q values must be in the range [0, 1] (got %g): the point budget max_all_points is "
                  "already exceeded", q);
        return -5;
    }
    uint32_t* keys[2] = {w.plan.in[0], w.plan.in[1]};
    uint32_t* vals[2] = {w.plan.in[2], w.plan.in[3]};
    masked_keys_kernel<<<blocks_for(P), 256, 0, stream>>>(P, values, w.flags, mask_bit, keys[0]);
    LSX_KERNEL_OK(stream, false);
    int res = 0;
    int rc = radix_sort_pairs_u32(keys, vals, P, 0, 32, true, w.temp, &res, stream, false);
    if (rc) return rc;
    const long long n = (long long)P + n_pad;
    const float rank = (float)q * (float)(double)(n - 1);
    const float fl = floorf(rank), ce = ceilf(rank);
    long long lo = (long long)fl, hi = (long long)ce;
    if (lo < 0) lo = 0;
    if (hi > n - 1) hi = n - 1;  // float32 rank can round up to n for huge n
    if (lo > n - 1) lo = n - 1;
    quantile_pick_kernel<<<1, 1, 0, stream>>>(keys[res], n_pad, (int)lo, (int)hi, rank - fl, w.thr);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

int read_counts(const Workspace& w, uint32_t* host, cudaStream_t stream) {
    LSX_CUDA_OK(cudaMemcpyAsync(host, w.counts, C_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    LSX_CUDA_OK(cudaStreamSynchronize(stream));
    return 0;
}

}  // namespace
}  // namespace lsx

using namespace lsx;

extern "C" int lsx_densify_stats_update(int32_t P, const float* dL_dmeans2D, const float* dL_dmeans2D_abs, const int32_t* radii,
                                        const int32_t* out_observe, float* grad_accum, float* grad_accum_abs, float* denom,
                                        float* max_radii2D, void* stream_) {
    if (P < 0 || (P > 0 && (!dL_dmeans2D || !dL_dmeans2D_abs || !radii || !grad_accum || !grad_accum_abs || !denom ||
                            !max_radii2D))) {
        set_error("lsx_densify_stats_update: bad arguments");
        return -1;
    }
    if (P == 0) return 0;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    stats_update_kernel<<<blocks_for(P), 256, 0, stream>>>(P, dL_dmeans2D, dL_dmeans2D_abs, radii, out_observe, grad_accum,
                                                           grad_accum_abs, denom, max_radii2D);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

extern "C" size_t lsx_densify_workspace_bytes(int32_t P) { return carve(nullptr, P).bytes; }

extern "C" int lsx_densify_plan(const lsx_densify_plan_args* a, lsx_densify_plan_result* out) {
    if (!a || !out || a->P < 0 ||
        (a->P > 0 && (!a->grad_accum || !a->grad_accum_abs || !a->denom || !a->max_radii2D || !a->scaling_raw ||
                      !a->opacity_raw || !a->workspace))) {
        set_error("lsx_densify_plan: bad arguments");
        return -1;
    }
    if (!(a->max_grad > 0.f) || !(a->abs_max_grad > 0.f)) {
        set_error("lsx_densify_plan: gradient thresholds must be > 0 (a zero threshold would select the zero-padded rows)");
        return -1;
    }
    *out = lsx_densify_plan_result{};
    const int P = a->P;
    if (P == 0) return 0;
    if (a->workspace_bytes < carve(nullptr, P).bytes) {
        set_error("lsx_densify_plan: workspace too small (%zu < %zu bytes)", a->workspace_bytes, carve(nullptr, P).bytes);
        return -1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(a->stream);
    const Workspace w = carve(static_cast<char*>(a->workspace), P);
    const long long max_all = a->max_all_points;
    LSX_CUDA_OK(cudaMemsetAsync(w.counts, 0, C_COUNT * sizeof(uint32_t), stream));

    ClassifyParams cp{};
    cp.P = P;
    cp.accum = a->grad_accum;
    cp.accum_abs = a->grad_accum_abs;
    cp.denom = a->denom;
    cp.max_radii = a->max_radii2D;
    cp.scaling = a->scaling_raw;
    cp.opacity = a->opacity_raw;
    cp.max_grad = a->max_grad;
    cp.abs_max_grad = a->abs_max_grad;
    cp.dense_extent = (float)((double)a->percent_dense * (double)a->extent);
    cp.abs_radii_thr = a->abs_split_radii2D_threshold;
    cp.min_opacity = a->min_opacity;
    cp.world_extent = a->prune_world_size ? (float)(0.1 * (double)a->extent) : -1.0f;
    cp.flags = w.flags;
    cp.g = w.g;
    cp.ga = w.ga;
    cp.counts = w.counts;
    classify_kernel<<<blocks_for(P), 256, 0, stream>>>(cp);
    LSX_KERNEL_OK(stream, false);

    uint32_t h[C_COUNT];
    int rc = read_counts(w, h, stream);
    if (rc) return rc;
    long long n_clone = h[C_CLONE], n_split = h[C_SPLIT], n_abs = h[C_ABS];

    // ---- the reference's budget branches (host decisions on three counts) ----
    auto reselect = [&](const float* values, uint32_t mask_bit, uint32_t set_bit, uint32_t clear_bits, int n_pad, double q,
                        long long* n_sel) -> int {
        int r = quantile_to_threshold(w, P, values, mask_bit, n_pad, q, stream);
        if (r) return r;
        LSX_CUDA_OK(cudaMemsetAsync(w.counts + C_RESEL, 0, sizeof(uint32_t), stream));
        reselect_kernel<<<blocks_for(P), 256, 0, stream>>>(P, values, mask_bit, w.thr, set_bit, clear_bits, w.flags,
                                                           w.counts + C_RESEL);
        LSX_KERNEL_OK(stream, false);
        r = read_counts(w, h, stream);
        if (r) return r;
        *n_sel = h[C_RESEL];
        return 0;
    };
    if (n_clone + P > max_all) {  // :671-677
        const double ratio = fmin((double)(max_all - P) / (double)P, 1.0);
        rc = reselect(w.g, F_CLONE, F_CLONE, 0u, 0, 1.0 - ratio, &n_clone);
        if (rc) return rc;
        out->clone_capped = 1;
    }
    const long long n_init = P + n_clone;
    if (n_split + n_init > max_all) {  // :625-630 (no abs pass in this branch)
        const double ratio = (double)(max_all - n_init) / (double)n_init;
        rc = reselect(w.g, F_SPLIT, F_SPLIT, F_ABS, (int)n_clone, 1.0 - ratio, &n_split);
        if (rc) return rc;
        n_abs = 0;
        out->split_capped = 1;
    } else {  // :632-642
        long long limited = max_all - n_init - n_split;
        if (a->max_abs_split_points < limited) limited = a->max_abs_split_points;
        if (n_abs > limited) {
            const double ratio = (double)limited / (double)n_init;
            rc = reselect(w.ga, F_ABS_ELIG, F_ABS, 0u, (int)n_clone, 1.0 - ratio, &n_abs);
            if (rc) return rc;
            out->abs_capped = 1;
        }
    }

    // ---- destination rows ----
    plan_flags_kernel<<<blocks_for(P), 256, 0, stream>>>(P, w.flags, w.plan);
    LSX_KERNEL_OK(stream, false);
    for (int k = 0; k < 5; ++k) {
        rc = exclusive_scan_u32(w.plan.in[k], nullptr, w.plan.out[k], P, w.counts + C_KO + k, w.temp, stream, false);
        if (rc) return rc;
    }
    plan_scatter_kernel<<<blocks_for(P), 256, 0, stream>>>(P, w.plan, w.counts, w.row_map, w.noise_index);
    LSX_KERNEL_OK(stream, false);
    rc = read_counts(w, h, stream);
    if (rc) return rc;
    out->n_clone = h[C_NC];
    out->n_split = h[C_NS];
    out->n_split_abs = (int64_t)n_abs;
    out->n_kept_original = h[C_KO];
    out->n_kept_clone = h[C_KC];
    out->n_kept_split = h[C_KS];
    out->P_new = (int64_t)h[C_KO] + h[C_KC] + 2ll * h[C_KS];
    out->row_map = w.row_map;
    out->noise_index = w.noise_index;
    if ((long long)h[C_NC] != n_clone || (long long)h[C_NS] != n_split + n_abs) {
        set_error("lsx_densify_plan: internal count mismatch (clone %u vs %lld, split %u vs %lld)", h[C_NC], n_clone, h[C_NS],
                  n_split + n_abs);
        return -6;
    }
    return 0;
}

extern "C" int lsx_densify_apply(const lsx_densify_apply_args* a) {
    if (!a || a->P_new < 0 || a->n_groups <= 0 || a->n_groups > LSX_ADAM_MAX_GROUPS || !a->old_begin || !a->new_begin ||
        !a->width || !a->role || (a->P_new > 0 && (!a->row_map || !a->noise_index || !a->old_params || !a->new_params))) {
        set_error("lsx_densify_apply: bad arguments");
        return -1;
    }
    if ((a->old_exp_avg == nullptr) != (a->new_exp_avg == nullptr) || (a->old_exp_avg_sq == nullptr) != (a->new_exp_avg_sq == nullptr)) {
        set_error("lsx_densify_apply: old and new moment arenas must be given together");
        return -1;
    }
    if (a->P_new == 0) return 0;
    int g_xyz = -1, g_scaling = -1, g_rotation = -1;
    for (int k = 0; k < a->n_groups; ++k) {
        if (a->role[k] == LSX_DENSIFY_ROLE_XYZ) g_xyz = k;
        if (a->role[k] == LSX_DENSIFY_ROLE_SCALING) g_scaling = k;
        if (a->role[k] == LSX_DENSIFY_ROLE_ROTATION) g_rotation = k;
    }
    const bool any_new = a->n_new_rows > 0;
    if (any_new && (g_xyz < 0 || g_scaling < 0 || g_rotation < 0 || a->width[g_xyz] != 3 || a->width[g_scaling] != 3 ||
                    a->width[g_rotation] != 4 || (!a->z_clone && !a->z_split))) {
        set_error("lsx_densify_apply: new rows need xyz (3), scaling (3), rotation (4) groups and their noise arrays");
        return -1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(a->stream);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    for (int k = 0; k < a->n_groups; ++k) {
        if (a->width[k] <= 0) continue;
        ApplyParams p{};
        p.n = (long long)a->P_new * a->width[k];
        p.n_pad = a->group_align > 1 ? (p.n + a->group_align - 1) / a->group_align * a->group_align : p.n;
        p.width = a->width[k];
        p.role = a->role[k];
        p.row_map = a->row_map;
        p.noise_index = a->noise_index;
        p.old_p = a->old_params + a->old_begin[k];
        p.new_p = a->new_params + a->new_begin[k];
        p.old_m = a->old_exp_avg ? a->old_exp_avg + a->old_begin[k] : nullptr;
        p.new_m = a->new_exp_avg ? a->new_exp_avg + a->new_begin[k] : nullptr;
        p.old_v = a->old_exp_avg_sq ? a->old_exp_avg_sq + a->old_begin[k] : nullptr;
        p.new_v = a->new_exp_avg_sq ? a->new_exp_avg_sq + a->new_begin[k] : nullptr;
        p.scaling = g_scaling >= 0 ? a->old_params + a->old_begin[g_scaling] : nullptr;
        p.rotation = g_rotation >= 0 ? a->old_params + a->old_begin[g_rotation] : nullptr;
        p.z_clone = a->z_clone;
        p.z_split = a->z_split;
        const long long want = (p.n + 1023) / 1024, cap = (long long)sms * 16;
        launch_apply(p, (int)(want < cap ? want : cap), stream);
        LSX_KERNEL_OK(stream, false);
    }
    return 0;
}

extern "C" int lsx_reset_opacity(int32_t P, float* opacity_raw, float* exp_avg, float* exp_avg_sq, void* stream_) {
    if (P < 0 || (P > 0 && !opacity_raw)) {
        set_error("lsx_reset_opacity: bad arguments");
        return -1;
    }
    if (P == 0) return 0;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    reset_opacity_kernel<<<blocks_for(P), 256, 0, stream>>>(P, opacity_raw, exp_avg, exp_avg_sq);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

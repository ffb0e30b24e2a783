// cls3d.cu — the 3-D neighbourhood regulariser of the language / instance features (second half).
//
// Reference behaviour restated: loss_cls_3d, field_construction/utils/loss_utils.py:158-186, called once per iteration
// of the language stage at field_construction/gaussian_field.py:461-465 and :482-485 with features = xyz (N,3),
// predictions = the (N,C) language or instance feature, k = 5, 800 samples:
//     q      = (predictions - min) / (max - min)          over ALL elements, only if max > min
//     S rows = a random sample of the points;  nbr = the k nearest points of each sample (the sample itself included)
//     loss   = lambda * mean | q[s] * (log(q[s] + 1e-10) - log(q[nbr] + 1e-10)) |      over S x k x C terms
// The reference materialises the S x N distance matrix with torch.cdist (1.6 GB at 800 x 500 k) and runs topk over it.
//
// Here:  forward = 4 launches, no S x N matrix, no host read
//   cls3d_minmax_kernel   per-block min / max of the predictions
//   cls3d_knn_kernel<K>   grid = (groups of 8 samples) x (point slices).  The CTA streams its slice through a 12 KB
//                         shared-memory tile; warp w scans the tile for sample w: every lane keeps the K best of its
//                         strided subset in registers (sorted insertion, rarely taken after the first tiles), then the 32
//                         lists are merged by K rounds of a lexicographic (distance, index) warp minimum.
//   cls3d_merge_kernel<K> one warp per sample: merges the slices' candidates, writes the neighbour indices and the
//                         sample's sum of |kl| terms
//   cls3d_finish_kernel   fixed-order sum of the S partial sums in double -> loss
// Distances are the exact squared differences (the reference's cdist uses the |x|^2 + |y|^2 - 2 x.y matmul form, whose
// rounding error is comparable to the neighbour spacing of a dense cloud; the neighbour SET is what matters for the
// loss, the order inside it does not).  Ties are broken towards the lower index.
// backward = zero-fill + scatter of the S x (k + 1) x C non-zero terms + tie count + one dense pass that applies the
// normalisation's chain rule (including the gradient through min and max, distributed evenly among ties like torch).
#include <float.h>
#include <limits.h>

#include "../../include/lsx_rasterizer.h"
#include "kernels.cuh"

namespace lsx {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxK = 8;
constexpr int kMmBlocks = 296;        // min/max partials
constexpr int kQueriesPerBlock = 8;   // one warp per sample
constexpr int kTilePoints = 1024;     // points per shared-memory tile (12 KB)
constexpr int kMaxSlices = 16;
constexpr float kEps = 1e-10f;

struct Cls3dLayout {
    int slices;
    size_t mm_off, cand_d_off, cand_i_off, qsum_off, acc_off, total;
};

int pick_slices(int N, int S) {
    const int groups = ceil_div(S, kQueriesPerBlock);
    int ns = ceil_div(4 * 148, groups);
    const int by_points = N / 2048 > 1 ? N / 2048 : 1;
    if (ns > by_points) ns = by_points;
    if (ns > kMaxSlices) ns = kMaxSlices;
    return ns < 1 ? 1 : ns;
}

Cls3dLayout make_layout(int N, int S, int k) {
    Cls3dLayout L;
    L.slices = pick_slices(N, S);
    size_t off = 0;
    L.mm_off = off;
    off = align_up(off + 2 * kMmBlocks * sizeof(float), 256);
    L.cand_d_off = off;
    off = align_up(off + (size_t)S * L.slices * k * sizeof(float), 256);
    L.cand_i_off = off;
    off = align_up(off + (size_t)S * L.slices * k * sizeof(int), 256);
    L.qsum_off = off;
    off = align_up(off + (size_t)S * sizeof(float), 256);
    L.acc_off = off;  // backward: double dm_rest, dM, G_min; uint32 count_min, count_max
    off = align_up(off + 3 * sizeof(double) + 2 * sizeof(uint32_t), 256);
    L.total = off;
    return L;
}

// ---- min / max ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cls3d_minmax_kernel(const long long n, const float* __restrict__ x,
                                                           float* __restrict__ partial) {
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(kFull, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(kFull, hi, o));
    }
    __shared__ float s_lo[8], s_hi[8];
    if ((threadIdx.x & 31) == 0) {
        s_lo[threadIdx.x >> 5] = lo;
        s_hi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) {
            lo = fminf(lo, s_lo[w]);
            hi = fmaxf(hi, s_hi[w]);
        }
        partial[2 * blockIdx.x] = lo;
        partial[2 * blockIdx.x + 1] = hi;
    }
}

// all lanes return the global {min, max}
__device__ __forceinline__ float2 warp_minmax(const float* __restrict__ partial, const int nblocks, const unsigned lane) {
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (int b = (int)lane; b < nblocks; b += 32) {
        lo = fminf(lo, partial[2 * b]);
        hi = fmaxf(hi, partial[2 * b + 1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(kFull, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(kFull, hi, o));
    }
    return make_float2(lo, hi);
}

// ---- k nearest neighbours of the samples ----------------------------------------------------------------------------
__device__ __forceinline__ bool cand_less(const float d0, const int i0, const float d1, const int i1) {
    return d0 < d1 || (d0 == d1 && i0 < i1);
}

// K rounds of a lexicographic warp minimum over the heads of the lanes' sorted lists; lane 0 stores round r's winner
template <int K>
__device__ __forceinline__ void warp_merge_sorted(float (&dist)[K], int (&id)[K], const unsigned lane, float* __restrict__ out_d,
                                                  int* __restrict__ out_i) {
#pragma unroll
    for (int r = 0; r < K; ++r) {
        float bd = dist[0];
        int bi = id[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(kFull, bd, o);
            const int oi = __shfl_xor_sync(kFull, bi, o);
            if (cand_less(od, oi, bd, bi)) {
                bd = od;
                bi = oi;
            }
        }
        if (bi == id[0] && bd == dist[0]) {  // this lane's head won: pop it (indices are unique across lanes; the
#pragma unroll                               // all-empty case pops everywhere, harmlessly)
            for (int j = 0; j + 1 < K; ++j) {
                dist[j] = dist[j + 1];
                id[j] = id[j + 1];
            }
            dist[K - 1] = FLT_MAX;
            id[K - 1] = INT_MAX;
        }
        if (lane == 0) {
            out_d[r] = bd;
            out_i[r] = bi;
        }
    }
}

template <int K>
__global__ void __launch_bounds__(256) cls3d_knn_kernel(const int N, const int S, const int slices,
                                                        const float* __restrict__ points,
                                                        const int* __restrict__ sample_idx, float* __restrict__ cand_d,
                                                        int* __restrict__ cand_i) {
    __shared__ float s_pts[kTilePoints * 3];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * kQueriesPerBlock + (int)warp;
    const int slice = blockIdx.y;
    const long long per = ((long long)N + slices - 1) / slices;
    const int begin = (int)(per * slice), end = (int)((per * (slice + 1) < N) ? per * (slice + 1) : N);

    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (q < S) {
        const int si = sample_idx[q];
        qx = points[3 * (size_t)si];
        qy = points[3 * (size_t)si + 1];
        qz = points[3 * (size_t)si + 2];
    }
    float dist[K];
    int id[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        dist[j] = FLT_MAX;
        id[j] = INT_MAX;
    }
    for (int t0 = begin; t0 < end; t0 += kTilePoints) {
        const int cnt = (end - t0 < kTilePoints) ? end - t0 : kTilePoints;
        __syncthreads();  // previous tile consumed
        for (int e = threadIdx.x; e < cnt * 3; e += 256) s_pts[e] = points[3 * (size_t)t0 + e];
        __syncthreads();
        if (q < S) {
            for (int i = (int)lane; i < cnt; i += 32) {
                const float dx = s_pts[3 * i] - qx, dy = s_pts[3 * i + 1] - qy, dz = s_pts[3 * i + 2] - qz;
                const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                if (d < dist[K - 1]) {  // a lane sees its indices in increasing order: equal distances keep the earlier one
                    const int gi = t0 + i;
#pragma unroll
                    for (int j = K - 1; j > 0; --j) {
                        const bool shift = d < dist[j - 1];
                        const bool here = !shift && d < dist[j];
                        if (shift) {
                            dist[j] = dist[j - 1];
                            id[j] = id[j - 1];
                        } else if (here) {
                            dist[j] = d;
                            id[j] = gi;
                        }
                    }
                    if (d < dist[0]) {
                        dist[0] = d;
                        id[0] = gi;
                    }
                }
            }
        }
    }
    if (q < S) {
        const size_t o = ((size_t)q * slices + slice) * K;
        warp_merge_sorted<K>(dist, id, lane, cand_d + o, cand_i + o);
    }
}

// one warp per sample: merge the slices' candidate lists, write the neighbour indices, sum the sample's |kl| terms
template <int K>
__global__ void __launch_bounds__(256) cls3d_merge_kernel(const int N, const int C, const int S, const int slices,
                                                          const float* __restrict__ preds, const int* __restrict__ sample_idx,
                                                          const float* __restrict__ cand_d, const int* __restrict__ cand_i,
                                                          const float* __restrict__ mm_partial, const int mm_blocks,
                                                          int* __restrict__ nbr_idx, float* __restrict__ minmax,
                                                          float* __restrict__ qsum) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * kQueriesPerBlock + (int)warp;
    if (q >= S) return;
    const float2 mm = warp_minmax(mm_partial, mm_blocks, lane);
    if (q == 0 && lane == 0) {
        minmax[0] = mm.x;
        minmax[1] = mm.y;
    }
    // candidates: slices * K <= 128, up to 4 per lane
    constexpr int kPer = kMaxSlices * kMaxK / 32;
    float cd[kPer];
    int ci[kPer];
    const int ncand = slices * K;
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
        const int e = u * 32 + (int)lane;
        cd[u] = e < ncand ? cand_d[(size_t)q * ncand + e] : FLT_MAX;
        ci[u] = e < ncand ? cand_i[(size_t)q * ncand + e] : INT_MAX;
    }
    int nbr[K];
#pragma unroll
    for (int r = 0; r < K; ++r) {
        float bd = cd[0];
        int bi = ci[0];
#pragma unroll
        for (int u = 1; u < kPer; ++u)
            if (cand_less(cd[u], ci[u], bd, bi)) {
                bd = cd[u];
                bi = ci[u];
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(kFull, bd, o);
            const int oi = __shfl_xor_sync(kFull, bi, o);
            if (cand_less(od, oi, bd, bi)) {
                bd = od;
                bi = oi;
            }
        }
#pragma unroll
        for (int u = 0; u < kPer; ++u)
            if (ci[u] == bi) {  // indices are unique among the candidates
                cd[u] = FLT_MAX;
                ci[u] = INT_MAX;
            }
        nbr[r] = bi;
        if (lane == 0) nbr_idx[(size_t)q * K + r] = bi;
    }
    // |kl| terms of this sample
    const bool norm = mm.y > mm.x;
    const float lo = norm ? mm.x : 0.f, range = norm ? __fsub_rn(mm.y, mm.x) : 1.f;
    const int si = sample_idx[q];
    float acc = 0.f;
    for (int e = (int)lane; e < K * C; e += 32) {
        const int r = e / C, c = e - r * C;
        int nb = nbr[0];
#pragma unroll
        for (int j = 1; j < K; ++j)
            if (j == r) nb = nbr[j];
        const float a = __fdiv_rn(__fsub_rn(preds[(size_t)si * C + c], lo), range);
        const float b = __fdiv_rn(__fsub_rn(preds[(size_t)nb * C + c], lo), range);
        const float kl = __fmul_rn(a, __fsub_rn(logf(__fadd_rn(a, kEps)), logf(__fadd_rn(b, kEps))));
        acc += fabsf(kl);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
    if (lane == 0) qsum[q] = acc;
}

__global__ void __launch_bounds__(256) cls3d_finish_kernel(const int S, const float* __restrict__ qsum, const double scale,
                                                           float* __restrict__ loss) {
    __shared__ double s_part[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < S; i += 256) acc += (double)qsum[i];
    s_part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s_part[threadIdx.x] += s_part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = (float)(s_part[0] * scale);
}

// ---- backward -------------------------------------------------------------------------------------------------------
// one thread per (sample, channel): adds d loss / d q into the zero-filled output for the sample's own row and its k
// neighbours, and accumulates (in double) the gradient through min and max:  q = (p - m) / (M - m)  =>
// dq/dm = (q - 1) / (M - m), dq/dM = -q / (M - m).  Terms that land on a minimum element (q == 0) are kept apart in G_min:
// there log(0 + 1e-10) makes d loss / d q ~ 1e10 * q_s, and the element's two paths (directly, and through m) cancel —
// exactly when the minimum is unique.  The reference's fp32 autograd leaves rounding noise of that cancellation in the
// minimum element's gradient; here it is carried out in double / analytically (see cls3d_finalize_grad_kernel).
__global__ void __launch_bounds__(256) cls3d_scatter_kernel(const int C, const int S, const int k,
                                                            const float* __restrict__ preds,
                                                            const int* __restrict__ sample_idx,
                                                            const int* __restrict__ nbr_idx,
                                                            const float* __restrict__ minmax, const float term_scale,
                                                            float* __restrict__ gq, double* __restrict__ acc) {
    const float m = minmax[0], M = minmax[1];
    const bool norm = M > m;
    const float lo = norm ? m : 0.f, range = norm ? __fsub_rn(M, m) : 1.f;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double dm = 0.0, dM = 0.0, gmin = 0.0;
    auto through_extrema = [&](const float gv, const float qv) {
        if (qv == 0.f) gmin += (double)gv;
        else dm += (double)gv * ((double)qv - 1.0);
        dM -= (double)gv * (double)qv;
    };
    if (t < S * C) {
        const int s = t / C, c = t - s * C;
        const int si = sample_idx[s];
        const float a = __fdiv_rn(__fsub_rn(preds[(size_t)si * C + c], lo), range);
        const float la = logf(__fadd_rn(a, kEps));
        float ga = 0.f;
        for (int j = 0; j < k; ++j) {
            const int nb = nbr_idx[(size_t)s * k + j];
            const float b = __fdiv_rn(__fsub_rn(preds[(size_t)nb * C + c], lo), range);
            const float diff = __fsub_rn(la, logf(__fadd_rn(b, kEps)));
            const float kl = __fmul_rn(a, diff);
            const float sg = kl > 0.f ? term_scale : (kl < 0.f ? -term_scale : 0.f);
            ga += sg * (diff + a / (a + kEps));
            const float gb = -sg * a / (b + kEps);
            if (gb != 0.f) {
                atomicAdd(gq + (size_t)nb * C + c, gb);
                through_extrema(gb, b);
            }
        }
        if (ga != 0.f) {
            atomicAdd(gq + (size_t)si * C + c, ga);
            through_extrema(ga, a);
        }
    }
    if (!norm) return;  // block-uniform
    __shared__ double s_red[3][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dm += __shfl_xor_sync(kFull, dm, o);
        dM += __shfl_xor_sync(kFull, dM, o);
        gmin += __shfl_xor_sync(kFull, gmin, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_red[0][threadIdx.x >> 5] = dm;
        s_red[1][threadIdx.x >> 5] = dM;
        s_red[2][threadIdx.x >> 5] = gmin;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            dm += s_red[0][w];
            dM += s_red[1][w];
            gmin += s_red[2][w];
        }
        atomicAdd(acc, dm / (double)range);
        atomicAdd(acc + 1, dM / (double)range);
        atomicAdd(acc + 2, gmin);
    }
}

__global__ void __launch_bounds__(256) cls3d_count_ties_kernel(const long long n, const float* __restrict__ preds,
                                                               const float* __restrict__ minmax,
                                                               uint32_t* __restrict__ counts) {
    const float m = minmax[0], M = minmax[1];
    uint32_t cm = 0, cM = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = preds[i];
        cm += v == m;
        cM += v == M;
    }
    cm = __reduce_add_sync(kFull, cm);
    cM = __reduce_add_sync(kFull, cM);
    if ((threadIdx.x & 31) == 0) {
        if (cm) atomicAdd(counts, cm);
        if (cM) atomicAdd(counts + 1, cM);
    }
}

// d loss / d preds = upstream * ( gq / (M - m) + [p == m] dm / #min + [p == M] dM / #max ).  With the minimum elements' own
// terms kept apart (dm = dm_rest - G_min / (M - m)) a minimum element gets  dm_rest / #min + (gq - G_min / #min) / (M - m),
// evaluated in double; for a unique minimum gq == G_min and the second term vanishes identically.
__global__ void __launch_bounds__(256) cls3d_finalize_grad_kernel(const long long n, const float* __restrict__ preds,
                                                                  const float* __restrict__ minmax,
                                                                  const double* __restrict__ acc,
                                                                  const uint32_t* __restrict__ counts,
                                                                  const float* __restrict__ upstream, float* __restrict__ g) {
    const float m = minmax[0], M = minmax[1];
    const bool norm = M > m;
    const float up = upstream ? *upstream : 1.0f;
    const float inv = norm ? up / __fsub_rn(M, m) : up;
    const double inv_d = norm ? 1.0 / (double)__fsub_rn(M, m) : 1.0;
    const double cm = norm ? (double)counts[0] : 1.0, cM = norm ? (double)counts[1] : 1.0;
    const double tm = norm ? acc[0] / cm : 0.0;
    const float tM = norm ? up * (float)(acc[1] / cM) : 0.f;
    const double gmin_share = norm ? acc[2] / cm : 0.0;
    const bool unique_min = norm && counts[0] == 1u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = preds[i];
        float o;
        if (norm && v == m) {
            const double own = unique_min ? 0.0 : ((double)g[i] - gmin_share) * inv_d;
            o = up * (float)(tm + own);
        } else {
            o = g[i] * inv;
            if (norm && v == M) o += tM;
        }
        g[i] = o;
    }
}

template <int K>
int launch_forward_k(const int N, const int C, const int S, const Cls3dLayout& L, const float* points, const float* preds,
                     const int* sample_idx, char* scratch, int* nbr_idx, float* minmax, const int mm_blocks,
                     const void* tree, cudaStream_t stream) {
    float* cand_d = reinterpret_cast<float*>(scratch + L.cand_d_off);
    int* cand_i = reinterpret_cast<int*>(scratch + L.cand_i_off);
    const int groups = ceil_div(S, kQueriesPerBlock);
    int slices = L.slices;
    if (tree != nullptr) {
        // neighbour search through a prebuilt Morton / box hierarchy of the points (knn.cu): one candidate list per sample
        slices = 1;
        const int rc = knn_tree_query(K, N, tree, S, points, sample_idx, cand_d, cand_i, stream);
        if (rc) return rc;
    } else {
        cls3d_knn_kernel<K><<<dim3(groups, L.slices), 256, 0, stream>>>(N, S, L.slices, points, sample_idx, cand_d, cand_i);
        LSX_KERNEL_OK(stream, false);
    }
    cls3d_merge_kernel<K><<<groups, 256, 0, stream>>>(N, C, S, slices, preds, sample_idx, cand_d, cand_i,
                                                      reinterpret_cast<const float*>(scratch + L.mm_off), mm_blocks, nbr_idx,
                                                      minmax, reinterpret_cast<float*>(scratch + L.qsum_off));
    LSX_KERNEL_OK(stream, false);
    return 0;
}

int dense_blocks(long long n) {
    const long long want = (n + 1023) / 1024;
    return (int)(want < 1 ? 1 : (want < kMmBlocks ? want : kMmBlocks));
}

bool bad_sizes(int N, int C, int S, int k) { return N <= 0 || C <= 0 || S <= 0 || k <= 0 || k > kMaxK || k > N; }

}  // namespace
}  // namespace lsx

using namespace lsx;

extern "C" int64_t lsx_cls3d_scratch_bytes(int32_t N, int32_t C, int32_t S, int32_t k) {
    if (bad_sizes(N, C, S, k)) return 0;
    return (int64_t)make_layout(N, S, k).total;
}

extern "C" int64_t lsx_knn_tree_bytes(int32_t N) { return N > 0 ? (int64_t)knn_temp_bytes(N) : 0; }

extern "C" int lsx_knn_tree_build(int32_t N, const float* points, void* tree, void* stream) {
    if (N <= 0 || !points || !tree) {
        set_error("lsx_knn_tree_build: bad arguments");
        return -1;
    }
    return knn_tree_build(N, points, tree, static_cast<cudaStream_t>(stream));
}

static int cls3d_forward_impl(int32_t N, int32_t C, int32_t S, int32_t k, float lambda_val, const float* points, const float* preds,
                              const int32_t* sample_idx, float* loss, int32_t* nbr_idx, float* minmax, void* scratch_,
                              const void* tree, void* stream_) {
    if (bad_sizes(N, C, S, k)) {
        set_error("lsx_cls3d_forward: need N, C, S > 0 and 1 <= k <= min(N, %d) (got N=%d C=%d S=%d k=%d)", kMaxK, N, C, S, k);
        return -1;
    }
    if (!points || !preds || !sample_idx || !loss || !nbr_idx || !minmax || !scratch_) {
        set_error("lsx_cls3d_forward: null argument");
        return -1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    char* scratch = static_cast<char*>(scratch_);
    const Cls3dLayout L = make_layout(N, S, k);
    const long long n = (long long)N * C;
    const int mm_blocks = dense_blocks(n);
    cls3d_minmax_kernel<<<mm_blocks, 256, 0, stream>>>(n, preds, reinterpret_cast<float*>(scratch + L.mm_off));
    LSX_KERNEL_OK(stream, false);
    int rc = 0;
    switch (k) {
#define LSX_CLS3D_CASE(KK) \
    case KK: rc = launch_forward_k<KK>(N, C, S, L, points, preds, sample_idx, scratch, nbr_idx, minmax, mm_blocks, tree, stream); break;
        LSX_CLS3D_CASE(1)
        LSX_CLS3D_CASE(2)
        LSX_CLS3D_CASE(3)
        LSX_CLS3D_CASE(4)
        LSX_CLS3D_CASE(5)
        LSX_CLS3D_CASE(6)
        LSX_CLS3D_CASE(7)
        LSX_CLS3D_CASE(8)
#undef LSX_CLS3D_CASE
    }
    if (rc) return rc;
    const double scale = (double)lambda_val / ((double)S * (double)k * (double)C);
    cls3d_finish_kernel<<<1, 256, 0, stream>>>(S, reinterpret_cast<const float*>(scratch + L.qsum_off), scale, loss);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

extern "C" int lsx_cls3d_forward(int32_t N, int32_t C, int32_t S, int32_t k, float lambda_val, const float* points,
                                 const float* preds, const int32_t* sample_idx, float* loss, int32_t* nbr_idx, float* minmax,
                                 void* scratch_, void* stream_) {
    return cls3d_forward_impl(N, C, S, k, lambda_val, points, preds, sample_idx, loss, nbr_idx, minmax, scratch_, nullptr, stream_);
}

extern "C" int lsx_cls3d_forward_tree(int32_t N, int32_t C, int32_t S, int32_t k, float lambda_val, const float* points,
                                      const float* preds, const int32_t* sample_idx, float* loss, int32_t* nbr_idx,
                                      float* minmax, void* scratch_, const void* tree, void* stream_) {
    if (!tree) {
        set_error("lsx_cls3d_forward_tree: null tree");
        return -1;
    }
    return cls3d_forward_impl(N, C, S, k, lambda_val, points, preds, sample_idx, loss, nbr_idx, minmax, scratch_, tree, stream_);
}

extern "C" int lsx_cls3d_backward(int32_t N, int32_t C, int32_t S, int32_t k, float lambda_val, const float* preds,
                                  const int32_t* sample_idx, const int32_t* nbr_idx, const float* minmax,
                                  const float* upstream, float* dL_dpreds, void* scratch_, void* stream_) {
    if (bad_sizes(N, C, S, k)) {
        set_error("lsx_cls3d_backward: bad sizes (N=%d C=%d S=%d k=%d)", N, C, S, k);
        return -1;
    }
    if (!preds || !sample_idx || !nbr_idx || !minmax || !dL_dpreds || !scratch_) {
        set_error("lsx_cls3d_backward: null argument");
        return -1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    char* scratch = static_cast<char*>(scratch_);
    const Cls3dLayout L = make_layout(N, S, k);
    const long long n = (long long)N * C;
    double* acc = reinterpret_cast<double*>(scratch + L.acc_off);
    uint32_t* counts = reinterpret_cast<uint32_t*>(acc + 3);
    LSX_CUDA_OK(cudaMemsetAsync(dL_dpreds, 0, (size_t)n * sizeof(float), stream));
    LSX_CUDA_OK(cudaMemsetAsync(acc, 0, 3 * sizeof(double) + 2 * sizeof(uint32_t), stream));
    const float term_scale = (float)((double)lambda_val / ((double)S * (double)k * (double)C));
    cls3d_scatter_kernel<<<ceil_div(S * C, 256), 256, 0, stream>>>(C, S, k, preds, sample_idx, nbr_idx, minmax, term_scale,
                                                                   dL_dpreds, acc);
    LSX_KERNEL_OK(stream, false);
    const int blocks = dense_blocks(n);
    cls3d_count_ties_kernel<<<blocks, 256, 0, stream>>>(n, preds, minmax, counts);
    LSX_KERNEL_OK(stream, false);
    cls3d_finalize_grad_kernel<<<blocks, 256, 0, stream>>>(n, preds, minmax, acc, counts, upstream, dL_dpreds);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

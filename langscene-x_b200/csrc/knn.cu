// knn.cu — mean squared distance to the 3 nearest neighbours (distCUDA2).
//
// Reference behaviour restated: simple-knn/simple_knn.cu:185-221
//   scene AABB (min/max reductions whose init value {0,0,0} forces the origin into the box, :191-199),
//   30-bit Morton codes (:45-61), sort by code (:206-213), then per point the exact 3 smallest squared
//   distances  d.x*d.x + d.y*d.y + d.z*d.z  to all OTHER points (self excluded by position, :131-183),
//   result (b0 + b1 + b2) / 3.0f with b0 <= b1 <= b2 (FLT_MAX entries when P < 4).
//   The reference prunes with one level of 1024-point boxes and lets every thread scan all boxes.
//
// B200 design: same Morton ordering, but
//   * the sorted points are gathered once into a float4 stream (coalesced, 16-B loads),
//   * a 3-level hierarchy of AABBs with fan-out 32 (leaf = 32 consecutive Morton points),
//   * one WARP per leaf: its 32 lanes are the 32 queries; box tests are lane-parallel (lane i tests
//     child i, one ballot selects the survivors), candidate leaves are staged in shared memory and
//     scanned by all lanes with broadcast LDS.128; the pruning bound is the warp maximum of the
//     per-lane 3rd-best distance (REDUX), widened by 1e-5 so float rounding can never drop a neighbour;
//     before a surviving leaf is scanned every lane tests its OWN point against the leaf's box and its
//     OWN 3rd-best distance, and the scan happens only if some lane votes for it (4x fewer scans).
//   The result is the exact 3-NN set, evaluated with the reference's distance expression, so the
//   output is bit-identical; no host synchronisation (the reference has two).
#include <cfloat>
#include <climits>

#include "kernels.cuh"

namespace lsx {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kFan = 32;

struct Box {
    float4 lo, hi;  // .w unused
};

__device__ __forceinline__ uint32_t spread_bits10(uint32_t x) {
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

// ---- scene bounds ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bounds_partial_kernel(int P, const float* __restrict__ pts, float* __restrict__ part) {
    float lo[3] = {0.f, 0.f, 0.f}, hi[3] = {0.f, 0.f, 0.f};  // init {0,0,0}: the origin is always inside
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = pts[3 * (size_t)i + a];
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    }
    __shared__ float s[8][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(kFull, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(kFull, hi[a], o));
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            s[warp][a] = lo[a];
            s[warp][3 + a] = hi[a];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fminf(v, s[w][threadIdx.x]) : fmaxf(v, s[w][threadIdx.x]);
        part[blockIdx.x * 6 + threadIdx.x] = v;
    }
}

__global__ void bounds_final_kernel(int nb, const float* __restrict__ part, float* __restrict__ bounds) {
    if (threadIdx.x < 6) {
        float v = part[threadIdx.x];
        for (int b = 1; b < nb; ++b) v = threadIdx.x < 3 ? fminf(v, part[b * 6 + threadIdx.x]) : fmaxf(v, part[b * 6 + threadIdx.x]);
        bounds[threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(256) morton_kernel(int P, const float* __restrict__ pts, const float* __restrict__ bounds,
                                                     uint32_t* __restrict__ codes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float3 lo = make_float3(bounds[0], bounds[1], bounds[2]);
    const float3 hi = make_float3(bounds[3], bounds[4], bounds[5]);
    const float3 c = make_float3(pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2]);
    const uint32_t x = spread_bits10(((c.x - lo.x) / (hi.x - lo.x)) * ((1 << 10) - 1));
    const uint32_t y = spread_bits10(((c.y - lo.y) / (hi.y - lo.y)) * ((1 << 10) - 1));
    const uint32_t z = spread_bits10(((c.z - lo.z) / (hi.z - lo.z)) * ((1 << 10) - 1));
    codes[i] = x | (y << 1) | (z << 2);
}

// `rank` (optional): position of every original point in the sorted stream (the K-NN queries start from their own leaf)
__global__ void __launch_bounds__(256) gather_points_kernel(int P, const float* __restrict__ pts,
                                                            const uint32_t* __restrict__ order, float4* __restrict__ spts,
                                                            uint32_t* __restrict__ rank) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const uint32_t i = order[k];
    spts[k] = make_float4(pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2], __uint_as_float(i));
    if (rank != nullptr) rank[i] = (uint32_t)k;
}

// ---- box hierarchy -------------------------------------------------------------------------------------
// level 0: one warp reduces 32 consecutive points; higher levels: one warp reduces 32 child boxes
__global__ void __launch_bounds__(256) build_boxes_kernel(int n_children, const float4* __restrict__ pts,
                                                          const Box* __restrict__ child_boxes, Box* __restrict__ out,
                                                          int n_out) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_out) return;
    const int c = w * kFan + lane;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (c < n_children) {
        if (pts) {
            const float4 p = pts[c];
            lo[0] = hi[0] = p.x;
            lo[1] = hi[1] = p.y;
            lo[2] = hi[2] = p.z;
        } else {
            const Box b = child_boxes[c];
            lo[0] = b.lo.x; lo[1] = b.lo.y; lo[2] = b.lo.z;
            hi[0] = b.hi.x; hi[1] = b.hi.y; hi[2] = b.hi.z;
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(kFull, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(kFull, hi[a], o));
        }
    }
    if (lane == 0) {
        Box b;
        b.lo = make_float4(lo[0], lo[1], lo[2], 0.f);
        b.hi = make_float4(hi[0], hi[1], hi[2], 0.f);
        out[w] = b;
    }
}

__device__ __forceinline__ float box_box_dist2(const Box& a, const Box& b) {
    const float gx = fmaxf(0.f, fmaxf(a.lo.x - b.hi.x, b.lo.x - a.hi.x));
    const float gy = fmaxf(0.f, fmaxf(a.lo.y - b.hi.y, b.lo.y - a.hi.y));
    const float gz = fmaxf(0.f, fmaxf(a.lo.z - b.hi.z, b.lo.z - a.hi.z));
    return gx * gx + gy * gy + gz * gz;
}

__device__ __forceinline__ void insert3(float (&best)[3], float dist) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (best[j] > dist) {
            const float t = best[j];
            best[j] = dist;
            dist = t;
        }
    }
}

constexpr int kQueryWarps = 8;

__global__ void __launch_bounds__(kQueryWarps * 32) knn_query_kernel(int P, const float4* __restrict__ spts,
                                                                     const Box* __restrict__ l1, int n1,
                                                                     const Box* __restrict__ l2, int n2,
                                                                     const Box* __restrict__ l3, int n3,
                                                                     float* __restrict__ out) {
    __shared__ float4 s_pts[kQueryWarps][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int leaf = blockIdx.x * kQueryWarps + warp;
    if (leaf >= n1) return;

    const int qi = leaf * kFan + lane;
    const bool valid = qi < P;
    const float4 q = valid ? spts[qi] : make_float4(0.f, 0.f, 0.f, 0.f);
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};

    auto scan_leaf = [&](int cl, bool is_self) {
        const int base = cl * kFan;
        const int cnt = min(kFan, P - base);
        __syncwarp();
        if (lane < cnt) s_pts[warp][lane] = spts[base + lane];
        __syncwarp();
        if (valid) {
            for (int k = 0; k < cnt; ++k) {
                if (is_self && k == lane) continue;
                const float4 c = s_pts[warp][k];
                const float dx = c.x - q.x, dy = c.y - q.y, dz = c.z - q.z;
                // operation order of the reference's SASS: fma(dz,dz, fma(dx,dx, dy*dy))
                insert3(best, __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
            }
        }
    };

    // seed the bound from this leaf and its Morton neighbours
    const int seed_lo = max(0, leaf - 1), seed_hi = min(n1 - 1, leaf + 1);
    for (int cl = seed_lo; cl <= seed_hi; ++cl) scan_leaf(cl, cl == leaf);

    const Box qbox = l1[leaf];
    // warp bound: the largest 3rd-best distance of any valid lane (non-negative floats order like uints)
    auto warp_bound = [&]() -> float {
        const unsigned bits = __reduce_max_sync(kFull, valid ? __float_as_uint(best[2]) : 0u);
        return __uint_as_float(bits);
    };
    float bound = warp_bound();
    const float kSlack = 1.0f - 1e-5f;

    for (int c3 = 0; c3 < n3; c3 += 32) {
        const int i3 = c3 + lane;
        bool hit3 = false;
        if (i3 < n3) hit3 = !(box_box_dist2(qbox, l3[i3]) * kSlack > bound);
        unsigned m3 = __ballot_sync(kFull, hit3);
        while (m3) {
            const int b3 = c3 + __ffs(m3) - 1;
            m3 &= m3 - 1;
            const int i2 = b3 * kFan + lane;
            bool hit2 = false;
            if (i2 < n2) hit2 = !(box_box_dist2(qbox, l2[i2]) * kSlack > bound);
            unsigned m2 = __ballot_sync(kFull, hit2);
            while (m2) {
                const int b2 = b3 * kFan + __ffs(m2) - 1;
                m2 &= m2 - 1;
                const int i1 = b2 * kFan + lane;
                bool hit1 = false;
                if (i1 < n1 && (i1 < seed_lo || i1 > seed_hi)) hit1 = !(box_box_dist2(qbox, l1[i1]) * kSlack > bound);
                unsigned m1 = __ballot_sync(kFull, hit1);
                while (m1) {
                    const int b1 = b2 * kFan + __ffs(m1) - 1;
                    m1 &= m1 - 1;
                    // the bound may have tightened since the ballot: re-test this leaf (warp-uniform) ...
                    const Box cb = l1[b1];
                    if (box_box_dist2(qbox, cb) * kSlack > bound) continue;
                    // ... and per query: the leaf is scanned only if SOME lane's own point is closer to its box than
                    // that lane's current 3rd-best distance (point-to-box >= box-to-box, so this prunes far more)
                    const float pgx = fmaxf(0.f, fmaxf(cb.lo.x - q.x, q.x - cb.hi.x));
                    const float pgy = fmaxf(0.f, fmaxf(cb.lo.y - q.y, q.y - cb.hi.y));
                    const float pgz = fmaxf(0.f, fmaxf(cb.lo.z - q.z, q.z - cb.hi.z));
                    const bool need = valid && !((pgx * pgx + pgy * pgy + pgz * pgz) * kSlack > best[2]);
                    if (!__any_sync(kFull, need)) continue;
                    scan_leaf(b1, false);
                    bound = warp_bound();
                }
            }
        }
    }
    if (valid) out[__float_as_uint(q.w)] = (best[0] + best[1] + best[2]) / 3.0f;
}

// ---- K nearest points of arbitrary dataset points (loss_cls_3d's neighbour search) -------------------------------------
// One CTA (8 warps) per query, in two phases around a FIXED pruning bound:
//   seed     warp 0 scans the 32 leaves (one leaf per lane) of the level-2 node that holds the query's own position in the Morton
//            order and publishes the K best (distance, index) pairs; the K-th is the bound of everything that follows;
//   collect  the warps share the level-3 nodes; children are pruned lane-parallel by their point-to-box distance against the
//            bound (kept when equal: an equally distant point may win the tie; same slack as the 3-NN kernel) and the surviving
//            LEAVES are appended to a queue in shared memory — no leaf is scanned yet, nothing depends on a scan's result;
//   scan     the queue is scanned 32 leaves at a time, one leaf per lane into that lane's private sorted K-list (16-B loads, four
//            in flight, candidates pre-filtered by the bound); the lists of a warp are merged once, the 8 warp lists and the seed
//            list once more.
// The bounding boxes of consecutive Morton ranges overlap (a range that straddles a high-order boundary of the curve is
// large), so a query's ball reaches ~30 level-2 nodes but only a few leaves in each: scanning node by node (rounds 1-2) spent
// a 32-leaf scan and a K-round merge on every one of them, 125 k warp instructions per query.  Pairs are ordered
// lexicographically (ties towards the lower index, like the brute-force scan of cls3d.cu) and distances use that kernel's
// expression, so both routes return the same bits.
__device__ __forceinline__ float point_box_dist2(const float qx, const float qy, const float qz, const Box& b) {
    const float gx = fmaxf(0.f, fmaxf(b.lo.x - qx, qx - b.hi.x));
    const float gy = fmaxf(0.f, fmaxf(b.lo.y - qy, qy - b.hi.y));
    const float gz = fmaxf(0.f, fmaxf(b.lo.z - qz, qz - b.hi.z));
    return gx * gx + gy * gy + gz * gz;
}

__device__ __forceinline__ bool pair_less(const float d0, const int i0, const float d1, const int i1) {
    return d0 < d1 || (d0 == d1 && i0 < i1);
}

// insert (d, id) into a lexicographically sorted K-list that it beats
template <int K>
__device__ __forceinline__ void pair_insert(float (&bd)[K], int (&bi)[K], const float d, const int id) {
#pragma unroll
    for (int j = K - 1; j > 0; --j) {
        const bool shift = pair_less(d, id, bd[j - 1], bi[j - 1]);
        const bool here = !shift && pair_less(d, id, bd[j], bi[j]);
        if (shift) {
            bd[j] = bd[j - 1];
            bi[j] = bi[j - 1];
        } else if (here) {
            bd[j] = d;
            bi[j] = id;
        }
    }
    if (pair_less(d, id, bd[0], bi[0])) {
        bd[0] = d;
        bi[0] = id;
    }
}

// the K smallest pairs of the lanes' sorted lists (ld, li), in order, to `out_d / out_i` (lane 0 writes): K rounds of a
// lexicographic warp minimum over the lists' heads; every lane whose head IS the winner pops it (equal pairs are the same
// point, so this also drops duplicates)
template <int K, typename OutD, typename OutI>
__device__ __forceinline__ void merge_lists_to(float (&ld)[K], int (&li)[K], OutD out_d, OutI out_i) {
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < K; ++r) {
        float hd = ld[0];
        int hi = li[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(kFull, hd, o);
            const int oi = __shfl_xor_sync(kFull, hi, o);
            if (pair_less(od, oi, hd, hi)) {
                hd = od;
                hi = oi;
            }
        }
        if (hi == li[0] && hd == ld[0]) {
#pragma unroll
            for (int j = 0; j + 1 < K; ++j) {
                ld[j] = ld[j + 1];
                li[j] = li[j + 1];
            }
            ld[K - 1] = FLT_MAX;
            li[K - 1] = INT_MAX;
        }
        if (lane == 0) {
            out_d[r] = hd;
            out_i[r] = hi;
        }
    }
}

constexpr int kLeafQueue = 1024;  // queued leaves per query; a warp that finds the queue full scans its leaves at once

template <int K>
__global__ void __launch_bounds__(kQueryWarps * 32) knn_tree_query_kernel(int P, const float4* __restrict__ spts,
                                                                          const Box* __restrict__ l1, int n1,
                                                                          const Box* __restrict__ l2, int n2,
                                                                          const Box* __restrict__ l3, int n3, int S,
                                                                          const float* __restrict__ points,
                                                                          const uint32_t* __restrict__ rank,
                                                                          const int* __restrict__ sample_idx,
                                                                          float* __restrict__ cand_d, int* __restrict__ cand_i) {
    __shared__ float s_md[kQueryWarps + 1][K];  // row kQueryWarps: the seed list
    __shared__ int s_mi[kQueryWarps + 1][K];
    __shared__ int s_queue[kLeafQueue];
    __shared__ int s_qn, s_qvalid, s_seed2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x;
    if (s >= S) return;
    const int si = sample_idx[s];
    LSX_CHECK_INDEX(si, P, "sampled point index");
    const float qx = points[3 * (size_t)si], qy = points[3 * (size_t)si + 1], qz = points[3 * (size_t)si + 2];
    float ld[K];  // this lane's private candidates
    int li[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        ld[j] = FLT_MAX;
        li[j] = INT_MAX;
    }
    float bound_d = FLT_MAX;  // candidates must beat (bound_d, bound_i): nothing during the seed scan, the seed's K-th afterwards
    int bound_i = INT_MAX;
    // this lane walks the 32 points of ITS leaf (independent 16-B loads, four at a time) into its private list
    auto scan_leaf = [&](const int leaf) {
        LSX_CHECK_INDEX(leaf, n1, "leaf of the box hierarchy");
        const int base = leaf * kFan;
        const int cnt = min(kFan, P - base);
        for (int k0 = 0; k0 < cnt; k0 += 4) {
            float4 c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) c[u] = (k0 + u < cnt) ? __ldg(spts + base + k0 + u) : make_float4(3e18f, 3e18f, 3e18f, 0.f);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float dx = c[u].x - qx, dy = c[u].y - qy, dz = c[u].z - qz;
                const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                const int id = (int)__float_as_uint(c[u].w);
                if (k0 + u < cnt && pair_less(d, id, bound_d, bound_i) && pair_less(d, id, ld[K - 1], li[K - 1]))
                    pair_insert<K>(ld, li, d, id);
            }
        }
    };
    // ---- seed (warp 0) ----
    if (warp == 0) {
        // the level-2 node of the query's own position in the Morton order: its 1024 neighbours along the curve hold most of
        // the true neighbours (the box NEAREST to the query is useless as a seed: the boxes overlap, many contain it)
        LSX_CHECK_INDEX(rank[si], P, "rank of the query in the Morton order");
        const int seed2 = (int)(rank[si] / (uint32_t)(kFan * kFan));
        const int leaf = seed2 * kFan + lane;
        if (leaf < n1) scan_leaf(leaf);
        merge_lists_to<K>(ld, li, s_md[kQueryWarps], s_mi[kQueryWarps]);
        if (lane == 0) {
            s_seed2 = seed2;
            s_qn = 0;
            s_qvalid = kLeafQueue;
        }
        // merge_lists_to consumed the lists' heads; start over for the scan phase
#pragma unroll
        for (int j = 0; j < K; ++j) {
            ld[j] = FLT_MAX;
            li[j] = INT_MAX;
        }
    }
    __syncthreads();
    bound_d = s_md[kQueryWarps][K - 1];
    bound_i = s_mi[kQueryWarps][K - 1];
    const int seed2 = s_seed2;
    // ---- collect: surviving leaves of every level-2 node but the seed's ----
    const float kSlack = 1.0f - 1e-5f;
    for (int i3 = warp; i3 < n3; i3 += kQueryWarps) {
        if (point_box_dist2(qx, qy, qz, l3[i3]) * kSlack > bound_d) continue;  // warp-uniform
        const int i2 = i3 * kFan + lane;
        unsigned m2 = __ballot_sync(kFull, i2 < n2 && i2 != seed2 && !(point_box_dist2(qx, qy, qz, l2[i2 < n2 ? i2 : 0]) * kSlack > bound_d));
        while (m2) {
            const int b2 = i3 * kFan + __ffs(m2) - 1;
            m2 &= m2 - 1;
            const int leaf = b2 * kFan + lane;
            const bool mine = leaf < n1 && !(point_box_dist2(qx, qy, qz, l1[leaf < n1 ? leaf : 0]) * kSlack > bound_d);
            const unsigned mm = __ballot_sync(kFull, mine);
            if (mm == 0u) continue;
            const int cnt = __popc(mm);
            int pos = 0;
            if (lane == 0) pos = atomicAdd(&s_qn, cnt);
            pos = __shfl_sync(kFull, pos, 0);
            if (pos + cnt <= kLeafQueue) {
                if (mine) {
                    LSX_CHECK_INDEX(pos + __popc(mm & ((1u << lane) - 1u)), kLeafQueue, "leaf queue slot");
                    s_queue[pos + __popc(mm & ((1u << lane) - 1u))] = leaf;
                }
            } else {  // queue full (degenerate data: many points at the bound): entries from `pos` on are invalid, scan now
                if (lane == 0) atomicMin(&s_qvalid, pos);
                if (mine) scan_leaf(leaf);
            }
        }
    }
    __syncthreads();
    // ---- scan the queue, one leaf per lane ----
    const int Q = min(s_qn, s_qvalid);
    for (int q0 = warp * 32; q0 < Q; q0 += kQueryWarps * 32)
        if (q0 + lane < Q) scan_leaf(s_queue[q0 + lane]);
    merge_lists_to<K>(ld, li, s_md[warp], s_mi[warp]);
    __syncthreads();
    if (warp != 0) return;
    // ---- the warps' lists and the seed list -> the query's K nearest ----
#pragma unroll
    for (int j = 0; j < K; ++j) {
        ld[j] = lane <= kQueryWarps ? s_md[lane][j] : FLT_MAX;
        li[j] = lane <= kQueryWarps ? s_mi[lane][j] : INT_MAX;
    }
    merge_lists_to<K>(ld, li, cand_d + (size_t)s * K, cand_i + (size_t)s * K);
}

struct KnnScratch {
    float* bounds_part;
    float* bounds;
    uint32_t* codes[2];
    uint32_t* order[2];
    void* sort_temp;
    float4* spts;
    uint32_t* rank;
    Box *l1, *l2, *l3;
    int n1, n2, n3, nb_bounds;
    size_t bytes;
};

KnnScratch carve_knn(char* base, int P) {
    KnnScratch s{};
    size_t off = 0;
    auto take = [&](size_t bytes) -> char* {
        off = align_up(off, 256);
        char* p = base ? base + off : nullptr;
        off += bytes;
        return p;
    };
    s.n1 = ceil_div(P, kFan);
    s.n2 = ceil_div(s.n1, kFan);
    s.n3 = ceil_div(s.n2, kFan);
    s.nb_bounds = min(1024, ceil_div(P, 256));
    s.bounds_part = (float*)take((size_t)s.nb_bounds * 6 * sizeof(float));
    s.bounds = (float*)take(8 * sizeof(float));
    s.codes[0] = (uint32_t*)take((size_t)P * 4);
    s.codes[1] = (uint32_t*)take((size_t)P * 4);
    s.order[0] = (uint32_t*)take((size_t)P * 4);
    s.order[1] = (uint32_t*)take((size_t)P * 4);
    s.sort_temp = take(radix_sort_temp_bytes(P));
    s.spts = (float4*)take((size_t)P * sizeof(float4));
    s.rank = (uint32_t*)take((size_t)P * 4);
    s.l1 = (Box*)take((size_t)s.n1 * sizeof(Box));
    s.l2 = (Box*)take((size_t)s.n2 * sizeof(Box));
    s.l3 = (Box*)take((size_t)s.n3 * sizeof(Box));
    s.bytes = align_up(off, 256);
    return s;
}

}  // namespace

size_t knn_temp_bytes(int P) { return carve_knn(nullptr, P > 0 ? P : 1).bytes; }

// Morton order + float4 stream + 3-level box hierarchy of `points` into `temp` (knn_temp_bytes(P) bytes)
int knn_tree_build(int P, const float* points, void* temp, cudaStream_t stream, bool with_rank) {
    if (P <= 0) return 0;
    KnnScratch s = carve_knn(static_cast<char*>(temp), P);
    bounds_partial_kernel<<<s.nb_bounds, 256, 0, stream>>>(P, points, s.bounds_part);
    LSX_KERNEL_OK(stream, false);
    bounds_final_kernel<<<1, 32, 0, stream>>>(s.nb_bounds, s.bounds_part, s.bounds);
    LSX_KERNEL_OK(stream, false);
    morton_kernel<<<ceil_div(P, 256), 256, 0, stream>>>(P, points, s.bounds, s.codes[0]);
    LSX_KERNEL_OK(stream, false);
    int res = 0;
    int rc = radix_sort_pairs_u32(s.codes, s.order, P, 0, 30, /*identity_vals=*/true, s.sort_temp, &res, stream, false);
    if (rc) return rc;
    gather_points_kernel<<<ceil_div(P, 256), 256, 0, stream>>>(P, points, s.order[res], s.spts, with_rank ? s.rank : nullptr);
    LSX_KERNEL_OK(stream, false);
    build_boxes_kernel<<<ceil_div(s.n1 * 32, 256), 256, 0, stream>>>(P, s.spts, nullptr, s.l1, s.n1);
    LSX_KERNEL_OK(stream, false);
    build_boxes_kernel<<<ceil_div(s.n2 * 32, 256), 256, 0, stream>>>(s.n1, nullptr, s.l1, s.l2, s.n2);
    LSX_KERNEL_OK(stream, false);
    build_boxes_kernel<<<ceil_div(s.n3 * 32, 256), 256, 0, stream>>>(s.n2, nullptr, s.l2, s.l3, s.n3);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

int knn_mean_dist2(int P, const float* points, float* out, void* temp, cudaStream_t stream) {
    if (P <= 0) return 0;
    const int rc = knn_tree_build(P, points, temp, stream, /*with_rank=*/false);
    if (rc) return rc;
    KnnScratch s = carve_knn(static_cast<char*>(temp), P);
    knn_query_kernel<<<ceil_div(s.n1, kQueryWarps), kQueryWarps * 32, 0, stream>>>(P, s.spts, s.l1, s.n1, s.l2, s.n2, s.l3,
                                                                                   s.n3, out);
    LSX_KERNEL_OK(stream, false);
    return 0;
}

// exact K nearest points (the query itself included) of S dataset points, from a tree built by knn_tree_build over the same
// `points`: cand_d / cand_i [S][K] sorted by (distance, index) — the layout cls3d.cu's merge kernel takes for one slice
int knn_tree_query(int K, int P, const void* tree, int S, const float* points, const int* sample_idx, float* cand_d, int* cand_i,
                   cudaStream_t stream) {
    if (P <= 0 || S <= 0) return 0;
    KnnScratch s = carve_knn(const_cast<char*>(static_cast<const char*>(tree)), P);
    const int blocks = S;  // one CTA per query
    switch (K) {
#define LSX_KNN_CASE(KK)                                                                                                       \
    case KK:                                                                                                                   \
        knn_tree_query_kernel<KK><<<blocks, kQueryWarps * 32, 0, stream>>>(P, s.spts, s.l1, s.n1, s.l2, s.n2, s.l3, s.n3, S,   \
                                                                           points, s.rank, sample_idx, cand_d, cand_i);        \
        break;
        LSX_KNN_CASE(1)
        LSX_KNN_CASE(2)
        LSX_KNN_CASE(3)
        LSX_KNN_CASE(4)
        LSX_KNN_CASE(5)
        LSX_KNN_CASE(6)
        LSX_KNN_CASE(7)
        LSX_KNN_CASE(8)
#undef LSX_KNN_CASE
        default:
            set_error("knn_tree_query: k = %d outside 1..8", K);
            return -1;
    }
    LSX_KERNEL_OK(stream, false);
    return 0;
}

}  // namespace lsx

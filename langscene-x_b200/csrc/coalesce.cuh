// coalesce.cuh — block-cooperative, fully coalesced movement of row-major (rows x row_floats) slabs between
// global memory and a padded shared-memory tile.
//
// The per-Gaussian kernels (preprocess forward / backward) own one Gaussian per thread, but the reference's
// tensors are array-of-structures: (P, M, 3) SH coefficients, (P, F) features, packed records ...  A thread
// walking its own 192-B row makes every warp-level access touch 32 different cache lines.  Instead the block's
// slab (contiguous in global memory) is moved with 16-B accesses in thread order — each warp instruction
// covers 512 contiguous bytes — and the threads then work on their rows in shared memory, whose row stride is
// chosen odd (in words) so that thread-per-row access is bank-conflict free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lsx {

// smem row stride (floats) for thread-per-row scalar access: odd => conflict free
__host__ __device__ static inline int padded_row(int row_floats) { return row_floats | 1; }

// global (rows x rowf, contiguous) -> smem (row stride srow).  `g` points at the slab's first element.
// U (default four) independent 16-B loads are issued per thread before any of them is consumed: these kernels are bound by
// memory latency (few resident warps because of the tile), so bytes in flight per thread is what matters.
template <int NT, int U = 4>
__device__ __forceinline__ void slab_load(float* __restrict__ s, const float* __restrict__ g, int rows, int rowf, int srow) {
    const int total = rows * rowf;
    const int tid = threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0 && rowf >= 4) {
        const int n4 = total >> 2;
        for (int i0 = tid; i0 < n4; i0 += NT * U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * NT;
                v[u] = (i < n4) ? __ldg(reinterpret_cast<const float4*>(g) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * NT;
                if (i < n4) {
                    const int e = i * 4;
                    const int r = e / rowf, c = e - r * rowf;
                    const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        int ck = c + k, rk = r;
                        if (ck >= rowf) {
                            ck -= rowf;
                            ++rk;
                        }
                        s[rk * srow + ck] = vv[k];
                    }
                }
            }
        }
        for (int e2 = (n4 << 2) + tid; e2 < total; e2 += NT) {
            const int r2 = e2 / rowf;
            s[r2 * srow + (e2 - r2 * rowf)] = __ldg(g + e2);
        }
    } else {
        for (int e2 = tid; e2 < total; e2 += NT) {
            const int r2 = e2 / rowf;
            s[r2 * srow + (e2 - r2 * rowf)] = __ldg(g + e2);
        }
    }
}

// smem (row stride srow, columns [col0, col0 + rowf)) -> global (rows x rowf, contiguous);
// accumulate: g += s instead of g = s
template <int NT>
__device__ __forceinline__ void slab_store(float* __restrict__ g, const float* __restrict__ s, int rows, int rowf, int srow,
                                           int col0 = 0, bool accumulate = false) {
    const int total = rows * rowf;
    const int tid = threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0 && rowf >= 4) {
        const int n4 = total >> 2;
        const int step = NT * 4;
        const int dq = step / rowf, dr = step - dq * rowf;
        int e = tid * 4;
        int r = e / rowf, c = e - r * rowf;
        for (int i = tid; i < n4; i += NT) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (accumulate) o = reinterpret_cast<const float4*>(g)[i];  // issued first: overlaps the shared-memory reads
            float vv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int ck = c + k, rk = r;
                if (ck >= rowf) {
                    ck -= rowf;
                    ++rk;
                }
                vv[k] = s[rk * srow + col0 + ck];
            }
            if (accumulate) {
                vv[0] += o.x;
                vv[1] += o.y;
                vv[2] += o.z;
                vv[3] += o.w;
            }
            reinterpret_cast<float4*>(g)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
            r += dq;
            c += dr;
            if (c >= rowf) {
                c -= rowf;
                ++r;
            }
        }
        for (int e2 = (n4 << 2) + tid; e2 < total; e2 += NT) {
            const int r2 = e2 / rowf;
            const float v = s[r2 * srow + col0 + (e2 - r2 * rowf)];
            g[e2] = accumulate ? g[e2] + v : v;
        }
    } else {
        for (int e2 = tid; e2 < total; e2 += NT) {
            const int r2 = e2 / rowf;
            const float v = s[r2 * srow + col0 + (e2 - r2 * rowf)];
            g[e2] = accumulate ? g[e2] + v : v;
        }
    }
}

// global -> global: dst (rows x rowf, contiguous) = columns [col0, col0 + rowf) of src rows of `src_stride` floats
// (accumulate: dst += ...).  Stores are coalesced 16-B accesses in thread order; the gathers walk contiguous runs of a
// row, so the sectors one load instruction touches are used up by the next three.
template <int NT>
__device__ __forceinline__ void slab_copy_columns(float* __restrict__ dst, const float* __restrict__ src, int rows, int rowf,
                                                  int src_stride, int col0, bool accumulate) {
    const int total = rows * rowf;
    const int tid = threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && rowf >= 4) {
        const int n4 = total >> 2;
        const int step = NT * 4;
        const int dq = step / rowf, dr = step - dq * rowf;
        int e = tid * 4;
        int r = e / rowf, c = e - r * rowf;
        for (int i = tid; i < n4; i += NT) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (accumulate) o = reinterpret_cast<const float4*>(dst)[i];
            float vv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int ck = c + k, rk = r;
                if (ck >= rowf) {
                    ck -= rowf;
                    ++rk;
                }
                vv[k] = __ldg(src + (size_t)rk * src_stride + col0 + ck);
            }
            reinterpret_cast<float4*>(dst)[i] = make_float4(vv[0] + o.x, vv[1] + o.y, vv[2] + o.z, vv[3] + o.w);
            r += dq;
            c += dr;
            if (c >= rowf) {
                c -= rowf;
                ++r;
            }
        }
        for (int e2 = (n4 << 2) + tid; e2 < total; e2 += NT) {
            const int r2 = e2 / rowf;
            const float v = __ldg(src + (size_t)r2 * src_stride + col0 + (e2 - r2 * rowf));
            dst[e2] = accumulate ? dst[e2] + v : v;
        }
    } else {
        for (int e2 = tid; e2 < total; e2 += NT) {
            const int r2 = e2 / rowf;
            const float v = __ldg(src + (size_t)r2 * src_stride + col0 + (e2 - r2 * rowf));
            dst[e2] = accumulate ? dst[e2] + v : v;
        }
    }
}

}  // namespace lsx

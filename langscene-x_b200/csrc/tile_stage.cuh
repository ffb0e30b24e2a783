// tile_stage.cuh — per-warp, double-buffered staging of a tile's Gaussian list into shared memory.
//
// The render kernels run ONE WARP PER CTA: warp = one 8x4 pixel block of a 16x16 tile.  A round covers CHUNK
// consecutive entries of the tile's sorted list: each lane reads one entry's sub-tile footprint mask (cull.cu);
// a ballot of "my block's bit is set" is the warp's work list for the round, and every surviving lane issues ONE
// bulk copy (TMA engine, cp.async.bulk -> UBLKCP) of its Gaussian's packed, sector-aligned record into the next
// free slot of the round's buffer, signalling the buffer's mbarrier with the byte count.  Two buffers alternate,
// so the copies of round r+1 are in flight while round r is blended; there is no CTA-wide barrier anywhere.
// Consumers address the buffers through 32-bit shared-window addresses (ld.shared with register + immediate).
#pragma once
#include "async_copy.cuh"
#include "common.cuh"

namespace lsx {

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
// N consecutive 16-B shared loads issued back to back from ONE asm block, so that their latencies overlap
// (separate volatile asm statements are kept in program order with their consumers in between).
template <int N>
__device__ __forceinline__ void lds128xN(uint32_t addr, float4* f) {
    static_assert(N >= 1 && N <= 4, "group size");
    if constexpr (N == 1) {
        f[0] = lds128(addr);
    } else if constexpr (N == 2) {
        asm volatile(
            "ld.shared.v4.f32 {%0, %1, %2, %3}, [%8];\n\t"
            "ld.shared.v4.f32 {%4, %5, %6, %7}, [%8+16];"
            : "=f"(f[0].x), "=f"(f[0].y), "=f"(f[0].z), "=f"(f[0].w), "=f"(f[1].x), "=f"(f[1].y), "=f"(f[1].z), "=f"(f[1].w)
            : "r"(addr)
            : "memory");
    } else if constexpr (N == 3) {
        asm volatile(
            "ld.shared.v4.f32 {%0, %1, %2, %3}, [%12];\n\t"
            "ld.shared.v4.f32 {%4, %5, %6, %7}, [%12+16];\n\t"
            "ld.shared.v4.f32 {%8, %9, %10, %11}, [%12+32];"
            : "=f"(f[0].x), "=f"(f[0].y), "=f"(f[0].z), "=f"(f[0].w), "=f"(f[1].x), "=f"(f[1].y), "=f"(f[1].z), "=f"(f[1].w),
              "=f"(f[2].x), "=f"(f[2].y), "=f"(f[2].z), "=f"(f[2].w)
            : "r"(addr)
            : "memory");
    } else {
        asm volatile(
            "ld.shared.v4.f32 {%0, %1, %2, %3}, [%16];\n\t"
            "ld.shared.v4.f32 {%4, %5, %6, %7}, [%16+16];\n\t"
            "ld.shared.v4.f32 {%8, %9, %10, %11}, [%16+32];\n\t"
            "ld.shared.v4.f32 {%12, %13, %14, %15}, [%16+48];"
            : "=f"(f[0].x), "=f"(f[0].y), "=f"(f[0].z), "=f"(f[0].w), "=f"(f[1].x), "=f"(f[1].y), "=f"(f[1].z), "=f"(f[1].w),
              "=f"(f[2].x), "=f"(f[2].y), "=f"(f[2].z), "=f"(f[2].w), "=f"(f[3].x), "=f"(f[3].y), "=f"(f[3].z), "=f"(f[3].w)
            : "r"(addr)
            : "memory");
    }
}
// all NQ float4 of a row, in groups of GROUP loads
template <int NQ, int GROUP = 4>
__device__ __forceinline__ void lds_row(uint32_t addr, float4* f) {
#pragma unroll
    for (int q = 0; q + GROUP <= NQ; q += GROUP) lds128xN<GROUP>(addr + q * 16, f + q);
    if constexpr (NQ % GROUP != 0) lds128xN<NQ % GROUP>(addr + (NQ / GROUP) * GROUP * 16, f + (NQ / GROUP) * GROUP);
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ int lds32i(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

constexpr unsigned kFullMask = 0xffffffffu;

template <int RS, int CHUNK>  // record stride in floats; list entries per round (<= 32)
struct WarpStage {
    static_assert(CHUNK == 8 || CHUNK == 16 || CHUNK == 32, "CHUNK must be 8, 16 or 32");
    static constexpr int kRecBytes = RS * 4;
    static constexpr int kBufBytes = CHUNK * kRecBytes;          // one record buffer
    static constexpr int kIdsOff = 2 * kBufBytes;                // int ids [2][CHUNK]
    static constexpr int kBarOff = kIdsOff + 2 * CHUNK * 4;      // u64 bar [2]
    static constexpr size_t kSmemBytes = kBarOff + 2 * sizeof(uint64_t);

    unsigned char* base;  // generic pointer (producer side)
    uint32_t sbase;       // shared-window address of `base` (consumer side)

    __device__ __forceinline__ uint64_t* bar(int buf) const { return reinterpret_cast<uint64_t*>(base + kBarOff) + buf; }

    // one warp per CTA: all 32 threads call this
    __device__ __forceinline__ void init(unsigned char* smem) {
        base = smem;
        sbase = smem_u32(smem);
        if (threadIdx.x == 0) {
            mbar_init(bar(0), 1);
            mbar_init(bar(1), 1);
            mbar_init_fence();
        }
        __syncwarp();
    }

    // Staging is software-pipelined: prefetch() loads the footprint mask and the Gaussian index of this lane's
    // candidate one round ahead (plain global loads, latency off the critical path); issue() consumes the
    // prefetched pair, stages the surviving entries of the round into buffer `buf` and returns the ballot of
    // surviving lanes: survivor number k (in lane order) occupies slot k.
    // `entry` = absolute position in point_list / masks of this lane's candidate, or -1 (lanes >= CHUNK and
    // positions outside the list).
    unsigned pre_mask;
    int pre_id;
    __device__ __forceinline__ void prefetch(long long entry, const uint32_t* __restrict__ point_list,
                                             const uint8_t* __restrict__ masks) {
        pre_mask = 0u;
        pre_id = 0;
        if (entry >= 0) {
            pre_mask = __ldg(masks + entry);
            pre_id = (int)__ldg(point_list + entry);
        }
    }
    __device__ __forceinline__ unsigned issue(int buf, int block_bit, const float* __restrict__ records) {
        const unsigned lane = threadIdx.x;
        const bool mine = ((pre_mask >> block_bit) & 1u) != 0u;
        const unsigned bits = __ballot_sync(kFullMask, mine);
        if (mine) {
            const int slot = __popc(bits & ((1u << lane) - 1u));
            reinterpret_cast<int*>(base + kIdsOff)[buf * CHUNK + slot] = pre_id;
            bulk_copy_g2s(base + (size_t)buf * kBufBytes + (size_t)slot * kRecBytes, records + (size_t)pre_id * RS,
                          kRecBytes, bar(buf));
        }
        if (lane == 0) mbar_arrive_expect_tx(bar(buf), (uint32_t)__popc(bits) * kRecBytes);
        __syncwarp();  // ids visible to the whole warp
        return bits;
    }

    __device__ __forceinline__ void wait(int buf, uint32_t parity) { mbar_wait(bar(buf), parity); }
    __device__ __forceinline__ uint32_t rec_addr(int buf) const { return sbase + (uint32_t)(buf * kBufBytes); }
    __device__ __forceinline__ uint32_t ids_addr(int buf) const { return sbase + (uint32_t)(kIdsOff + buf * CHUNK * 4); }
};

}  // namespace lsx

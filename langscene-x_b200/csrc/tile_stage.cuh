// tile_stage.cuh — double-buffered staging of a tile's Gaussian list into shared memory.
//
// A batch is STAGE_BATCH list entries.  For each entry one thread reads the Gaussian index and the
// sub-tile footprint mask (cull.cu) from the sorted list and issues ONE bulk copy (TMA engine,
// cp.async.bulk -> UBLKCP) of that Gaussian's packed, sector-aligned record into the batch buffer,
// signalling an mbarrier with the byte count.  Two buffers alternate so the copies of batch b+1 are in
// flight while batch b is blended.  Consumers address the buffers through 32-bit shared-window
// addresses (ld.shared with register + immediate operands, no generic-pointer arithmetic in the loops).
#pragma once
#include "async_copy.cuh"
#include "common.cuh"

namespace lsx {

constexpr int STAGE_BATCH = 128;

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
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
__device__ __forceinline__ unsigned lds8u(uint32_t addr) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

template <int RS>  // record stride in floats
struct TileStage {
    static constexpr int kRecBytes = RS * 4;
    static constexpr int kRecBuf = STAGE_BATCH * kRecBytes;                 // bytes of one record buffer
    static constexpr int kIdsOff = 2 * kRecBuf;                             // int   ids  [2][STAGE_BATCH]
    static constexpr int kMaskOff = kIdsOff + 2 * STAGE_BATCH * 4;          // uint8 mask [2][STAGE_BATCH]
    static constexpr int kBarOff = kMaskOff + 2 * STAGE_BATCH;              // u64   bar  [2]
    static constexpr size_t kSmemBytes = kBarOff + 2 * sizeof(uint64_t);

    unsigned char* base;  // generic pointer (producer side)
    uint32_t sbase;       // shared-window address of `base` (consumer side)

    __device__ __forceinline__ uint64_t* bar(int buf) const { return reinterpret_cast<uint64_t*>(base + kBarOff) + buf; }

    __device__ __forceinline__ void init(unsigned char* smem) {
        base = smem;
        sbase = smem_u32(smem);
        if (threadIdx.x == 0) {
            mbar_init(bar(0), STAGE_BATCH);
            mbar_init(bar(1), STAGE_BATCH);
            mbar_init_fence();
        }
        __syncthreads();
    }

    // Issue the copies of batch b.  `entry` = position in point_list for this thread's slot, or -1.
    // Must be called by (at least) threads 0..STAGE_BATCH-1 of the block, after every thread has finished
    // reading the buffer (b & 1) from batch b-2 (i.e. after a __syncthreads).
    __device__ __forceinline__ void issue(int b, long long entry, const uint32_t* __restrict__ point_list,
                                          const uint8_t* __restrict__ masks, const float* __restrict__ records) {
        if (threadIdx.x < STAGE_BATCH) {
            const int buf = b & 1;
            const int slot = buf * STAGE_BATCH + threadIdx.x;
            if (entry >= 0) {
                const int id = (int)__ldg(point_list + entry);
                reinterpret_cast<int*>(base + kIdsOff)[slot] = id;
                (base + kMaskOff)[slot] = __ldg(masks + entry);
                mbar_arrive_expect_tx(bar(buf), kRecBytes);
                bulk_copy_g2s(base + (size_t)slot * kRecBytes, records + (size_t)id * RS, kRecBytes, bar(buf));
            } else {
                (base + kMaskOff)[slot] = 0;
                mbar_arrive(bar(buf));
            }
        }
    }

    __device__ __forceinline__ void wait(int b) { mbar_wait(bar(b & 1), (uint32_t)((b >> 1) & 1)); }
    // shared-window addresses of batch b's arrays
    __device__ __forceinline__ uint32_t rec_addr(int b) const { return sbase + (uint32_t)((b & 1) * kRecBuf); }
    __device__ __forceinline__ uint32_t ids_addr(int b) const { return sbase + (uint32_t)(kIdsOff + (b & 1) * STAGE_BATCH * 4); }
    __device__ __forceinline__ uint32_t mask_addr(int b) const { return sbase + (uint32_t)(kMaskOff + (b & 1) * STAGE_BATCH); }
};

}  // namespace lsx

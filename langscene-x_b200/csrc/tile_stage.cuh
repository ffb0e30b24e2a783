// tile_stage.cuh — double-buffered staging of a tile's Gaussian list into shared memory.
//
// A batch is STAGE_BATCH list entries.  For each entry one thread reads the Gaussian index from the
// sorted point list and issues ONE bulk copy (TMA engine) of that Gaussian's packed, sector-aligned
// record into the batch buffer, signalling an mbarrier with the byte count.  Two buffers alternate so
// the copies of batch b+1 are in flight while batch b is blended.
#pragma once
#include "async_copy.cuh"
#include "common.cuh"

namespace lsx {

constexpr int STAGE_BATCH = 128;

template <int RS>  // record stride in floats
struct TileStage {
    static constexpr int kRecBytes = RS * 4;
    static constexpr size_t kSmemBytes = 2 * (size_t)STAGE_BATCH * (kRecBytes + 4) + 2 * sizeof(uint64_t);

    float* rec;     // [2][STAGE_BATCH][RS]
    int* ids;       // [2][STAGE_BATCH]
    uint64_t* bar;  // [2]

    __device__ __forceinline__ void init(unsigned char* smem) {
        rec = reinterpret_cast<float*>(smem);
        ids = reinterpret_cast<int*>(rec + 2 * STAGE_BATCH * RS);
        bar = reinterpret_cast<uint64_t*>(ids + 2 * STAGE_BATCH);
        if (threadIdx.x == 0) {
            mbar_init(&bar[0], STAGE_BATCH);
            mbar_init(&bar[1], STAGE_BATCH);
            mbar_init_fence();
        }
        __syncthreads();
    }

    // Issue the copies of batch b.  `entry` = position in point_list for this thread's slot, or -1.
    // Must be called by (at least) threads 0..STAGE_BATCH-1 of the block, after every thread has finished
    // reading the buffer (b & 1) from batch b-2 (i.e. after a __syncthreads).
    __device__ __forceinline__ void issue(int b, long long entry, const uint32_t* __restrict__ point_list,
                                          const float* __restrict__ records) {
        if (threadIdx.x < STAGE_BATCH) {
            const int buf = b & 1;
            if (entry >= 0) {
                const int id = (int)point_list[entry];
                ids[buf * STAGE_BATCH + threadIdx.x] = id;
                mbar_arrive_expect_tx(&bar[buf], kRecBytes);
                bulk_copy_g2s(rec + ((size_t)buf * STAGE_BATCH + threadIdx.x) * RS, records + (size_t)id * RS, kRecBytes,
                              &bar[buf]);
            } else {
                mbar_arrive(&bar[buf]);
            }
        }
    }

    __device__ __forceinline__ void wait(int b) { mbar_wait(&bar[b & 1], (uint32_t)((b >> 1) & 1)); }
    __device__ __forceinline__ const float* rec_buf(int b) const { return rec + (size_t)(b & 1) * STAGE_BATCH * RS; }
    __device__ __forceinline__ const int* id_buf(int b) const { return ids + (b & 1) * STAGE_BATCH; }
};

}  // namespace lsx

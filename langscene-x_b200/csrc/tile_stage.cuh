// tile_stage.cuh — per-warp, double-buffered staging of a block's compacted Gaussian list into shared memory.
//
// The render kernels run ONE WARP PER CTA: warp = one 8x4 pixel block of a 16x16 tile.  cull.cu leaves, per block, the
// compacted list of the tile-list entries whose footprint reaches it; ListStage (below) moves the packed,
// sector-aligned records of 16 consecutive list elements per round into one of two buffers with per-lane asynchronous
// copies, so the copies of round r+1 are in flight while round r is blended; there is no CTA-wide barrier anywhere.
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
__device__ __forceinline__ void sts32i(uint32_t addr, int v) {
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

constexpr unsigned kFullMask = 0xffffffffu;

// ---------------------------------------------------------------------------------------------------------------------
// ListStage — staging from the block's COMPACTED list (cull.cu writes, per 8x4 block of every tile, the positions of the
// list entries whose footprint reaches the block, in depth order).  A round is 16 consecutive elements of that list, all of
// them visited: no mask loads, no ballot / slot bookkeeping, and every slot of the buffer is used.  Records move with
// per-lane 16-B asynchronous copies (two lanes per record), completion is tracked with cp.async groups (one group per
// round, at most two outstanding).  The element -> Gaussian id lookup is a dependent load (list -> point_list), so the
// pipeline runs three rounds ahead for positions and two for ids.
// Traversal index t = 16 q + lane (q = round) maps to list element first + dir * t  (dir = +1 forward, -1 backward).
template <int RS>
struct ListStage {
    static constexpr int CHUNK = 16;
    static constexpr int kRecBytes = RS * 4;
    static constexpr int kBufBytes = CHUNK * kRecBytes;
    static constexpr int kIdsOff = 2 * kBufBytes;  // int ids [2][CHUNK]
    static constexpr size_t kSmemBytes = kIdsOff + 2 * CHUNK * 4;
    static constexpr int kPerLane = RS / 8;  // 16-B pieces per lane
    static_assert(RS % 8 == 0, "record stride must be a multiple of 8 floats");

    uint32_t sbase;
    const uint32_t* list;    // the block's compacted list
    const uint32_t* plist;   // point_list + range.x (positions are relative to the tile's range)
    const float* records;
    int first, dir, count;
    int chk_range, chk_points;  // LSX_BOUNDS_CHECK: length of the tile's range, number of Gaussians
    uint32_t id_next, pos_next;  // id of this lane's element of round cur + 1, position of its element of round cur + 2

    __device__ __forceinline__ uint32_t load_pos(int q) const {
        const int t = q * CHUNK + (int)threadIdx.x;
        if (threadIdx.x < CHUNK && t < count) LSX_CHECK_INDEX(first + dir * t, chk_range, "block list element");
        return (threadIdx.x < CHUNK && t < count) ? __ldg(list + (first + dir * t)) : 0u;
    }
    __device__ __forceinline__ uint32_t load_id(uint32_t pos, int q) const {
        const int t = q * CHUNK + (int)threadIdx.x;
        if (threadIdx.x < CHUNK && t < count) LSX_CHECK_INDEX(pos, chk_range, "tile list position");
        return (threadIdx.x < CHUNK && t < count) ? __ldg(plist + pos) : 0u;
    }
    __device__ __forceinline__ int round_size(int q) const {
        const int left = count - q * CHUNK;
        return left < 0 ? 0 : (left < CHUNK ? left : CHUNK);
    }
    // stage round q (ids of its elements in lanes 0..15) into buffer q & 1 and close the round's copy group
    __device__ __forceinline__ void issue(int q, uint32_t id) {
        const unsigned lane = threadIdx.x;
        const int m = round_size(q);
        const uint32_t buf = (uint32_t)(q & 1);
        if ((int)lane < m) sts32i(sbase + kIdsOff + (buf * CHUNK + lane) * 4u, (int)id);
        const int slot = (int)(lane >> 1);
        const uint32_t id_s = __shfl_sync(kFullMask, id, slot);
        if (slot < m) {
            LSX_CHECK_INDEX(id_s, chk_points, "Gaussian id of a list entry");
            const uint32_t piece = (lane & 1u) * (uint32_t)(kPerLane * 16);
            const char* src = reinterpret_cast<const char*>(records + (size_t)id_s * RS) + piece;
            const uint32_t dst = sbase + buf * (uint32_t)kBufBytes + (uint32_t)slot * (uint32_t)kRecBytes + piece;
#pragma unroll
            for (int j = 0; j < kPerLane; ++j) cp_async16(dst + j * 16, src + j * 16);
        }
        cp_async_commit();
    }
    // all 32 threads; the record buffers must be free (not aliased by live data) from here on
    __device__ __forceinline__ void start(unsigned char* smem, const uint32_t* list_, const uint32_t* plist_,
                                          const float* records_, int first_, int dir_, int count_, int range_len = 0x7fffffff,
                                          int points = 0x7fffffff) {
        chk_range = range_len;
        chk_points = points;
        sbase = smem_u32(smem);
        list = list_;
        plist = plist_;
        records = records_;
        first = first_;
        dir = dir_;
        count = count_;
        const uint32_t p0 = load_pos(0), p1 = load_pos(1);
        pos_next = load_pos(2);
        const uint32_t i0 = load_id(p0, 0);
        id_next = load_id(p1, 1);
        issue(0, i0);
    }
    // top of iteration r: put round r + 1 in flight, advance the look-ahead loads, then wait for round r
    __device__ __forceinline__ void advance(int r) {
        issue(r + 1, id_next);  // an empty group when round r + 1 does not exist
        id_next = load_id(pos_next, r + 2);
        pos_next = load_pos(r + 3);
        cp_async_wait<1>();
        __syncwarp();
    }
    __device__ __forceinline__ void drain() { cp_async_wait<0>(); }
    __device__ __forceinline__ uint32_t rec_addr(int buf) const { return sbase + (uint32_t)(buf * kBufBytes); }
    __device__ __forceinline__ uint32_t ids_addr(int buf) const { return sbase + (uint32_t)(kIdsOff + buf * CHUNK * 4); }
};

}  // namespace lsx

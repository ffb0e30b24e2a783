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

#ifndef LSX_STAGE_MODE
#define LSX_STAGE_MODE 0  // 0: two lanes per record (shipped); 1: lanes on consecutive pieces; 2: one bulk copy per record
#endif

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
// ---- packed fp32 pairs (sm_100: FFMA2 executes two IEEE fp32 FMAs from ONE issue slot) ----------------------------------
// The render kernels are bound by instruction issue, not by the FP32 pipes; their FMA runs act on consecutive channels, which
// sit in consecutive registers anyway.  Each component is rounded exactly like fmaf, so results do not change.
typedef unsigned long long f32x2;  // (lo, hi) floats in an aligned register pair
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(f32x2 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// N consecutive 16-B shared loads as 2 N packed pairs, issued back to back
template <int N>
__device__ __forceinline__ void lds128xN_pairs(uint32_t addr, f32x2* f) {
    static_assert(N >= 1 && N <= 4, "group size");
    if constexpr (N == 1) {
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(f[0]), "=l"(f[1]) : "r"(addr) : "memory");
    } else if constexpr (N == 2) {
        asm volatile(
            "ld.shared.v2.b64 {%0, %1}, [%4];\n\t"
            "ld.shared.v2.b64 {%2, %3}, [%4+16];"
            : "=l"(f[0]), "=l"(f[1]), "=l"(f[2]), "=l"(f[3])
            : "r"(addr)
            : "memory");
    } else if constexpr (N == 3) {
        asm volatile(
            "ld.shared.v2.b64 {%0, %1}, [%6];\n\t"
            "ld.shared.v2.b64 {%2, %3}, [%6+16];\n\t"
            "ld.shared.v2.b64 {%4, %5}, [%6+32];"
            : "=l"(f[0]), "=l"(f[1]), "=l"(f[2]), "=l"(f[3]), "=l"(f[4]), "=l"(f[5])
            : "r"(addr)
            : "memory");
    } else {
        asm volatile(
            "ld.shared.v2.b64 {%0, %1}, [%8];\n\t"
            "ld.shared.v2.b64 {%2, %3}, [%8+16];\n\t"
            "ld.shared.v2.b64 {%4, %5}, [%8+32];\n\t"
            "ld.shared.v2.b64 {%6, %7}, [%8+48];"
            : "=l"(f[0]), "=l"(f[1]), "=l"(f[2]), "=l"(f[3]), "=l"(f[4]), "=l"(f[5]), "=l"(f[6]), "=l"(f[7])
            : "r"(addr)
            : "memory");
    }
}
template <int NQ, int GROUP = 4>
__device__ __forceinline__ void lds_row_pairs(uint32_t addr, f32x2* f) {
#pragma unroll
    for (int q = 0; q + GROUP <= NQ; q += GROUP) lds128xN_pairs<GROUP>(addr + q * 16, f + 2 * q);
    if constexpr (NQ % GROUP != 0) lds128xN_pairs<NQ % GROUP>(addr + (NQ / GROUP) * GROUP * 16, f + 2 * (NQ / GROUP) * GROUP);
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
#if LSX_STAGE_MODE == 2
    static constexpr int kBarOff = kIdsOff + 2 * CHUNK * 4;  // two mbarriers
    static constexpr size_t kSmemBytes = kBarOff + 16;
    int issued, waited;
#else
    static constexpr size_t kSmemBytes = kIdsOff + 2 * CHUNK * 4;
#endif
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
#if LSX_STAGE_MODE == 1
        // consecutive lanes copy consecutive 16-B pieces: piece t = 32 j + lane of the round's buffer
        constexpr int NP = RS / 4;  // 16-B pieces per record
#pragma unroll
        for (int j = 0; j < kPerLane; ++j) {
            const int t = j * 32 + (int)lane;
            const int slot = t / NP, piece = t - slot * NP;
            const uint32_t id_s = __shfl_sync(kFullMask, id, slot);
            if (slot < m) {
                LSX_CHECK_INDEX(id_s, chk_points, "Gaussian id of a list entry");
                cp_async16(sbase + buf * (uint32_t)kBufBytes + (uint32_t)t * 16u,
                           reinterpret_cast<const char*>(records + (size_t)id_s * RS) + piece * 16);
            }
        }
        cp_async_commit();
#elif LSX_STAGE_MODE == 2
        // one bulk copy per record, issued by the lane that holds its id; completion on the buffer's mbarrier
        if (m > 0) {
            const uint32_t bar = sbase + kBarOff + buf * 8u;
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(m * kRecBytes)) : "memory");
            if ((int)lane < m) {
                LSX_CHECK_INDEX(id, chk_points, "Gaussian id of a list entry");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 sbase + buf * (uint32_t)kBufBytes + lane * (uint32_t)kRecBytes),
                             "l"(records + (size_t)id * RS), "n"(kRecBytes), "r"(bar)
                             : "memory");
            }
            issued = q;
        }
#else
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
#endif
    }
#if LSX_STAGE_MODE == 2
    __device__ __forceinline__ void wait_round(int q) {
        const uint32_t bar = sbase + kBarOff + (uint32_t)(q & 1) * 8u;
        const uint32_t parity = (uint32_t)(q >> 1) & 1u;
        uint32_t done;
        do {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(parity)
                : "memory");
        } while (!done);
        waited = q;
    }
#endif
    // all 32 threads; the record buffers must be free (not aliased by live data) from here on
    __device__ __forceinline__ void start(unsigned char* smem, const uint32_t* list_, const uint32_t* plist_,
                                          const float* records_, int first_, int dir_, int count_, int range_len = 0x7fffffff,
                                          int points = 0x7fffffff) {
        chk_range = range_len;
        chk_points = points;
        sbase = smem_u32(smem);
#if LSX_STAGE_MODE == 2
        issued = waited = -1;
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbase + kBarOff) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbase + kBarOff + 8u) : "memory");
        }
        // the barriers' initialisation and earlier generic writes to the buffers (the backward's transposition scratch) must be
        // visible to the copy engine
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
#endif
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
#if LSX_STAGE_MODE == 2
        __syncwarp();  // every lane is done reading the buffer the copy engine is about to overwrite
#endif
        issue(r + 1, id_next);  // an empty group when round r + 1 does not exist
        id_next = load_id(pos_next, r + 2);
        pos_next = load_pos(r + 3);
#if LSX_STAGE_MODE == 2
        wait_round(r);
#else
        cp_async_wait<1>();
#endif
        __syncwarp();
    }
#if LSX_STAGE_MODE == 2
    __device__ __forceinline__ void drain() {
        if (issued > waited) wait_round(issued);
    }
#else
    __device__ __forceinline__ void drain() { cp_async_wait<0>(); }
#endif
    __device__ __forceinline__ uint32_t rec_addr(int buf) const { return sbase + (uint32_t)(buf * kBufBytes); }
    __device__ __forceinline__ uint32_t ids_addr(int buf) const { return sbase + (uint32_t)(kIdsOff + buf * CHUNK * 4); }
};

}  // namespace lsx

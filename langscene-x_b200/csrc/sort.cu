// sort.cu — device-wide exclusive scan and stable LSD radix sort of (u32 key, u32 value) pairs.
//
// These replace the reference's CUB calls on the hot path
//   cub::DeviceScan::InclusiveSum        rasterizer_impl.cu:287
//   cub::DeviceRadixSort::SortPairs      rasterizer_impl.cu:313-318 (64-bit tile|depth keys)
//   cub::DeviceRadixSort::SortPairs      simple_knn.cu:210-213     (30-bit Morton codes)
//
// The 64-bit (tile << 32 | depth) sort of the reference is an LSD radix sort: the low 32 bits (depth)
// are processed first, the tile bits last.  Because all duplicates of one Gaussian carry the same depth
// and are emitted contiguously in index order, the depth passes can run on the P Gaussians BEFORE
// duplication (4 passes over P pairs) and only the tile passes (ceil(bit/8), 2 at 1080p) have to touch
// the R duplicated pairs; the sorted output is identical (see binning.cu).  Both stages use this one
// 32-bit-key kernel, as does the Morton sort of the KNN initialisation.
//
// Pass structure (8-bit digits): per-block digit histogram (+ digit and block-group totals by atomics) -> stable
// scatter that derives its own offsets.  In the scatter kernel each warp owns a contiguous 512-key segment and ranks its keys with
// per-bit warp ballots against warp-private shared-memory counters; a 256-thread step turns the per-warp
// counters into block-local and global positions; the block's 4096 pairs are regrouped by digit in shared
// memory and leave as contiguous runs (coalesced stores instead of 4-byte scatters).  No inter-block spinning anywhere (see B200_PROFILING.md on why).
#include "kernels.cuh"

namespace lsx {

namespace {

constexpr int kThreads = 256;
constexpr int kItems = 16;
constexpr int kTile = kThreads * kItems;  // 4096 elements per block
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// scan
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t& block_total) {
    // returns the exclusive prefix of v over the block's threads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t wsum = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0;
    uint32_t winc = wsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, winc, o);
        if (lane >= o) winc += t;
    }
    const uint32_t warp_off = __shfl_sync(kFull, winc - wsum, warp);
    block_total = __shfl_sync(kFull, winc, (blockDim.x >> 5) - 1);
    __syncthreads();
    return warp_off + inc - v;
}

__global__ void __launch_bounds__(kThreads) scan_reduce_kernel(const uint32_t* __restrict__ in,
                                                               const uint32_t* __restrict__ gather, int n,
                                                               uint32_t* __restrict__ partials) {
    __shared__ uint32_t s_warp[32];
    pdl_wait();
    const int base = blockIdx.x * kTile;
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int idx = base + i * kThreads + threadIdx.x;
        if (idx < n) sum += gather ? in[gather[idx]] : in[idx];
    }
    uint32_t total;
    block_exclusive_scan(sum, s_warp, total);
    if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_partials_kernel(uint32_t* __restrict__ partials, int nb,
                                                             uint32_t* __restrict__ total_out) {
    __shared__ uint32_t s_warp[32];
    pdl_wait();
    uint32_t carry = 0;
    for (int base = 0; base < nb; base += 1024) {
        const int idx = base + threadIdx.x;
        const uint32_t v = idx < nb ? partials[idx] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, s_warp, total);
        if (idx < nb) partials[idx] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(kThreads) scan_apply_kernel(const uint32_t* __restrict__ in,
                                                              const uint32_t* __restrict__ gather,
                                                              uint32_t* __restrict__ out, int n,
                                                              const uint32_t* __restrict__ partials) {
    __shared__ uint32_t s_warp[32];
    pdl_wait();
    const int base = blockIdx.x * kTile + threadIdx.x * kItems;  // blocked arrangement
    uint32_t v[kItems];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int idx = base + i;
        v[i] = idx < n ? (gather ? in[gather[idx]] : in[idx]) : 0;
        sum += v[i];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan(sum, s_warp, total) + partials[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int idx = base + i;
        if (idx < n) out[idx] = run;
        run += v[i];
    }
}

// ---------------------------------------------------------------------------------------------
// radix sort pass
// ---------------------------------------------------------------------------------------------
// A pass is TWO launches, whatever the size (sorts on this path are bound by launch latency as much as by bandwidth):
// the histogram kernel writes block-major per-block counts hist[b * 256 + d] and adds them, with global atomics, to
// the digit totals gsum[d] and to the totals of its group of kGroup consecutive blocks gsum[256 * (1 + b / kGroup) + d];
// the scatter kernel derives its own offsets: digit base (block scan of the totals) + earlier groups + earlier blocks
// of its own group — at most nb / kGroup + kGroup - 1 coalesced 1-KB rows instead of a device-wide scan (3 more
// launches) over the digit-major table.  No inter-block waiting anywhere.
constexpr int kGroup = 64;

// IT = keys per thread (a block sorts 256 IT keys): 16 for long arrays; 8 / 4 for short ones, whose few blocks are bound by the
// latency of one block's serial ranking work, not by throughput (radix_items below).
template <int IT>
__global__ void __launch_bounds__(kThreads) radix_hist_kernel(const uint32_t* __restrict__ keys, int n, int shift,
                                                              uint32_t mask, uint32_t* __restrict__ hist,
                                                              uint32_t* __restrict__ gsum) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    pdl_wait();
    __syncthreads();
    const int base = blockIdx.x * (kThreads * IT);
    // all keys of the thread are fetched before the first shared-memory atomic (interleaved, every atomic waited for
    // its own load: the kernel was a chain of 16 exposed global-load latencies)
    uint32_t k[IT];
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        const int idx = base + i * kThreads + threadIdx.x;
        k[i] = idx < n ? __ldg(keys + idx) : 0u;
    }
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        const int idx = base + i * kThreads + threadIdx.x;
        if (idx < n) atomicAdd(&h[(k[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    const uint32_t c = h[threadIdx.x];
    hist[blockIdx.x * 256 + threadIdx.x] = c;
    if (c) {
        atomicAdd(&gsum[threadIdx.x], c);
        atomicAdd(&gsum[256 * (1 + blockIdx.x / kGroup) + threadIdx.x], c);
    }
}

// 3 CTAs / SM (80 registers): measured 0.169 -> 0.151 ms for the two tile passes at C3 against 2 CTAs / SM; 4 spills
// MINB = 3: 80 registers; MINB = 4: 64 registers and 14 spilled words — pays only on long arrays (several waves of blocks):
// C5's tile sort (26 M pairs) 0.552 -> 0.480 ms, its depth sort (5 M) 0.223 -> 0.216; at 1 M keys it costs 3-5 %
template <int IT, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) radix_scatter_kernel(const uint32_t* __restrict__ keys_in,
                                                                 const uint32_t* __restrict__ vals_in,
                                                                 uint32_t* __restrict__ keys_out,
                                                                 uint32_t* __restrict__ vals_out, int n, int shift,
                                                                 uint32_t mask, const uint32_t* __restrict__ hist,
                                                                 const uint32_t* __restrict__ gsum) {
    __shared__ uint32_t cnt[kWarps][256];
    __shared__ uint32_t s_warp[32];
    constexpr int TILE = kThreads * IT;
    static_assert(TILE >= 1024, "the offset step reuses the first 1024 words of s_key");
    pdl_wait();
    __shared__ __align__(16) uint32_t s_key[TILE], s_val[TILE];   // the block's pairs regrouped by digit before they leave
    __shared__ uint32_t s_lstart[256], s_gbase[256];  // per digit: start inside the block / in the output array
    uint32_t my_start;  // global position of this block's first key with digit threadIdx.x
    {
        uint32_t total;
        const uint32_t digit_base = block_exclusive_scan(gsum[threadIdx.x], s_warp, total);
        // keys with each digit in earlier blocks = the earlier groups' totals + the earlier blocks of this group: a list
        // of 1-KB rows, summed by four 64-thread row partitions with 16-B loads (<= (nb / kGroup + kGroup) / 4 loads per
        // thread, 8 in flight), combined through s_key (not yet live)
        const int grp = blockIdx.x / kGroup, in_grp = blockIdx.x - grp * kGroup;
        const int part = threadIdx.x >> 6, col4 = threadIdx.x & 63;
        uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 8
        for (int r = part; r < grp + in_grp; r += 4) {
            const uint32_t* row = r < grp ? gsum + 256 * (1 + r) : hist + (size_t)(grp * kGroup + (r - grp)) * 256;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(row) + col4);
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
        }
        reinterpret_cast<uint4*>(s_key)[part * 64 + col4] = acc;
        __syncthreads();
        my_start = digit_base + s_key[threadIdx.x] + s_key[256 + threadIdx.x] + s_key[512 + threadIdx.x] + s_key[768 + threadIdx.x];
    }
    for (int i = threadIdx.x; i < kWarps * 256; i += kThreads) (&cnt[0][0])[i] = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wbase = blockIdx.x * TILE + warp * (32 * IT);
    const uint32_t lt_mask = (1u << lane) - 1u;

    // keys AND values are fetched up front, 32 independent loads per thread in flight (fetching a value only when it
    // is regrouped exposes one global-load latency per item: 40 % of this kernel's stall samples before the change)
    uint32_t key[IT], val[IT];
    uint32_t rank[IT / 2];  // two 16-bit ranks per register (a warp segment holds 512 keys)
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        const int idx = wbase + i * 32 + lane;
        key[i] = idx < n ? keys_in[idx] : 0u;
    }
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        const int idx = wbase + i * 32 + lane;
        val[i] = (vals_in && idx < n) ? vals_in[idx] : (uint32_t)idx;
    }
    // ranking in groups of 8 items: the warp matches of a group are independent and issue back to back; only the
    // counter updates (leader lane per distinct digit) form a serial chain
    constexpr int G = IT < 8 ? IT : 8;
#pragma unroll
    for (int g = 0; g < IT; g += G) {
        uint32_t peers[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int idx = wbase + (g + j) * 32 + lane;
            // lanes with the same digit, from 8 ballots (one per digit bit) instead of MATCH.ANY, whose result the warp waits
            // for (half of this kernel's stall samples with it; C3 tile sort 0.151 -> 0.128 ms, C5 0.69 -> 0.59); digit bits
            // above the pass width are zero in every lane
            const uint32_t d = (key[g + j] >> shift) & mask;
            uint32_t pm = __ballot_sync(kFull, idx < n);
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const bool bit = (d >> b) & 1u;
                const uint32_t v = __ballot_sync(kFull, bit);
                pm &= bit ? v : ~v;
            }
            peers[j] = pm;
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int idx = wbase + (g + j) * 32 + lane;
            const bool valid = idx < n;
            const uint32_t d = (key[g + j] >> shift) & mask;
            const int leader = __ffs(peers[j]) - 1;
            uint32_t old = 0;
            if (lane == leader && valid) {
                old = cnt[warp][d];
                cnt[warp][d] = old + __popc(peers[j]);
            }
            old = __shfl_sync(kFull, old, leader);
            const uint32_t r = old + __popc(peers[j] & lt_mask);
            if (j & 1) rank[(g + j) >> 1] |= r << 16;
            else rank[(g + j) >> 1] = r;
            __syncwarp();
        }
    }
    __syncthreads();
    // per digit: block total -> block-local start (exclusive scan over digits), per-warp local starts, global start
    {
        const uint32_t d = threadIdx.x;
        uint32_t total_d = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) total_d += cnt[w][d];
        uint32_t block_total;
        const uint32_t lstart = block_exclusive_scan(total_d, s_warp, block_total);
        uint32_t run = lstart;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t t = cnt[w][d];
            cnt[w][d] = run;
            run += t;
        }
        s_lstart[d] = lstart;
        s_gbase[d] = d <= mask ? my_start : 0u;
    }
    __syncthreads();
    // local shuffle: the block's keys (and values) land in shared memory grouped by digit, in stable order
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        const int idx = wbase + i * 32 + lane;
        if (idx < n) {
            const uint32_t d = (key[i] >> shift) & mask;
            const uint32_t lp = cnt[warp][d] + ((rank[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
            s_key[lp] = key[i];
            s_val[lp] = val[i];
        }
    }
    __syncthreads();
    // write out: consecutive threads write consecutive positions of a digit's run (coalesced within runs)
    const int count = min(TILE, n - (int)blockIdx.x * TILE);
    for (int j = threadIdx.x; j < count; j += kThreads) {
        const uint32_t k = s_key[j];
        const uint32_t d = (k >> shift) & mask;
        const uint32_t pos = s_gbase[d] + ((uint32_t)j - s_lstart[d]);
        LSX_CHECK_INDEX(pos, n, "radix scatter destination");
        keys_out[pos] = k;
        vals_out[pos] = s_val[j];
    }
}

}  // namespace

size_t scan_temp_bytes(int n) { return align_up((size_t)ceil_div(n > 0 ? n : 1, kTile) * sizeof(uint32_t), 256); }

int exclusive_scan_u32(const uint32_t* in, const uint32_t* gather, uint32_t* out, int n, uint32_t* total, void* temp,
                       cudaStream_t stream, bool debug) {
    if (n <= 0) {
        if (total) LSX_CUDA_OK(cudaMemsetAsync(total, 0, sizeof(uint32_t), stream));
        return 0;
    }
    const int nb = ceil_div(n, kTile);
    uint32_t* partials = static_cast<uint32_t*>(temp);
    launch_pdl(scan_reduce_kernel, nb, kThreads, 0, stream, in, gather, n, partials);
    LSX_KERNEL_OK(stream, debug);
    launch_pdl(scan_partials_kernel, 1, 1024, 0, stream, partials, nb, total);
    LSX_KERNEL_OK(stream, debug);
    launch_pdl(scan_apply_kernel, nb, kThreads, 0, stream, in, gather, out, n, partials);
    LSX_KERNEL_OK(stream, debug);
    return 0;
}

constexpr int kMaxPasses = 4;  // 32-bit keys, 8-bit digits

static size_t gsum_words(int nb) { return (size_t)256 * (1 + ceil_div(nb, kGroup)); }

// keys per thread for an array of n keys: a short array gets smaller blocks, so that its sort is not one partial wave of blocks
// that each rank 16 keys per thread one after the other (444 blocks are resident at 3 per SM).  Depth sort: C1 (10 k keys)
// 0.063 -> 0.042 ms, C2 (100 k) 0.070 -> 0.055, C4 (500 k) 0.072 -> 0.063 (profiles/r7a_stage_times.jsonl); what is left is
// launch latency: 8 kernels of ~5-8 us (4 keys per thread at 500 k changes nothing)
static int radix_items(int n) { return n <= 300 * 1024 ? 4 : (n <= 880 * 1024 ? 8 : 16); }

// sized for the smallest block whatever n is, so that the size stays monotonic in n (lsx_binning_capacity inverts it)
size_t radix_sort_temp_bytes(int n) {
    const int nb = ceil_div(n > 0 ? n : 1, kThreads * 4);
    return align_up((size_t)256 * nb * sizeof(uint32_t), 256) + kMaxPasses * gsum_words(nb) * sizeof(uint32_t);
}

int radix_sort_pairs_u32(uint32_t* keys[2], uint32_t* vals[2], int n, int begin_bit, int end_bit, bool identity_vals,
                         void* temp, int* result_buf, cudaStream_t stream, bool debug) {
    *result_buf = 0;
    if (n <= 0) return 0;
    if (radix_sort_num_passes(begin_bit, end_bit) > kMaxPasses) {
        set_error("radix_sort_pairs_u32: more than %d passes requested (bits %d..%d)", kMaxPasses, begin_bit, end_bit);
        return -1;
    }
    const int items = radix_items(n);
    const int nb = ceil_div(n, kThreads * items);
    uint32_t* hist = static_cast<uint32_t*>(temp);
    uint32_t* gsum_all = reinterpret_cast<uint32_t*>(static_cast<char*>(temp) + align_up((size_t)256 * nb * sizeof(uint32_t), 256));
    const size_t gw = gsum_words(nb);
    LSX_CUDA_OK(cudaMemsetAsync(gsum_all, 0, kMaxPasses * gw * sizeof(uint32_t), stream));
    int cur = 0;
    bool first = true;
    int pass = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8, ++pass) {
        const int bits = (end_bit - shift) < 8 ? (end_bit - shift) : 8;
        const uint32_t mask = (1u << bits) - 1u;
        uint32_t* gsum = gsum_all + pass * gw;
        const uint32_t* vin = (first && identity_vals) ? nullptr : vals[cur];
        if (items == 4) {
            launch_pdl(radix_hist_kernel<4>, nb, kThreads, 0, stream, keys[cur], n, shift, mask, hist, gsum);
            LSX_KERNEL_OK(stream, debug);
            launch_pdl(radix_scatter_kernel<4, 3>, nb, kThreads, 0, stream, keys[cur], vin, keys[cur ^ 1], vals[cur ^ 1], n, shift, mask,
                                                                 hist, gsum);
        } else if (items == 8) {
            launch_pdl(radix_hist_kernel<8>, nb, kThreads, 0, stream, keys[cur], n, shift, mask, hist, gsum);
            LSX_KERNEL_OK(stream, debug);
            launch_pdl(radix_scatter_kernel<8, 3>, nb, kThreads, 0, stream, keys[cur], vin, keys[cur ^ 1], vals[cur ^ 1], n, shift, mask,
                                                                 hist, gsum);
        } else {
            launch_pdl(radix_hist_kernel<16>, nb, kThreads, 0, stream, keys[cur], n, shift, mask, hist, gsum);
            LSX_KERNEL_OK(stream, debug);
            if (n >= 3 * 1024 * 1024)
                launch_pdl(radix_scatter_kernel<16, 4>, nb, kThreads, 0, stream, keys[cur], vin, keys[cur ^ 1], vals[cur ^ 1], n, shift,
                                                                         mask, hist, gsum);
            else
                launch_pdl(radix_scatter_kernel<16, 3>, nb, kThreads, 0, stream, keys[cur], vin, keys[cur ^ 1], vals[cur ^ 1], n, shift,
                                                                         mask, hist, gsum);
        }
        LSX_KERNEL_OK(stream, debug);
        cur ^= 1;
        first = false;
    }
    *result_buf = cur;
    return 0;
}

}  // namespace lsx

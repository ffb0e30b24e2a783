// api.cu — the C ABI of liblsx_b200.so (see include/lsx_rasterizer.h) and the host-side orchestration
// that replaces CudaRasterizer::Rasterizer::{forward,backward,markVisible}
// (diff-langsurf-rasterizer/cuda_rasterizer/rasterizer_impl.cu:141-153,198-362,366-491) and
// SimpleKNN::knn (simple-knn/simple_knn.cu:185-221).
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/lsx_rasterizer.h"
#include "kernels.cuh"

namespace lsx {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ---- optional per-stage device timing (bench.py's live roofline numbers) ------------------------------
// When enabled, every stage of forward/backward is bracketed by CUDA events recorded on the caller's
// stream; lsx_profile_read() synchronises them and returns the accumulated milliseconds per stage.
// Disabled (the default) this costs one relaxed atomic load per stage.
static std::atomic<int> g_profile{0};
static std::mutex g_profile_mu;
struct StageSpan {
    int stage;
    cudaEvent_t start, stop;
};
static std::vector<StageSpan> g_spans;
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t take_event() {
    if (!g_event_pool.empty()) {
        cudaEvent_t e = g_event_pool.back();
        g_event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct StageTimer {
    int stage;
    cudaStream_t stream;
    cudaEvent_t start = nullptr;
    StageTimer(int st, cudaStream_t s) : stage(st), stream(s) {
        if (g_profile.load(std::memory_order_relaxed)) {
            std::lock_guard<std::mutex> lk(g_profile_mu);
            start = take_event();
            cudaEventRecord(start, stream);
        }
    }
    ~StageTimer() {
        if (start) {
            std::lock_guard<std::mutex> lk(g_profile_mu);
            cudaEvent_t stop = take_event();
            cudaEventRecord(stop, stream);
            g_spans.push_back({stage, start, stop});
        }
    }
};

namespace {

constexpr size_t kAlign = 256;

struct Carver {  // carves aligned sub-arrays out of one allocation; with base == nullptr it just measures
    char* base;
    size_t off = 0;
    explicit Carver(char* b) : base(b) {}
    template <typename T>
    T* take(size_t count, size_t* offset_out = nullptr) {
        off = align_up(off, kAlign);
        if (offset_out) *offset_out = off;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
    size_t total() const { return align_up(off, kAlign); }
};

struct GeomScratch {
    float* depths;
    uint8_t* clamped;
    float2* means2D;
    float* cov3D;
    float4* conic_opacity;
    float* rgb;
    uint32_t* tiles_touched;
    float* records;
    float* grad_records;  // packed per-Gaussian gradient accumulators for the backward pass (zeroed there)
    // forward-only temporaries
    uint32_t* depth_keys[2];
    uint32_t* order[2];  // Gaussian indices, ping-pong; the depth-sorted order ends in one of them
    uint32_t* offsets;
    uint32_t* total;
    void* sort_temp;
    void* scan_temp;
    float* pose_partials;  // raw mode backward: one row of 16 pose-gradient partial sums per preprocess block
    size_t bytes;
    lsx_scratch_layout lay;
};

GeomScratch carve_geom(char* base, int P, int rec_stride) {
    GeomScratch g{};
    Carver c(base);
    const size_t p = (size_t)P;
    g.depths = c.take<float>(p, &g.lay.depths);
    g.clamped = c.take<uint8_t>(p, &g.lay.clamped);
    g.means2D = c.take<float2>(p, &g.lay.means2D);
    g.cov3D = c.take<float>(6 * p, &g.lay.cov3D);
    g.conic_opacity = c.take<float4>(p, &g.lay.conic_opacity);
    g.rgb = c.take<float>(3 * p, &g.lay.rgb);
    g.tiles_touched = c.take<uint32_t>(p, &g.lay.tiles_touched);
    g.records = c.take<float>(p * rec_stride, &g.lay.records);
    g.lay.record_stride = rec_stride;
    g.grad_records = c.take<float>(p * (size_t)(rec_stride - REC_HEAD + 8));  // >= round_up4(channels) + 8 per Gaussian
    g.depth_keys[0] = c.take<uint32_t>(p);
    g.depth_keys[1] = c.take<uint32_t>(p);
    g.order[0] = c.take<uint32_t>(p);
    g.order[1] = c.take<uint32_t>(p);
    g.offsets = c.take<uint32_t>(p);
    g.total = c.take<uint32_t>(64);
    g.sort_temp = c.take<char>(radix_sort_temp_bytes(P));
    g.scan_temp = c.take<char>(scan_temp_bytes(P));
    g.pose_partials = c.take<float>((size_t)ceil_div(P > 0 ? P : 1, LSX_PRE_BWD_THREADS) * 16);
    g.bytes = c.total();
    g.lay.geom_bytes = g.bytes;
    return g;
}

struct ImageScratch {
    float* final_T;
    uint32_t* n_contrib;
    uint2* ranges;
    uint32_t* k_contrib;  // per pixel: length of its block's compacted list up to its last contributor (render_fwd -> render_bwd)
    size_t bytes;
};

ImageScratch carve_image(char* base, int W, int H, lsx_scratch_layout* lay) {
    ImageScratch s{};
    Carver c(base);
    const size_t n = (size_t)W * H;
    const size_t tiles = (size_t)ceil_div(W, TILE_X) * ceil_div(H, TILE_Y);
    size_t o0, o1, o2;
    s.final_T = c.take<float>(n, &o0);
    s.n_contrib = c.take<uint32_t>(n, &o1);
    s.ranges = c.take<uint2>(tiles, &o2);
    size_t ok = 0;
    s.k_contrib = c.take<uint32_t>(n, &ok);
    s.bytes = c.total();
    if (lay) {
        lay->final_T = o0;
        lay->n_contrib = o1;
        lay->ranges = o2;
        lay->k_contrib = ok;
        lay->image_bytes = s.bytes;
    }
    return s;
}

struct BinningScratch {
    uint32_t* point_list;
    uint8_t* masks;  // sub-tile footprint mask per list entry (cull.cu), kept for the backward pass
    uint32_t* vals_alt;
    uint32_t* tile_keys[2];
    void* sort_temp;
    uint32_t* blk_list;  // 8 x R: per 8x4 block of every tile, the compacted list of its entries (cull.cu); stride R
    uint32_t* blk_cnt;   // 8 per tile
    size_t bytes;
};

BinningScratch carve_binning(char* base, int R, int num_tiles, lsx_scratch_layout* lay) {
    BinningScratch s{};
    Carver c(base);
    const size_t r = (size_t)(R > 0 ? R : 0);
    size_t o0;
    s.point_list = c.take<uint32_t>(r, &o0);
    size_t om = 0;
    s.masks = c.take<uint8_t>(r, &om);
    s.vals_alt = c.take<uint32_t>(r);
    s.tile_keys[0] = c.take<uint32_t>(r);
    s.tile_keys[1] = c.take<uint32_t>(r);
    s.sort_temp = c.take<char>(radix_sort_temp_bytes(R));
    size_t obl = 0, obc = 0;
    s.blk_list = c.take<uint32_t>(8 * r, &obl);
    s.blk_cnt = c.take<uint32_t>(8 * (size_t)(num_tiles > 0 ? num_tiles : 0), &obc);
    s.bytes = c.total() + kAlign;
    if (lay) {
        lay->point_list = o0;
        lay->binning_bytes = s.bytes;
        lay->masks = om;
        lay->blk_list = obl;
        lay->blk_cnt = obc;
    }
    return s;
}

// number of bits needed to hold `n` (the reference's getHigherMsb, rasterizer_impl.cu:35-50)
int bits_for(uint32_t n) {
    int b = 0;
    while (n >> b) ++b;
    return b;
}

// ---- binning capacity ------------------------------------------------------------------------------------------
// The binning scratch is carved for a CAPACITY of list slots, a multiple of 64: num_rendered rounded up, or the
// caller's speculative hint.  The backward pass (and the debug / statistics helpers) recover the capacity from the
// size of the binning buffer, which is strictly increasing over multiples of 64.
int round_up64(long long r) { return (int)((r + 63) / 64 * 64); }

int capacity_from_bytes(size_t bytes, int num_tiles) {
    long long lo = 0, hi = (0x7fffffffll / 64);  // capacity = 64 * x
    while (lo < hi) {
        const long long mid = (lo + hi + 1) / 2;
        if (carve_binning(nullptr, (int)(mid * 64), num_tiles, nullptr).bytes <= bytes) lo = mid; else hi = mid - 1;
    }
    const int cap = (int)(lo * 64);
    return carve_binning(nullptr, cap, num_tiles, nullptr).bytes == bytes ? cap : -1;
}

// pinned word + event for the asynchronous read of num_rendered in speculative mode (host-side pool, no device state)
struct HostSlot {
    uint32_t* word = nullptr;
    cudaEvent_t ev = nullptr;
};
std::mutex g_slot_mu;
std::vector<HostSlot> g_slots;

bool acquire_slot(HostSlot* out) {
    {
        std::lock_guard<std::mutex> lk(g_slot_mu);
        if (!g_slots.empty()) {
            *out = g_slots.back();
            g_slots.pop_back();
            return true;
        }
    }
    HostSlot s;
    if (cudaHostAlloc(reinterpret_cast<void**>(&s.word), 64, cudaHostAllocDefault) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming) != cudaSuccess) {
        cudaFreeHost(s.word);
        return false;
    }
    *out = s;
    return true;
}
void release_slot(const HostSlot& s) {
    std::lock_guard<std::mutex> lk(g_slot_mu);
    g_slots.push_back(s);
}

int blend_channels(int include_feature, int render_geo, int F, int Fi) {
    return 3 + (include_feature ? F + Fi : 0) + (render_geo ? 5 : 0);
}

}  // namespace
}  // namespace lsx

using namespace lsx;

extern "C" {

int lsx_abi_version(void) { return LSX_ABI_VERSION; }
const char* lsx_last_error(void) { return lsx::g_error; }
uint64_t lsx_kernel_launch_count(void) { return lsx::g_launches.load(std::memory_order_relaxed); }

void lsx_profile_enable(int enable) { lsx::g_profile.store(enable ? 1 : 0, std::memory_order_relaxed); }

int lsx_profile_read(float* ms_out, int n) {
    std::lock_guard<std::mutex> lk(lsx::g_profile_mu);
    for (int i = 0; i < n; ++i) ms_out[i] = 0.f;
    for (auto& sp : lsx::g_spans) {
        float ms = 0.f;
        if (cudaEventSynchronize(sp.stop) == cudaSuccess && cudaEventElapsedTime(&ms, sp.start, sp.stop) == cudaSuccess &&
            sp.stage >= 0 && sp.stage < n)
            ms_out[sp.stage] += ms;
        lsx::g_event_pool.push_back(sp.start);
        lsx::g_event_pool.push_back(sp.stop);
    }
    const int count = (int)lsx::g_spans.size();
    lsx::g_spans.clear();
    return count;
}

int lsx_scratch_layout_query(int32_t P, int32_t W, int32_t H, int32_t R, int32_t n_blend_channels,
                             lsx_scratch_layout* out) {
    if (!out || P < 0 || W <= 0 || H <= 0 || R < 0 || n_blend_channels < 3 || n_blend_channels > LSX_MAX_BLEND_CHANNELS) {
        set_error("lsx_scratch_layout_query: bad arguments");
        return -1;
    }
    GeomScratch g = carve_geom(nullptr, P, record_stride(n_blend_channels));
    *out = g.lay;
    carve_image(nullptr, W, H, out);
    carve_binning(nullptr, R, ceil_div(W, TILE_X) * ceil_div(H, TILE_Y), out);
    return 0;
}

int lsx_rasterize_forward(const lsx_forward_args* a, int32_t* num_rendered) {
    if (!a || !num_rendered) {
        set_error("lsx_rasterize_forward: null argument block");
        return -1;
    }
    *num_rendered = 0;
    cudaStream_t stream = static_cast<cudaStream_t>(a->stream);
    const bool debug = a->debug != 0;
    const int P = a->P, W = a->W, H = a->H;
    if (P < 0 || W <= 0 || H <= 0) {
        set_error("lsx_rasterize_forward: bad sizes P=%d W=%d H=%d", P, W, H);
        return -1;
    }
    const int F = a->include_feature ? a->F : 0, Fi = a->include_feature ? a->Fi : 0;
    const int nch = blend_channels(a->include_feature, a->render_geo, F, Fi);
    if (F < 0 || Fi < 0 || nch > LSX_MAX_BLEND_CHANNELS) {
        set_error("lsx_rasterize_forward: %d blended channels exceed the supported maximum of %d", nch,
                  LSX_MAX_BLEND_CHANNELS);
        return -1;
    }
    if (!a->out_color || !a->out_all_map || !a->out_plane_depth ||
        (P > 0 && (!a->radii || !a->out_observe || !a->background || !a->viewmatrix || !a->projmatrix || !a->campos))) {
        set_error("lsx_rasterize_forward: a required pointer is null");
        return -1;
    }
    const size_t HW = (size_t)W * H;
    if (P == 0) {
        // the reference skips the rasterizer entirely and returns its zero-initialised outputs
        LSX_CUDA_OK(cudaMemsetAsync(a->out_color, 0, 3 * HW * sizeof(float), stream));
        if (a->include_feature) {
            LSX_CUDA_OK(cudaMemsetAsync(a->out_language_feature, 0, (size_t)F * HW * sizeof(float), stream));
            LSX_CUDA_OK(cudaMemsetAsync(a->out_language_feature_instance, 0, (size_t)Fi * HW * sizeof(float), stream));
        }
        LSX_CUDA_OK(cudaMemsetAsync(a->out_all_map, 0, 5 * HW * sizeof(float), stream));
        LSX_CUDA_OK(cudaMemsetAsync(a->out_plane_depth, 0, HW * sizeof(float), stream));
        return 0;
    }
    if (!a->means3D || !a->opacities || (!a->shs && !a->colors_precomp) ||
        (!a->cov3D_precomp && (!a->scales || !a->rotations)) ||
        (a->include_feature && (!a->language_feature || !a->language_feature_instance || !a->out_language_feature ||
                                !a->out_language_feature_instance)) ||
        (a->render_geo && !a->all_map && !a->raw_params)) {
        set_error("lsx_rasterize_forward: a required input pointer is null");
        return -1;
    }
    if (a->raw_params && (a->cov3D_precomp || a->all_map || !a->scales || !a->rotations)) {
        set_error("lsx_rasterize_forward: raw_params needs scales + rotations and takes neither cov3D_precomp nor all_map");
        return -1;
    }
    if (reinterpret_cast<uintptr_t>(a->rotations) & 15u) {
        set_error("lsx_rasterize_forward: rotations must be 16-byte aligned (read as one 16-B word per Gaussian)");
        return -1;
    }
    if (!a->colors_precomp && a->M < (a->D + 1) * (a->D + 1)) {
        set_error("lsx_rasterize_forward: sh has %d coefficients but degree %d needs %d", a->M, a->D,
                  (a->D + 1) * (a->D + 1));
        return -1;
    }
    if (!a->geom_alloc || !a->binning_alloc || !a->image_alloc) {
        set_error("lsx_rasterize_forward: scratch allocators are required");
        return -1;
    }

    const float focal_y = H / (2.0f * a->tanfovy);
    const float focal_x = W / (2.0f * a->tanfovx);
    const uint32_t grid_x = (uint32_t)ceil_div(W, TILE_X), grid_y = (uint32_t)ceil_div(H, TILE_Y);
    const int num_tiles = (int)(grid_x * grid_y);
    const int rs = record_stride(nch);

    GeomScratch gm = carve_geom(nullptr, P, rs);
    char* gbase = a->geom_alloc(a->geom_user, gm.bytes);
    ImageScratch im = carve_image(nullptr, W, H, nullptr);
    char* ibase = a->image_alloc(a->image_user, im.bytes);
    if (!gbase || !ibase) {
        set_error("lsx_rasterize_forward: scratch allocation failed");
        return -4;
    }
    gm = carve_geom(gbase, P, rs);
    im = carve_image(ibase, W, H, nullptr);

    // ---- K1: per-Gaussian preprocess ------------------------------------------------------------
    PreprocessFwdParams pp{};
    pp.P = P; pp.D = a->D; pp.M = a->M; pp.W = W; pp.H = H; pp.F = F; pp.Fi = Fi;
    pp.grid_x = grid_x; pp.grid_y = grid_y;
    pp.focal_x = focal_x; pp.focal_y = focal_y; pp.tan_fovx = a->tanfovx; pp.tan_fovy = a->tanfovy;
    pp.scale_modifier = a->scale_modifier;
    pp.prefiltered = a->prefiltered; pp.render_geo = a->render_geo; pp.include_feature = a->include_feature;
    pp.rec_stride = rs; pp.n_channels = nch;
    pp.raw_params = a->raw_params; pp.pose = a->raw_params ? a->pose : nullptr;
    pp.means3D = a->means3D; pp.scales = a->scales; pp.rotations = a->rotations; pp.opacities = a->opacities;
    pp.shs = a->shs; pp.cov3D_precomp = a->cov3D_precomp; pp.colors_precomp = a->colors_precomp;
    pp.language_feature = a->language_feature; pp.language_feature_instance = a->language_feature_instance;
    pp.all_map = a->all_map; pp.view = a->viewmatrix; pp.proj = a->projmatrix; pp.campos = a->campos;
    pp.radii = a->radii; pp.out_observe = a->out_observe; pp.depths = gm.depths; pp.depth_keys = gm.depth_keys[0];
    pp.clamped = gm.clamped; pp.means2D = gm.means2D; pp.cov3D = gm.cov3D; pp.conic_opacity = gm.conic_opacity;
    pp.rgb = gm.rgb; pp.tiles_touched = gm.tiles_touched; pp.records = gm.records;
    int rc = 0;
    {
        StageTimer _t(LSX_STAGE_PREPROCESS_FWD, stream);
        rc = launch_preprocess_fwd(pp, stream, debug);
    }
    if (rc) return rc;

    // ---- depth-major presort of the Gaussians (the low-32-bit passes of the 64-bit key sort) ----
    int order_buf = 0;
    {
        StageTimer _t(LSX_STAGE_DEPTH_SORT, stream);
        rc = radix_sort_pairs_u32(gm.depth_keys, gm.order, P, 0, 32, /*identity_vals=*/true, gm.sort_temp, &order_buf,
                                  stream, debug);
    }
    if (rc) return rc;
    const uint32_t* order = gm.order[order_buf];

    // ---- K2: duplicate offsets in depth order; the total (num_rendered) stays on the device ------
    {
        StageTimer _t(LSX_STAGE_OFFSETS_SCAN, stream);
        rc = exclusive_scan_u32(gm.tiles_touched, order, gm.offsets, P, gm.total, gm.scan_temp, stream, debug);
        if (rc) return rc;
    }

    // ---- K3..K6 for a given list capacity: emit (tile, idx) pairs in depth order, stable tile sort, tile ranges,
    //      footprint masks + per-block lists, tile render.  Every kernel is sized by `cap`, never by num_rendered. ----
    const int tile_bits = bits_for((uint32_t)num_tiles);
    const int passes = radix_sort_num_passes(0, tile_bits);
    auto bin_and_render = [&](const int cap) -> int {
        BinningScratch bn = carve_binning(nullptr, cap, num_tiles, nullptr);
        char* bbase = a->binning_alloc(a->binning_user, bn.bytes);
        if (!bbase) {
            set_error("lsx_rasterize_forward: binning scratch allocation failed");
            return -4;
        }
        bn = carve_binning(bbase, cap, num_tiles, nullptr);
        uint32_t* vals[2];
        vals[passes & 1] = bn.point_list;  // so that the sorted values end in point_list
        vals[(passes & 1) ^ 1] = bn.vals_alt;
        int r2 = 0;
        if (cap > 0) {
            {
                StageTimer _t(LSX_STAGE_EMIT, stream);
                r2 = launch_emit_tile_pairs(P, order, gm.offsets, gm.means2D, a->radii, grid_x, grid_y, bn.tile_keys[0], vals[0],
                                            (uint32_t)cap, gm.total, stream, debug);
            }
            if (r2) return r2;
            int res = 0;
            {
                StageTimer _t(LSX_STAGE_TILE_SORT, stream);
                r2 = radix_sort_pairs_u32(bn.tile_keys, vals, cap, 0, tile_bits, false, bn.sort_temp, &res, stream, debug);
            }
            if (r2) return r2;
            {
                StageTimer _t(LSX_STAGE_TILE_RANGES, stream);
                r2 = launch_tile_ranges(cap, bn.tile_keys[res], im.ranges, num_tiles, stream, debug);
            }
            if (r2) return r2;
            {
                StageTimer _t(LSX_STAGE_FOOTPRINT_MASKS, stream);
                r2 = launch_footprint_masks(num_tiles, im.ranges, bn.point_list, gm.records, rs, grid_x, bn.masks, bn.blk_list,
                                            (size_t)cap, bn.blk_cnt, P, stream, debug);
            }
            if (r2) return r2;
        } else {
            r2 = launch_tile_ranges(0, nullptr, im.ranges, num_tiles, stream, debug);
            if (r2) return r2;
        }
        RenderParams rp{};
        rp.W = W; rp.H = H; rp.grid_x = grid_x; rp.grid_y = grid_y; rp.focal_x = focal_x; rp.focal_y = focal_y;
        rp.P = P; rp.R = cap;
        rp.F = F; rp.Fi = Fi; rp.include_feature = a->include_feature; rp.render_geo = a->render_geo;
        rp.n_channels = nch; rp.rec_stride = rs;
        rp.ranges = im.ranges; rp.point_list = bn.point_list; rp.masks = bn.masks; rp.records = gm.records;
        rp.blk_list = bn.blk_list; rp.list_stride = (size_t)cap; rp.blk_cnt = bn.blk_cnt; rp.k_contrib = im.k_contrib;
        rp.bg = a->background;
        rp.final_T = im.final_T; rp.n_contrib = im.n_contrib;
        rp.out_color = a->out_color; rp.out_language_feature = a->out_language_feature;
        rp.out_language_feature_instance = a->out_language_feature_instance; rp.out_observe = a->out_observe;
        rp.out_all_map = a->out_all_map; rp.out_plane_depth = a->out_plane_depth;
        {
            StageTimer _t(LSX_STAGE_RENDER_FWD, stream);
            r2 = launch_render_fwd(rp, stream, debug);
        }
        if (r2) return r2;
        if (!a->render_geo) {
            LSX_CUDA_OK(cudaMemsetAsync(a->out_all_map, 0, 5 * HW * sizeof(float), stream));
            LSX_CUDA_OK(cudaMemsetAsync(a->out_plane_depth, 0, HW * sizeof(float), stream));
        }
        return 0;
    };

    uint32_t R_host = 0;
    if (a->binning_capacity_hint > 0 && !debug) {
        // Speculative: everything behind the scan is enqueued for the hinted capacity BEFORE the host knows
        // num_rendered; the host then waits only for the event behind the scan (the GPU is already busy with the
        // kernels behind it), so the stream never drains in the middle of a forward call.  If the hint was too
        // small the binning + render are repeated with the exact capacity: results are always those of the exact path.
        HostSlot slot;
        if (!acquire_slot(&slot)) {
            set_error("lsx_rasterize_forward: could not allocate the pinned result word");
            return -2;
        }
        LSX_CUDA_OK(cudaMemcpyAsync(slot.word, gm.total, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        LSX_CUDA_OK(cudaEventRecord(slot.ev, stream));
        const int cap = round_up64(a->binning_capacity_hint);
        rc = bin_and_render(cap);
        const cudaError_t ee = cudaEventSynchronize(slot.ev);
        R_host = *slot.word;
        release_slot(slot);
        if (rc) return rc;
        if (ee != cudaSuccess) {
            set_error("lsx_rasterize_forward: waiting for num_rendered failed: %s", cudaGetErrorString(ee));
            return -2;
        }
        if (R_host > 0x7fffffffu - 64u) {
            set_error("lsx_rasterize_forward: %u duplicated splats overflow the 31-bit list index", R_host);
            return -5;
        }
        if (R_host > (uint32_t)cap) {  // hint too small: redo with the exact size (out_observe was counted into)
            LSX_CUDA_OK(cudaMemsetAsync(a->out_observe, 0, (size_t)P * sizeof(int), stream));
            rc = bin_and_render(round_up64(R_host));
            if (rc) return rc;
        }
    } else {
        LSX_CUDA_OK(cudaMemcpyAsync(&R_host, gm.total, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        LSX_CUDA_OK(cudaStreamSynchronize(stream));
        if (R_host > 0x7fffffffu - 64u) {
            set_error("lsx_rasterize_forward: %u duplicated splats overflow the 31-bit list index", R_host);
            return -5;
        }
        rc = bin_and_render(round_up64(R_host));
        if (rc) return rc;
    }
    *num_rendered = (int)R_host;
    return 0;
}

int lsx_rasterize_backward(const lsx_backward_args* a) {
    if (!a) {
        set_error("lsx_rasterize_backward: null argument block");
        return -1;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(a->stream);
    const bool debug = a->debug != 0;
    const int P = a->P, W = a->W, H = a->H, R = a->R;
    if (P < 0 || W <= 0 || H <= 0 || R < 0) {
        set_error("lsx_rasterize_backward: bad sizes P=%d W=%d H=%d R=%d", P, W, H, R);
        return -1;
    }
    if (P == 0) return 0;
    const int F = a->include_feature ? a->F : 0, Fi = a->include_feature ? a->Fi : 0;
    const int nch = blend_channels(a->include_feature, a->render_geo, F, Fi);
    if (nch > LSX_MAX_BLEND_CHANNELS) {
        set_error("lsx_rasterize_backward: too many blended channels (%d)", nch);
        return -1;
    }
    if (!a->geom_buffer || !a->image_buffer || (R > 0 && !a->binning_buffer) || !a->radii || !a->means3D ||
        !a->dL_dout_color || !a->dL_dmeans2D || !a->dL_dmeans2D_abs || !a->dL_dconic || !a->dL_dopacity ||
        !a->dL_dcolors || !a->dL_dmeans3D || (!a->dL_dcov3D && !a->raw_params) || !a->dL_dscales || !a->dL_drotations ||
        (!a->dL_dall_map && !a->raw_params) ||
        (a->M > 0 && !a->dL_dsh) ||
        (a->include_feature && (!a->dL_dout_language_feature || !a->dL_dout_language_feature_instance ||
                                !a->dL_dlanguage_feature || !a->dL_dlanguage_feature_instance)) ||
        (a->render_geo && (!a->dL_dout_all_map || !a->dL_dout_plane_depth || !a->out_all_map))) {
        set_error("lsx_rasterize_backward: a required pointer is null");
        return -1;
    }
    const float focal_y = H / (2.0f * a->tanfovy);
    const float focal_x = W / (2.0f * a->tanfovx);
    const uint32_t grid_x = (uint32_t)ceil_div(W, TILE_X), grid_y = (uint32_t)ceil_div(H, TILE_Y);
    const int rs = record_stride(nch);
    GeomScratch gm = carve_geom(const_cast<char*>(a->geom_buffer), P, rs);
    ImageScratch im = carve_image(const_cast<char*>(a->image_buffer), W, H, nullptr);
    int cap = round_up64(R);
    if (a->binning_bytes != 0) {
        cap = capacity_from_bytes((size_t)a->binning_bytes, (int)(grid_x * grid_y));
        if (cap < R) {
            set_error("lsx_rasterize_backward: binning buffer of %llu bytes does not belong to a forward call of this size",
                      (unsigned long long)a->binning_bytes);
            return -1;
        }
    }
    BinningScratch bn = carve_binning(const_cast<char*>(a->binning_buffer), cap, (int)(grid_x * grid_y), nullptr);

    const int gs = round_up4(nch) + 8;  // floats per packed gradient record
    {
        StageTimer _t(LSX_STAGE_BWD_ZERO, stream);
        LSX_CUDA_OK(cudaMemsetAsync(gm.grad_records, 0, (size_t)P * gs * sizeof(float), stream));
    }

    int rc = 0;
    if (R > 0) {
        RenderParams rp{};
        rp.W = W; rp.H = H; rp.grid_x = grid_x; rp.grid_y = grid_y; rp.focal_x = focal_x; rp.focal_y = focal_y;
        rp.P = P; rp.R = cap;
        rp.F = F; rp.Fi = Fi; rp.include_feature = a->include_feature; rp.render_geo = a->render_geo;
        rp.n_channels = nch; rp.rec_stride = rs;
        rp.ranges = im.ranges; rp.point_list = bn.point_list; rp.masks = bn.masks; rp.records = gm.records;
        rp.blk_list = bn.blk_list; rp.list_stride = (size_t)cap; rp.blk_cnt = bn.blk_cnt; rp.k_contrib = im.k_contrib;
        rp.bg = a->background;
        rp.final_T = im.final_T; rp.n_contrib = im.n_contrib;
        rp.dL_dout_color = a->dL_dout_color; rp.dL_dout_language_feature = a->dL_dout_language_feature;
        rp.dL_dout_language_feature_instance = a->dL_dout_language_feature_instance;
        rp.dL_dout_all_map = a->dL_dout_all_map; rp.dL_dout_plane_depth = a->dL_dout_plane_depth;
        rp.all_map_pixels = a->out_all_map;
        rp.grad_records = gm.grad_records;
        {
            StageTimer _t(LSX_STAGE_RENDER_BWD, stream);
            rc = launch_render_bwd(rp, stream, debug);
        }
        if (rc) return rc;
    }

    PreprocessBwdParams bp{};
    bp.P = P; bp.D = a->D; bp.M = a->M;
    bp.focal_x = focal_x; bp.focal_y = focal_y; bp.tan_fovx = a->tanfovx; bp.tan_fovy = a->tanfovy;
    bp.scale_modifier = a->scale_modifier;
    bp.means3D = a->means3D; bp.radii = a->radii; bp.shs = a->shs; bp.clamped = gm.clamped;
    bp.scales = a->scales; bp.rotations = a->rotations; bp.cov3D = gm.cov3D; bp.cov3D_precomp = a->cov3D_precomp;
    bp.view = a->viewmatrix; bp.proj = a->projmatrix; bp.campos = a->campos;
    bp.grad_records = gm.grad_records; bp.grad_stride = gs; bp.n_channels_pad = round_up4(nch);
    bp.F = F; bp.Fi = Fi; bp.include_feature = a->include_feature; bp.render_geo = a->render_geo;
    bp.accumulate = a->accumulate_param_grads;
    bp.raw_params = a->raw_params; bp.pose = a->raw_params ? a->pose : nullptr; bp.conic_opacity = gm.conic_opacity;
    bp.pose_partials = gm.pose_partials; bp.dL_dpose = a->dL_dpose; bp.accumulate_pose = a->accumulate_pose;
    if (a->raw_params && (a->cov3D_precomp || !a->scales || !a->rotations)) {
        set_error("lsx_rasterize_backward: raw_params needs scales + rotations and no cov3D_precomp");
        return -1;
    }
    if ((reinterpret_cast<uintptr_t>(a->rotations) | reinterpret_cast<uintptr_t>(a->dL_drotations)) & 15u) {
        set_error("lsx_rasterize_backward: rotations and dL_drotations must be 16-byte aligned");
        return -1;
    }
    bp.geo_scale_x = 0.5f * (float)W; bp.geo_scale_y = 0.5f * (float)H;
    bp.dL_dmean2D = a->dL_dmeans2D; bp.dL_dmean2D_abs = a->dL_dmeans2D_abs; bp.dL_dconic = a->dL_dconic;
    bp.dL_dopacity = a->dL_dopacity; bp.dL_dcolor = a->dL_dcolors;
    bp.dL_dlanguage_feature = a->dL_dlanguage_feature;
    bp.dL_dlanguage_feature_instance = a->dL_dlanguage_feature_instance; bp.dL_dall_map = a->dL_dall_map;
    bp.dL_dmeans3D = a->dL_dmeans3D; bp.dL_dcov3D = a->dL_dcov3D; bp.dL_dsh = a->dL_dsh;
    bp.dL_dscales = a->dL_dscales; bp.dL_drotations = a->dL_drotations;
    StageTimer _t(LSX_STAGE_PREPROCESS_BWD, stream);
    return launch_preprocess_bwd(bp, stream, debug);
}

int lsx_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream) {
    (void)projmatrix;  // the reference computes the projected point but only tests view-space z
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) {
        set_error("lsx_mark_visible: bad arguments");
        return -1;
    }
    return launch_mark_visible(P, means3D, viewmatrix, present, static_cast<cudaStream_t>(stream));
}

int lsx_knn_mean_dist2(int32_t P, const float* points, float* out, lsx_alloc_fn alloc, void* alloc_user, void* stream) {
    if (P < 0 || (P > 0 && (!points || !out || !alloc))) {
        set_error("lsx_knn_mean_dist2: bad arguments");
        return -1;
    }
    if (P == 0) return 0;
    char* temp = alloc(alloc_user, knn_temp_bytes(P));
    if (!temp) {
        set_error("lsx_knn_mean_dist2: scratch allocation failed");
        return -4;
    }
    StageTimer _t(LSX_STAGE_KNN, static_cast<cudaStream_t>(stream));
    return knn_mean_dist2(P, points, out, temp, static_cast<cudaStream_t>(stream));
}

int32_t lsx_binning_capacity(size_t binning_bytes, int32_t W, int32_t H) {
    if (W <= 0 || H <= 0) return -1;
    return capacity_from_bytes(binning_bytes, ceil_div(W, TILE_X) * ceil_div(H, TILE_Y));
}

int lsx_debug_sorted_keys(int32_t P, int32_t W, int32_t H, int32_t R, int32_t n_blend_channels, const char* geom_buffer,
                          const char* binning_buffer, size_t binning_bytes, const char* image_buffer, uint64_t* keys_out,
                          void* stream) {
    if (P <= 0 || R <= 0 || !geom_buffer || !binning_buffer || !image_buffer || !keys_out) {
        set_error("lsx_debug_sorted_keys: bad arguments");
        return -1;
    }
    GeomScratch gm = carve_geom(const_cast<char*>(geom_buffer), P, record_stride(n_blend_channels));
    ImageScratch im = carve_image(const_cast<char*>(image_buffer), W, H, nullptr);
    const int tiles = ceil_div(W, TILE_X) * ceil_div(H, TILE_Y);
    const int cap = binning_bytes ? capacity_from_bytes(binning_bytes, tiles) : round_up64(R);
    if (cap < R) {
        set_error("lsx_debug_sorted_keys: binning buffer size does not match");
        return -1;
    }
    BinningScratch bn = carve_binning(const_cast<char*>(binning_buffer), cap, tiles, nullptr);
    return launch_debug_keys(tiles, im.ranges, bn.point_list, gm.depths, keys_out, static_cast<cudaStream_t>(stream));
}

int lsx_render_stats(int32_t P, int32_t W, int32_t H, int32_t R, int32_t n_blend_channels, const char* geom_buffer,
                     const char* binning_buffer, size_t binning_bytes, const char* image_buffer, uint64_t* stats_out,
                     void* stream) {
    if (P < 0 || W <= 0 || H <= 0 || R < 0 || !stats_out || n_blend_channels < 3 || n_blend_channels > LSX_MAX_BLEND_CHANNELS ||
        (P > 0 && (!geom_buffer || !image_buffer)) || (R > 0 && !binning_buffer)) {
        set_error("lsx_render_stats: bad arguments");
        return -1;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P == 0 || R == 0) {
        LSX_CUDA_OK(cudaMemsetAsync(stats_out, 0, 8 * sizeof(uint64_t), st));
        return 0;
    }
    const int rs = record_stride(n_blend_channels);
    GeomScratch gm = carve_geom(const_cast<char*>(geom_buffer), P, rs);
    ImageScratch im = carve_image(const_cast<char*>(image_buffer), W, H, nullptr);
    const uint32_t grid_x = (uint32_t)ceil_div(W, TILE_X), grid_y = (uint32_t)ceil_div(H, TILE_Y);
    const int cap = binning_bytes ? capacity_from_bytes(binning_bytes, (int)(grid_x * grid_y)) : round_up64(R);
    if (cap < R) {
        set_error("lsx_render_stats: binning buffer size does not match");
        return -1;
    }
    BinningScratch bn = carve_binning(const_cast<char*>(binning_buffer), cap, (int)(grid_x * grid_y), nullptr);
    RenderParams rp{};
    rp.W = W; rp.H = H; rp.grid_x = grid_x; rp.grid_y = grid_y; rp.n_channels = n_blend_channels; rp.rec_stride = rs;
    rp.ranges = im.ranges; rp.point_list = bn.point_list; rp.records = gm.records;
    rp.blk_list = bn.blk_list; rp.list_stride = (size_t)cap; rp.blk_cnt = bn.blk_cnt; rp.k_contrib = im.k_contrib;
    rp.n_contrib = im.n_contrib;
    return launch_render_stats(rp, reinterpret_cast<unsigned long long*>(stats_out), st);
}

}  // extern "C"

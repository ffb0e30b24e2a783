// kernels.cuh — parameter blocks and launcher prototypes shared by the .cu files of liblsx_b200.so
#pragma once
#include <cstdio>
#include "../../include/lsx_rasterizer.h"
#include "common.cuh"

namespace lsx {

// ---- per-Gaussian preprocess -----------------------------------------------------------------
struct PreprocessFwdParams {
    int P, D, M, W, H, F, Fi;
    uint32_t grid_x, grid_y;
    float focal_x, focal_y, tan_fovx, tan_fovy, scale_modifier;
    int prefiltered, render_geo, include_feature;
    int raw_params;     // fused render-wrapper mode: raw parameters in, activations / plane normal / all_map computed in-kernel
    const float* pose;  // raw mode: optional camera pose [quaternion | translation] (7 floats, device)
    int rec_stride;
    int n_channels;  // blended channels: 3 + F + Fi + (render_geo ? 5 : 0)
    const float* means3D;
    const float* scales;
    const float* rotations;
    const float* opacities;
    const float* shs;
    const float* cov3D_precomp;
    const float* colors_precomp;
    const float* language_feature;
    const float* language_feature_instance;
    const float* all_map;
    const float* view;
    const float* proj;
    const float* campos;
    // outputs
    int* radii;
    int* out_observe;
    float* depths;
    uint32_t* depth_keys;
    uint8_t* clamped;
    float2* means2D;
    float* cov3D;
    float4* conic_opacity;
    float* rgb;
    uint32_t* tiles_touched;
    float* records;
};

struct PreprocessBwdParams {
    int P, D, M;
    float focal_x, focal_y, tan_fovx, tan_fovy, scale_modifier;
    const float* means3D;
    const int* radii;
    const float* shs;
    const uint8_t* clamped;
    const float* scales;
    const float* rotations;
    const float* cov3D;          // scratch copy written by the forward pass
    const float* cov3D_precomp;  // or the caller's
    const float* view;
    const float* proj;
    const float* campos;
    // packed gradient records written by the tile backward pass (see RenderParams::grad_records)
    const float* grad_records;
    int grad_stride;  // floats per record
    int n_channels_pad;  // round_up4(blended channels) = offset of the geometry terms inside a record
    int F, Fi, include_feature, render_geo;
    // fused render-wrapper mode (see PreprocessFwdParams::raw_params): means3D / scales / rotations are the raw parameters, the
    // activated opacity comes from the forward pass's conic_opacity; pose_partials: blocks x 16 floats of scratch
    int raw_params;
    const float* pose;
    const float4* conic_opacity;
    float* pose_partials;
    float* dL_dpose;      // 7 floats
    int accumulate_pose;  // dL_dpose += instead of =
    int accumulate;  // LSX_ACC_* bit mask: parameter gradients with out += value (multi-view accumulation) instead of out = value
    // the tile pass leaves the constant factors of the geometry terms (0.5 W / 0.5 H of the mean2D
    // terms, -0.5 of the conic terms) to this kernel, which applies them once per Gaussian instead of once per visit
    float geo_scale_x, geo_scale_y;
    // per-tensor outputs in the reference's layouts (every row written; zeros for culled splats)
    float* dL_dmean2D;      // (P,3)
    float* dL_dmean2D_abs;  // (P,3)
    float* dL_dconic;       // (P,4) as {x,y,0,w}
    float* dL_dopacity;     // (P)
    float* dL_dcolor;       // (P,3)
    float* dL_dlanguage_feature;
    float* dL_dlanguage_feature_instance;
    float* dL_dall_map;     // (P,5)
    float* dL_dmeans3D;
    float* dL_dcov3D;
    float* dL_dsh;
    float* dL_dscales;
    float* dL_drotations;
};

int launch_preprocess_fwd(const PreprocessFwdParams& p, cudaStream_t stream, bool debug);
int launch_preprocess_bwd(const PreprocessBwdParams& p, cudaStream_t stream, bool debug);
// pose.cu: d_pose (7) from `rows` x 16 partial sums (fixed-order sum in double, chain rule through R(q / |q|))
int launch_pose_finish(int rows, const float* pose, const float* partials, float* d_pose, int accumulate, cudaStream_t stream);
int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present, cudaStream_t stream);

// ---- device-wide primitives (sort.cu) ---------------------------------------------------------
// scratch requirement (bytes) of radix_sort_pairs_u32 for n elements
size_t radix_sort_temp_bytes(int n);
// Stable LSD radix sort of (key,value) u32 pairs on key bits [begin_bit, end_bit).
// keys/vals are ping-pong buffers (index 0 holds the input; identity_vals means "input values = 0..n-1",
// vals[0] is then only used as a ping-pong target).  *result_buf = index of the buffer holding the output.
int radix_sort_pairs_u32(uint32_t* keys[2], uint32_t* vals[2], int n, int begin_bit, int end_bit, bool identity_vals,
                         void* temp, int* result_buf, cudaStream_t stream, bool debug);
// number of 8-bit passes radix_sort_pairs_u32 will run
static inline int radix_sort_num_passes(int begin_bit, int end_bit) { return (end_bit - begin_bit + 7) / 8; }

size_t scan_temp_bytes(int n);
// out[i] = sum_{j<i} in[gather ? gather[j] : j]; *total (device) = full sum
int exclusive_scan_u32(const uint32_t* in, const uint32_t* gather, uint32_t* out, int n, uint32_t* total, void* temp,
                       cudaStream_t stream, bool debug);

// ---- binning (binning.cu) ---------------------------------------------------------------------
int launch_emit_tile_pairs(int P, const uint32_t* sorted_idx, const uint32_t* offsets, const float2* means2D,
                           const int* radii, uint32_t grid_x, uint32_t grid_y, uint32_t* tile_keys, uint32_t* vals,
                           uint32_t capacity, const uint32_t* total_pairs, cudaStream_t stream, bool debug);
// R = slots of the key array (pairs + 0xffffffff padding)
int launch_tile_ranges(int R, const uint32_t* sorted_tile_keys, uint2* ranges, int num_tiles, cudaStream_t stream,
                       bool debug);
int launch_footprint_masks(int num_tiles, const uint2* ranges, const uint32_t* point_list, const float* records,
                           int rec_stride, uint32_t grid_x, uint8_t* masks, uint32_t* blk_list, size_t list_stride,
                           uint32_t* blk_cnt, int num_points, cudaStream_t stream, bool debug);
int launch_debug_keys(int num_tiles, const uint2* ranges, const uint32_t* point_list, const float* depths,
                      uint64_t* keys_out, cudaStream_t stream);

// ---- tile renderers (render_fwd.cu / render_bwd.cu) --------------------------------------------
struct RenderParams {
    int W, H;
    int P, R;  // Gaussians, list slots (capacity) — read by the LSX_BOUNDS_CHECK build only
    uint32_t grid_x, grid_y;
    float focal_x, focal_y;
    int F, Fi;  // feature widths (0 when include_feature is false)
    int include_feature, render_geo;
    int n_channels;  // 3 + F + Fi + (render_geo ? 5 : 0)
    int rec_stride;
    const uint2* ranges;
    const uint32_t* point_list;
    const uint8_t* masks;  // per list entry: bit w set iff the splat may blend inside block w of its tile (cull.cu)
    // per 8x4 block w of tile t: the positions (inside the tile's range) of the entries with bit w set, in list order, at
    // blk_list[w * list_stride + ranges[t].x + k], k < blk_cnt[8 t + w]  (cull.cu)
    const uint32_t* blk_list;
    size_t list_stride;
    const uint32_t* blk_cnt;
    // per pixel: number of elements of its block's compacted list up to and including its last contributor (forward ->
    // backward; n_contrib keeps the reference's meaning: position in the tile list + 1)
    uint32_t* k_contrib;
    const float* records;
    const float* bg;
    // forward outputs / backward inputs
    float* final_T;
    uint32_t* n_contrib;
    float* out_color;
    float* out_language_feature;
    float* out_language_feature_instance;
    int* out_observe;
    float* out_all_map;
    float* out_plane_depth;
    // backward only
    const float* dL_dout_color;
    const float* dL_dout_language_feature;
    const float* dL_dout_language_feature_instance;
    const float* dL_dout_all_map;
    const float* dL_dout_plane_depth;
    const float* all_map_pixels;
    // packed per-Gaussian gradient accumulators, stride round_up4(n_channels) + 8 floats, zero on entry:
    // [channels in record order | mean2D.x, mean2D.y, |mean2D|.x, |mean2D|.y, conic.x, conic.y, conic.w, opacity]
    float* grad_records;
};

int launch_render_fwd(const RenderParams& p, cudaStream_t stream, bool debug);
int launch_render_bwd(const RenderParams& p, cudaStream_t stream, bool debug);
// stats.cu: S, B, V, Vb, L of a rendered view (needs ranges, point_list, blk_list/blk_cnt, records, n_contrib, k_contrib)
int launch_render_stats(const RenderParams& p, unsigned long long* out8, cudaStream_t stream);

// ---- KNN (knn.cu) -------------------------------------------------------------------------------
size_t knn_temp_bytes(int P);
int knn_mean_dist2(int P, const float* points, float* out, void* temp, cudaStream_t stream);
// the search structure alone (Morton order, float4 stream, 3-level boxes) into `temp` (knn_temp_bytes(P) bytes), and exact
// K-nearest queries of S dataset points against it: cand_d / cand_i [S][K] sorted by (distance, index)
int knn_tree_build(int P, const float* points, void* temp, cudaStream_t stream, bool with_rank = true);
int knn_tree_query(int K, int P, const void* tree, int S, const float* points, const int* sample_idx, float* cand_d, int* cand_i,
                   cudaStream_t stream);

}  // namespace lsx

/*
 * lsx_rasterizer.h — C ABI of the B200-native LangSurf rasterizer + KNN initialisation library
 * (liblsx_b200.so).  Plain C: raw device pointers, sizes, an opaque stream handle, int status.
 * No torch / C++ types cross this boundary, no exception ever leaves the library.
 *
 * Each entry point replaces one function of the reference's native layer
 * (paths relative to field_construction/submodules/ in the reference tree):
 *
 *   lsx_rasterize_forward    <- CudaRasterizer::Rasterizer::forward   diff-langsurf-rasterizer/cuda_rasterizer/rasterizer.h:34-66
 *                               as called by RasterizeGaussiansCUDA     diff-langsurf-rasterizer/rasterize_points.cu:35-143
 *   lsx_rasterize_backward   <- CudaRasterizer::Rasterizer::backward  diff-langsurf-rasterizer/cuda_rasterizer/rasterizer.h:68-107
 *                               as called by RasterizeGaussiansBackwardCUDA  rasterize_points.cu:145-259
 *   lsx_mark_visible         <- CudaRasterizer::Rasterizer::markVisible  rasterizer.h:27-32 / rasterize_points.cu:261-280
 *   lsx_knn_mean_dist2       <- SimpleKNN::knn / distCUDA2             simple-knn/simple_knn.h:16, simple-knn/spatial.cu:15-25
 *   lsx_alloc_fn             <- std::function<char*(size_t)> from resizeFunctional  rasterize_points.cu:27-33
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless the name ends in _host;
 *     NULL means "absent" exactly like the reference's empty tensors (forward.cu:205,241).
 *   - all float tensors are contiguous fp32, integer tensors int32, row-major as in the reference
 *     (means3D (P,3), shs (P,M,3), opacities (P,1), scales (P,3), rotations (P,4) un-normalised,
 *     cov3D_precomp (P,6), all_map (P,5), viewmatrix/projmatrix 16 floats in the reference's
 *     transposed ("row-vector") memory order, images planar CHW).
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it.  The library holds no
 *     device-global state between calls: several forwards may be outstanding before their backwards.
 *   - return value 0 = success, negative = error; lsx_last_error() returns a thread-local message.
 *   - OUTPUT INITIALISATION: callers pass UNINITIALISED output / gradient buffers.  The library writes
 *     every element (zero where the reference's torch::full / torch::zeros would have left zero).
 */
#ifndef LSX_RASTERIZER_H_INCLUDED
#define LSX_RASTERIZER_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LSX_API __attribute__((visibility("default")))
#else
#define LSX_API
#endif

#define LSX_ABI_VERSION 5
#define LSX_MAX_BLEND_CHANNELS 40 /* 3 + F + Fi + 5 must not exceed this */

/* scratch allocation callback: must return a device pointer to at least `bytes` bytes, aligned to
 * >= 256 B, that stays valid until the matching backward call has been enqueued. */
typedef char* (*lsx_alloc_fn)(void* user, size_t bytes);

typedef struct lsx_forward_args {
    /* sizes / scalars (GaussianRasterizationSettings, diff_LangSurf_rasterization/__init__.py:189-203) */
    int32_t P;             /* number of Gaussians                               */
    int32_t D;             /* active SH degree (0..3)                           */
    int32_t M;             /* SH coefficients per colour stored in `shs` (0 if shs absent) */
    int32_t W, H;          /* image width / height                              */
    int32_t F;             /* language feature width (run-time; reference: compile-time 3)  */
    int32_t Fi;            /* instance feature width (reference: 3)            */
    float tanfovx, tanfovy;
    float scale_modifier;
    int32_t prefiltered;
    int32_t render_geo;
    int32_t debug;         /* !=0: synchronise + check after every stage (auxiliary.h:166-173) */
    int32_t include_feature;
    /* inputs */
    const float* background;                 /* 3  */
    const float* means3D;                    /* P*3 */
    const float* shs;                        /* P*M*3 or NULL */
    const float* colors_precomp;             /* P*3 or NULL */
    const float* language_feature;           /* P*F  (include_feature) */
    const float* language_feature_instance;  /* P*Fi (include_feature) */
    const float* opacities;                  /* P */
    const float* scales;                     /* P*3 or NULL */
    const float* rotations;                  /* P*4 or NULL; 16-byte aligned (every other array: 4-byte aligned is enough,
                                              * rows that happen to be 16-byte aligned are moved with 16-byte accesses) */
    const float* cov3D_precomp;              /* P*6 or NULL */
    const float* all_map;                    /* P*5 (render_geo) */
    const float* viewmatrix;                 /* 16 */
    const float* projmatrix;                 /* 16 */
    const float* campos;                     /* 3  */
    /* outputs (uninitialised on entry, fully written on return) */
    float* out_color;                        /* 3*H*W  */
    float* out_language_feature;             /* F*H*W  if include_feature else untouched */
    float* out_language_feature_instance;    /* Fi*H*W if include_feature else untouched */
    int32_t* radii;                          /* P */
    int32_t* out_observe;                    /* P */
    float* out_all_map;                      /* 5*H*W (zeros when !render_geo) */
    float* out_plane_depth;                  /* H*W   (zeros when !render_geo) */
    /* scratch (private layout; contents are consumed by lsx_rasterize_backward) */
    lsx_alloc_fn geom_alloc;    void* geom_user;
    lsx_alloc_fn binning_alloc; void* binning_user;
    lsx_alloc_fn image_alloc;   void* image_user;
    void* stream;
    /* ABI v5 — speculative binning capacity (0 = off).  The list length num_rendered is only known on the device after the
     * duplicate-offset scan; the reference reads it back in the middle of the forward pass and lets the stream drain
     * (rasterizer_impl.cu:291).  With a hint > 0 (e.g. 1.25 x the previous call's num_rendered for this scene) the library
     * sizes the binning scratch and every kernel behind the scan for that capacity, enqueues them at once and only then
     * waits for the (long finished) scan: the stream never drains.  If the hint turns out too small, binning and render are
     * repeated with the exact size before the call returns — outputs and *num_rendered are always those of the exact path. */
    int32_t binning_capacity_hint;
    /* ABI v5 — fused render-wrapper mode ("activations, all_map construction ... pose transform" folded
     * into the per-Gaussian kernels).  With raw_params != 0 the arrays means3D / scales / rotations / opacities hold the
     * reference's RAW nn.Parameters (positions, log-scales, un-normalised quaternions, opacity logits;
     * field_construction/scene/gaussian_model.py:193-213), all_map and cov3D_precomp must be NULL, and the preprocess kernel
     * itself applies the optional camera pose (`pose`: 7 floats [quaternion | translation], NULL = none;
     * gaussian_renderer/__init__.py:79-87), exp / normalize / sigmoid, the plane normal and the all_map row
     * (gaussian_renderer/__init__.py:188-196) — what lsx_pose_transform_forward + lsx_gaussian_head_forward compute in two
     * extra passes over P.  Outputs are unchanged. */
    int32_t raw_params;
    const float* pose;
} lsx_forward_args;

/* Returns 0 and stores the number of (Gaussian,tile) duplicates in *num_rendered.
 * Performs exactly one host<-device read (num_rendered), like rasterizer_impl.cu:291. */
LSX_API int lsx_rasterize_forward(const lsx_forward_args* args, int32_t* num_rendered);

typedef struct lsx_backward_args {
    int32_t P, D, M, W, H, F, Fi;
    int32_t R;                 /* num_rendered returned by the forward call */
    float tanfovx, tanfovy, scale_modifier;
    int32_t render_geo, debug, include_feature;
    /* forward inputs again */
    const float* background;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* language_feature;
    const float* language_feature_instance;
    const float* all_map;
    const float* scales;
    const float* rotations;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    const int32_t* radii;
    /* saved forward results */
    const float* out_all_map;          /* "all_map_pixels", 5*H*W */
    const char* geom_buffer;
    const char* binning_buffer;
    const char* image_buffer;
    /* upstream gradients, planar like the forward outputs */
    const float* dL_dout_color;                      /* 3*H*W */
    const float* dL_dout_language_feature;           /* F*H*W  (include_feature) */
    const float* dL_dout_language_feature_instance;  /* Fi*H*W (include_feature) */
    const float* dL_dout_all_map;                    /* 5*H*W  (render_geo) */
    const float* dL_dout_plane_depth;                /* H*W    (render_geo) */
    /* gradient outputs (uninitialised on entry, fully written on return) */
    float* dL_dmeans2D;      /* P*3, .z = 0 */
    float* dL_dmeans2D_abs;  /* P*3, .z = 0 */
    float* dL_dconic;        /* P*4 as {x,y,0,w} (backward.cu:669-671) */
    float* dL_dopacity;      /* P */
    float* dL_dcolors;       /* P*3 */
    float* dL_dlanguage_feature;           /* P*F  if include_feature else untouched */
    float* dL_dlanguage_feature_instance;  /* P*Fi if include_feature else untouched */
    float* dL_dmeans3D;      /* P*3 */
    float* dL_dcov3D;        /* P*6 */
    float* dL_dsh;           /* P*M*3 (may be NULL when M == 0) */
    float* dL_dscales;       /* P*3 */
    float* dL_drotations;    /* P*4 */
    float* dL_dall_map;      /* P*5 */
    void* stream;
    /* Multi-view gradient accumulation (no reference counterpart: the reference returns fresh tensors and lets autograd add
     * them).  Bit mask (LSX_ACC_*, below) of the PARAMETER gradients that are ADDED to their buffer's current contents
     * (which must then be initialised) instead of overwriting it; 0 = overwrite everything.  The per-view screen-space
     * outputs dL_dmeans2D, dL_dmeans2D_abs and dL_dconic are always overwritten.  (ABI v5: was a boolean for all groups.) */
    int32_t accumulate_param_grads;
    /* ABI v5: size in bytes of the binning buffer the forward call allocated through binning_alloc (its list capacity is
     * recovered from it); 0 = the forward call ran without a capacity hint (capacity = R rounded up to 64). */
    uint64_t binning_bytes;
    /* ABI v5 — fused render-wrapper mode, backward (the forward call must have used the same raw_params / pose).  The inputs
     * means3D / scales / rotations are the RAW parameters; dL_dmeans3D, dL_dscales, dL_drotations and dL_dopacity receive the
     * gradients of the RAW parameters (the chain rule through activations, plane normal, all_map and the pose transform is
     * applied in the per-Gaussian kernel: what lsx_gaussian_head_backward + lsx_pose_transform_backward compute in two extra
     * passes); dL_dcov3D and dL_dall_map may be NULL (not written).  dL_dpose (7 floats, may be NULL) receives the pose
     * gradient — added to its contents when accumulate_pose != 0. */
    int32_t raw_params;
    const float* pose;
    float* dL_dpose;
    int32_t accumulate_pose;
} lsx_backward_args;

#define LSX_ACC_MEANS3D 0x001
#define LSX_ACC_SH 0x002
#define LSX_ACC_OPACITY 0x004
#define LSX_ACC_SCALES 0x008
#define LSX_ACC_ROTATIONS 0x010
#define LSX_ACC_COLORS 0x020
#define LSX_ACC_LANG 0x040
#define LSX_ACC_INST 0x080
#define LSX_ACC_ALL_MAP 0x100
#define LSX_ACC_COV3D 0x200
#define LSX_ACC_ALL 0x3ff

LSX_API int lsx_rasterize_backward(const lsx_backward_args* args);

/* present[i] = (view-space z of means3D[i]) > 0.2   (auxiliary.h:139-164, rasterizer_impl.cu:54-66) */
LSX_API int lsx_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream);

/* out[i] = mean of the 3 smallest squared distances from points[i] to the other points
 * (simple_knn.cu:147-183).  Scratch comes from `alloc` (one call).  */
LSX_API int lsx_knn_mean_dist2(int32_t P, const float* points, float* out, lsx_alloc_fn alloc, void* alloc_user,
                       void* stream);

/* ---- next row: plane depth -> depth normal ---------------------------------------------------
 * Replaces render_normal() -> normal_from_depth_image() -> depth_pcd2normal() (offset == None) of the render wrapper
 * (field_construction/gaussian_renderer/__init__.py:28-40, field_construction/utils/graphics_utils.py:16-75) together with
 * the `* rendered_alpha.detach()` of its call site (gaussian_renderer/__init__.py:233-235).
 *   depth       H*W   plane depth
 *   alpha       H*W   or NULL: per-pixel factor multiplied into the normal (treated as a constant: no gradient)
 *   out_normal  3*H*W planar; zero on the one-pixel border
 * K = [[fx,0,cx],[0,fy,cy],[0,0,1]] as built by Camera.get_calib_matrix_nerf (field_construction/scene/cameras.py:153-156). */
LSX_API int lsx_depth_normal_forward(int32_t W, int32_t H, float fx, float fy, float cx, float cy, const float* depth,
                                     const float* alpha, float* out_normal, void* stream);
/* dL_ddepth (H*W, fully written) from dL_dnormal (3*H*W); same depth / alpha as the forward call. */
LSX_API int lsx_depth_normal_backward(int32_t W, int32_t H, float fx, float fy, float cx, float cy, const float* depth,
                                      const float* alpha, const float* dL_dnormal, float* dL_ddepth, void* stream);

/* ---- next row: fused L1 + SSIM image loss ------------------------------------------------------
 * Replaces l1_loss and ssim (11x11 Gaussian window, sigma 1.5, zero padding; field_construction/utils/loss_utils.py:20-75)
 * as combined at field_construction/gaussian_field.py:238-246.  Images are planar C*H*W fp32.
 * forward : writes three derivative maps (3*C*H*W floats, consumed by backward) and per-block partial sums
 *           partial[0..nblk) = sums of the SSIM map, partial[nblk..2 nblk) = sums of |img1 - img2|, nblk =
 *           lsx_image_loss_num_blocks(C,H,W); the caller adds them up (deterministic) and divides by C*H*W.
 * backward: dL_dssim_mean / dL_dl1_mean are DEVICE scalars (may be NULL = 0): the upstream gradients of the two means, e.g.
 *           -l * g and (1-l) * g for loss = (1-l) * mean|.| + l * (1 - mean ssim); dL_dimg1 (C*H*W) is fully written. */
LSX_API int64_t lsx_image_loss_num_blocks(int32_t C, int32_t H, int32_t W);
LSX_API int lsx_image_loss_forward(int32_t C, int32_t H, int32_t W, const float* img1, const float* img2, float* dmaps,
                                   float* partial, void* stream);
LSX_API int lsx_image_loss_backward(int32_t C, int32_t H, int32_t W, const float* img1, const float* img2, const float* dmaps,
                                    const float* dL_dssim_mean, const float* dL_dl1_mean, float* dL_dimg1, void* stream);

/* ---- next row (per-Gaussian half of the render wrapper) -----------------------------------------------------
 * forward : scales = exp(scaling_raw) (P*3), rotations = normalize(rotation_raw) (P*4), opacity = sigmoid(opacity_raw) (P),
 *           all_map (P*5) = [plane normal in camera space, 1, |<normal, camera-space position>|], the plane normal being the
 *           rotation-matrix column of the smallest scale, flipped towards the camera
 *           (field_construction/scene/gaussian_model.py:53-61,193-236; field_construction/gaussian_renderer/__init__.py:188-196).
 * backward: gradients w.r.t. the raw parameters from the gradients the rasterizer returns for scales / rotations / opacities /
 *           all_map (any of them may be NULL = zero); dL_dmeans3D (may be NULL) is added to the position gradient.
 * viewmatrix (16 floats, the reference's world_view_transform in its own memory order) and campos (3) are DEVICE arrays, as in
 * lsx_forward_args. */
LSX_API int lsx_gaussian_head_forward(int32_t P, const float* viewmatrix, const float* campos, const float* xyz,
                                      const float* scaling_raw, const float* rotation_raw, const float* opacity_raw,
                                      float* scales, float* rotations, float* opacity, float* all_map, void* stream);
LSX_API int lsx_gaussian_head_backward(int32_t P, const float* viewmatrix, const float* campos, const float* xyz,
                                       const float* scaling_raw, const float* rotation_raw, const float* opacity_raw,
                                       const float* dL_dscales, const float* dL_drotations, const float* dL_dopacity,
                                       const float* dL_dall_map, const float* dL_dmeans3D, float* dL_dxyz,
                                       float* dL_dscaling_raw, float* dL_drotation_raw, float* dL_dopacity_raw, void* stream);
/* Same, with `accumulate_mask`: bit 0 dL_dxyz, bit 1 dL_dscaling_raw, bit 2 dL_drotation_raw, bit 3 dL_dopacity_raw are ADDED
 * to the buffers' contents (the views of a multi-view gradient arena) instead of written. */
LSX_API int lsx_gaussian_head_backward_acc(int32_t P, const float* viewmatrix, const float* campos, const float* xyz,
                                           const float* scaling_raw, const float* rotation_raw, const float* opacity_raw,
                                           const float* dL_dscales, const float* dL_drotations, const float* dL_dopacity,
                                           const float* dL_dall_map, const float* dL_dmeans3D, float* dL_dxyz,
                                           float* dL_dscaling_raw, float* dL_drotation_raw, float* dL_dopacity_raw,
                                           int32_t accumulate_mask, void* stream);

/* Adam step over flat fp32 arenas (parameters, gradients, both moments share one layout of n elements).  Replaces
 * torch.optim.Adam(groups, lr=0.0, eps=1e-15).step() of the reference (field_construction/scene/gaussian_model.py:313-328,
 * field_construction/gaussian_field.py:537-543): default betas, no weight decay, no amsgrad, one learning rate per group.
 * group_begin_host[0..n_groups] are ascending element offsets (multiples of 4; elements outside [begin[0], begin[n_groups])
 * get lr 0 but their moments are still updated), group_lr_host[0..n_groups) the learning rates; both are HOST arrays.
 * `step` is the 1-based step count used for the bias corrections. */
#define LSX_ADAM_MAX_GROUPS 16
LSX_API int lsx_arena_adam_step(int64_t n, int32_t n_groups, const int64_t* group_begin_host, const float* group_lr_host,
                                int32_t step, float beta1, float beta2, float eps, float* params, const float* grads,
                                float* exp_avg, float* exp_avg_sq, void* stream);

/* ---- next row: densification / pruning on the flat arenas -------------------------------------------------
 * Replaces GaussianModel.add_densification_stats / densify_and_prune / reset_opacity and the optimizer-state surgery under them
 * (field_construction/scene/gaussian_model.py:443-446,506-724; call sites field_construction/gaussian_field.py:519-535).
 *
 * lsx_densify_stats_update: one view's statistics.  For rows with radii > 0: grad_accum += |dL_dmeans2D.xy|, grad_accum_abs +=
 *   |dL_dmeans2D_abs.xy|, denom += 1 (the reference's denom_abs always equals denom); max_radii2D = max(max_radii2D, radii)
 *   where additionally out_observe > 0 (NULL = no such filter).  dL_dmeans2D* are the (P,3) tensors the rasterizer returns. */
LSX_API int lsx_densify_stats_update(int32_t P, const float* dL_dmeans2D, const float* dL_dmeans2D_abs, const int32_t* radii,
                                     const int32_t* out_observe, float* grad_accum, float* grad_accum_abs, float* denom,
                                     float* max_radii2D, void* stream);

/* lsx_densify_plan: decides, for the P current rows, which are cloned, split (N = 2) and pruned, exactly as
 * densify_and_prune(max_grad, abs_max_grad, min_opacity, extent, max_screen_size) does, and lays out the rows of the new set:
 *   [surviving originals | surviving clones | surviving first children | surviving second children]  (the reference's order).
 * Device scratch: `workspace` of lsx_densify_workspace_bytes(P) bytes, which also holds the two result arrays
 *   row_map[P_new]     = kind << 30 | source row   (kind 0 original, 1 clone, 2 / 3 first / second split child)
 *   noise_index[P_new] = row of z_clone (kind 1) / z_split (kind 2, 3) the reference's sampling order assigns, -1 for kind 0
 * valid until the workspace is released.  Synchronises the stream (the host needs the counts: 2 reads, +1 per capped branch).
 * Thresholds must be > 0.  prune_world_size = bool(max_screen_size): the reference zeroes max_radii2D before its final prune,
 * so the screen-size value itself never matters and only the world-size test (> 0.1 * extent) is switched — reproduced. */
typedef struct lsx_densify_plan_args {
    int32_t P;
    const float* grad_accum;      /* (P) */
    const float* grad_accum_abs;  /* (P) */
    const float* denom;           /* (P) */
    const float* max_radii2D;     /* (P) */
    const float* scaling_raw;     /* (P,3) log scales  */
    const float* opacity_raw;     /* (P)   logit opacity */
    float max_grad, abs_max_grad, min_opacity, extent;
    float percent_dense, abs_split_radii2D_threshold;
    int64_t max_all_points, max_abs_split_points;
    int32_t prune_world_size;
    void* workspace;
    size_t workspace_bytes;
    void* stream;
} lsx_densify_plan_args;

typedef struct lsx_densify_plan_result {
    int64_t P_new;
    int64_t n_clone, n_split;  /* selected (pre-prune): z_clone needs n_clone rows, z_split 2 * n_split rows */
    int64_t n_split_abs;       /* of n_split, selected by the abs-gradient pass */
    int64_t n_kept_original, n_kept_clone, n_kept_split;
    int32_t clone_capped, split_capped, abs_capped;  /* which max_all_points / max_abs_split_points branches were taken */
    const uint32_t* row_map;      /* device, inside the workspace */
    const int32_t* noise_index;   /* device, inside the workspace */
} lsx_densify_plan_result;

LSX_API size_t lsx_densify_workspace_bytes(int32_t P);
LSX_API int lsx_densify_plan(const lsx_densify_plan_args* args, lsx_densify_plan_result* out);

/* lsx_densify_apply: builds the new arenas with ONE gather per group: new[new_begin[k] + dst * width[k] + c] =
 * old[old_begin[k] + src * width[k] + c]; exp_avg / exp_avg_sq likewise for surviving originals and zero for new rows (all four
 * moment pointers may be NULL).  Roles: new rows of the XYZ group get xyz + R(q/|q|) (exp(s) * z) (z = unit normal noise,
 * (n_clone,3) and (2 n_split,3) as the reference draws it: first children, then second children), split children of the SCALING
 * group get log(exp(s) / 1.6); ROTATION marks the quaternion group.  old_begin / new_begin / width / role are HOST arrays. */
enum lsx_densify_role { LSX_DENSIFY_ROLE_COPY = 0, LSX_DENSIFY_ROLE_XYZ = 1, LSX_DENSIFY_ROLE_SCALING = 2,
                        LSX_DENSIFY_ROLE_ROTATION = 3 };
typedef struct lsx_densify_apply_args {
    int64_t P_new;
    int64_t n_new_rows;           /* P_new - n_kept_original */
    int32_t n_groups;             /* <= LSX_ADAM_MAX_GROUPS */
    const int64_t* old_begin;     /* element offsets of the groups inside the old / new arenas */
    const int64_t* new_begin;
    const int32_t* width;         /* floats per row */
    const int32_t* role;
    const uint32_t* row_map;
    const int32_t* noise_index;
    const float* z_clone;
    const float* z_split;
    const float* old_params;
    const float* old_exp_avg;
    const float* old_exp_avg_sq;
    float* new_params;
    float* new_exp_avg;
    float* new_exp_avg_sq;
    void* stream;
    int32_t group_align;          /* > 1: every group of the NEW arenas is zero-filled up to the next multiple of this many
                                   * elements past its last row (the arenas may then be allocated uninitialised) */
} lsx_densify_apply_args;
LSX_API int lsx_densify_apply(const lsx_densify_apply_args* args);

/* opacity_raw = inverse_sigmoid(min(sigmoid(opacity_raw), 0.01)), moments of that group zeroed (NULL = absent)
 * (GaussianModel.reset_opacity, gaussian_model.py:443-446 + replace_tensor_to_optimizer :506-518). */
LSX_API int lsx_reset_opacity(int32_t P, float* opacity_raw, float* exp_avg, float* exp_avg_sq, void* stream);

/* ---- next row (pose half of the render wrapper) -------------------------------------------------------------
 * render(..., camera_pose=pose) of field_construction/gaussian_renderer/__init__.py:79-87 with pose = [quaternion(4) | T(3)]
 * (a DEVICE array of 7 floats, one row of GaussianModel.P):
 *   out_means3D   = R(q/|q|) xyz + T              (get_camera_from_tensor / quad2rotation, field_construction/utils/pose_utils.py:13-87)
 *   out_rotations = q (x) rotation_raw            (quadmultiply, pose_utils.py:89-107: Hamilton product with the RAW pose quaternion)
 * rotation_raw / out_rotations may both be NULL.  backward: dL_dxyz (P,3), dL_drotation_raw (P,4, may be NULL) and dL_dpose (7)
 * from the gradients the rasterizer returns for means3D / rotations (either may be NULL = zero); `partials` is device scratch of
 * lsx_pose_num_partials() floats.  The pose gradient is reduced in a fixed order (deterministic, no atomics). */
LSX_API int32_t lsx_pose_num_partials(void);
LSX_API int lsx_pose_transform_forward(int32_t P, const float* pose, const float* xyz, const float* rotation_raw,
                                       float* out_means3D, float* out_rotations, void* stream);
LSX_API int lsx_pose_transform_backward(int32_t P, const float* pose, const float* xyz, const float* rotation_raw,
                                        const float* dL_dmeans3D, const float* dL_drotations, float* dL_dxyz,
                                        float* dL_drotation_raw, float* dL_dpose, float* partials, void* stream);

/* ---- next row: masked L1 of the language-feature loss ------------------------------------------------------
 * l1_loss(a * mask, b * mask) = mean |a*mask - b*mask| over C*H*W (field_construction/gaussian_field.py:450-451,
 * field_construction/utils/loss_utils.py:20-21); a, b planar C*H*W, mask H*W (mask_channels 1, broadcast) or C*H*W, NULL = ones.
 * forward : partial[0 .. lsx_masked_l1_num_blocks(C*H*W)) = per-block sums; the caller adds them and divides by C*H*W.
 * backward: dL_da = *upstream / (C*H*W) * sign(a*mask - b*mask) * mask, fully written; upstream is a DEVICE scalar (NULL = 1). */
LSX_API int32_t lsx_masked_l1_num_blocks(int64_t n);
LSX_API int lsx_masked_l1_forward(int32_t C, int32_t H, int32_t W, int32_t mask_channels, const float* a, const float* b,
                                  const float* mask, float* partial, void* stream);
LSX_API int lsx_masked_l1_backward(int32_t C, int32_t H, int32_t W, int32_t mask_channels, const float* a, const float* b,
                                   const float* mask, const float* upstream, float* dL_da, void* stream);

/* ---- next row (second half): 3-D neighbourhood regulariser of the language / instance features --------------
 * loss_cls_3d(features = xyz (N,3), predictions (N,C), k, lambda_val, max_points, sample_size) of
 * field_construction/utils/loss_utils.py:158-186 (called at field_construction/gaussian_field.py:461-465,482-485):
 *   q = (predictions - min) / (max - min) over all elements (only if max > min); for each of the S sampled rows the k
 *   nearest points (exact squared distances, the sample itself included, ties towards the lower index);
 *   loss = lambda_val * mean over S*k*C of | q[s] * (log(q[s] + 1e-10) - log(q[nbr] + 1e-10)) |.
 * The random choices of the reference (torch.randperm) stay with the caller: sample_idx is a DEVICE array of S row indices.
 * forward : *loss (device scalar), nbr_idx (S,k) int32 and minmax[2] are written and must be handed to backward;
 *           scratch = device buffer of lsx_cls3d_scratch_bytes() bytes (contents need not survive until backward).
 * backward: dL_dpreds (N,C) fully written (zero rows for points that are neither sampled nor a neighbour), including the
 *           gradient through min and max (split evenly among tied elements, like torch); upstream = DEVICE scalar (NULL = 1).
 * 1 <= k <= 8 and k <= N, else -1.  No S x N distance matrix is materialised and the host never waits. */
LSX_API int64_t lsx_cls3d_scratch_bytes(int32_t N, int32_t C, int32_t S, int32_t k);
LSX_API int lsx_cls3d_forward(int32_t N, int32_t C, int32_t S, int32_t k, float lambda_val, const float* points,
                              const float* preds, const int32_t* sample_idx, float* loss, int32_t* nbr_idx, float* minmax,
                              void* scratch, void* stream);
/* The same with the neighbour search through a prebuilt search structure over `points` (Morton order + 3-level box hierarchy,
 * the one lsx_knn_mean_dist2 builds for itself): build it once with lsx_knn_tree_build into a device buffer of
 * lsx_knn_tree_bytes(N) bytes and reuse it for every call made before the points change — e.g. once per optimisation step for
 * all views of the step.  Exact: the same neighbours, in the same order, as the brute-force search of lsx_cls3d_forward. */
LSX_API int64_t lsx_knn_tree_bytes(int32_t N);
LSX_API int lsx_knn_tree_build(int32_t N, const float* points, void* tree, void* stream);
LSX_API int lsx_cls3d_forward_tree(int32_t N, int32_t C, int32_t S, int32_t k, float lambda_val, const float* points,
                           const float* preds, const int32_t* sample_idx, float* loss, int32_t* nbr_idx, float* minmax,
                           void* scratch, const void* tree, void* stream);
LSX_API int lsx_cls3d_backward(int32_t N, int32_t C, int32_t S, int32_t k, float lambda_val, const float* preds,
                               const int32_t* sample_idx, const int32_t* nbr_idx, const float* minmax, const float* upstream,
                               float* dL_dpreds, void* scratch, void* stream);

/* ---- next row: row pack / unpack for the on-disk formats ----------------------------------------------------
 * GaussianModel.save_ply / load_ply (field_construction/scene/gaussian_model.py:415-441,448-504) convert between the per-group
 * parameter tensors and one row of ncols floats per Gaussian.  rows[r * ncols + c] = arena[col_begin[c] + r * col_stride[c]]
 * (col_begin[c] < 0: the column is constant 0 on pack and ignored on unpack — the PLY's unused normals).  col_begin / col_stride
 * are DEVICE arrays of ncols entries; rows is a device buffer of P * ncols floats (the PLY body, ready for one D2H copy). */
LSX_API int lsx_rows_pack(int64_t P, int32_t ncols, const int64_t* col_begin, const int32_t* col_stride, const float* arena,
                          float* rows, void* stream);
LSX_API int lsx_rows_unpack(int64_t P, int32_t ncols, const int64_t* col_begin, const int32_t* col_stride, const float* rows,
                            float* arena, void* stream);

/* ---- parity / introspection helpers (used by the tests; not on the hot path) ------------------- */

/* Offsets (bytes from the buffer base) of the private scratch arrays, so tests can read the
 * bit-exact intermediates the way SURVEY.md Appendix B reads the reference's. */
typedef struct lsx_scratch_layout {
    /* geometry buffer, P rows */
    size_t depths;         /* f32[P]   view-space z                              */
    size_t clamped;        /* u8[P]    bit c set if SH colour channel c was clamped at 0 */
    size_t means2D;        /* f32[2P]  pixel-space centre                        */
    size_t cov3D;          /* f32[6P]                                            */
    size_t conic_opacity;  /* f32[4P]                                            */
    size_t rgb;            /* f32[3P]                                            */
    size_t tiles_touched;  /* u32[P]                                             */
    size_t records;        /* f32[P*record_stride] packed per-Gaussian blend record */
    int32_t record_stride; /* floats */
    size_t geom_bytes;
    /* image buffer */
    size_t final_T;        /* f32[W*H] */
    size_t n_contrib;      /* u32[W*H] */
    size_t ranges;         /* u32[2*tiles] (start,end) */
    size_t image_bytes;
    /* binning buffer, R rows */
    size_t point_list;     /* u32[R] sorted Gaussian indices */
    size_t binning_bytes;
    size_t masks;          /* u8[R]  sub-tile footprint mask per list entry (binning buffer) */
    /* ABI v4: the per-block compacted lists the render kernels walk (built from the masks) */
    size_t blk_list;       /* u32[8*R] binning buffer: for block w (8x4 pixels) of tile t, the positions inside the tile's range
                            *          of the entries with mask bit w set, in list order, at [w*R + ranges[t].start + k] */
    size_t blk_cnt;        /* u32[8*tiles] binning buffer: list lengths, [8*t + w] */
    size_t k_contrib;      /* u32[W*H] image buffer: elements of the pixel's block list up to its last contributor */
} lsx_scratch_layout;

LSX_API int lsx_scratch_layout_query(int32_t P, int32_t W, int32_t H, int32_t R, int32_t n_blend_channels,
                             lsx_scratch_layout* out);

/* Materialise the sorted 64-bit keys (tile << 32 | depth bits) the reference keeps in
 * BinningState::point_list_keys (rasterizer_impl.cu:102-106), from this library's scratch. */
LSX_API int lsx_debug_sorted_keys(int32_t P, int32_t W, int32_t H, int32_t R, int32_t n_blend_channels,
                          const char* geom_buffer, const char* binning_buffer, size_t binning_bytes,
                          const char* image_buffer, uint64_t* keys_out, void* stream);

/* List capacity (a multiple of 64, >= num_rendered) of a binning buffer of `binning_bytes` bytes allocated by a forward call at
 * this image size, or -1 if no capacity has that size.  lsx_scratch_layout_query takes this value as its R. */
LSX_API int32_t lsx_binning_capacity(size_t binning_bytes, int32_t W, int32_t H);

/* Workload counters of one rendered view, counted on the device from the scratch of a forward call (S, B, R
 * accompany every number).  stats_out: DEVICE array of 8 uint64, written as
 *   [0] S   sum of n_contrib = (pixel, entry) tests per pass of the reference's render loops (forward.cu:335-407)
 *   [1] B   (pixel, entry) pairs actually blended (alpha >= 1/255, power <= 0, at or before the pixel's last contributor)
 *   [2] V   (8x4 block, entry) visits of this library's backward pass      [3] Vb  visits in which some pixel blends
 *   [4] L   total length of the per-block compacted lists
 *   [5] Hmax, [6] Hsum   what-if counters for 4x4-pixel cells (two per block): sum over blocks of the larger / of both halves'
 *                        counts of visits in which that half blends                                   [7] reserved (0) */
LSX_API int lsx_render_stats(int32_t P, int32_t W, int32_t H, int32_t R, int32_t n_blend_channels, const char* geom_buffer,
                     const char* binning_buffer, size_t binning_bytes, const char* image_buffer, uint64_t* stats_out,
                     void* stream);

/* number of kernels launched by this library since process start (bench.py "gpu_launches") */
LSX_API uint64_t lsx_kernel_launch_count(void);

/* ---- optional per-stage device timing (events on the caller's stream; used by bench.py) ---------- */
enum lsx_stage {
    LSX_STAGE_PREPROCESS_FWD = 0, LSX_STAGE_DEPTH_SORT, LSX_STAGE_OFFSETS_SCAN, LSX_STAGE_EMIT, LSX_STAGE_TILE_SORT,
    LSX_STAGE_TILE_RANGES, LSX_STAGE_FOOTPRINT_MASKS, LSX_STAGE_RENDER_FWD, LSX_STAGE_BWD_ZERO, LSX_STAGE_RENDER_BWD, LSX_STAGE_PREPROCESS_BWD,
    LSX_STAGE_KNN, LSX_NUM_STAGES
};
LSX_API void lsx_profile_enable(int enable);
/* Synchronises the recorded events, writes the milliseconds accumulated per stage since the last read into
 * ms_out[0..n) and returns the number of spans consumed. */
LSX_API int lsx_profile_read(float* ms_out, int n);

/* Device peaks MEASURED_PEAKS.json lacks.  kind 0: FP32 FFMA TFLOP/s, 1: MUFU.EX2 Gop/s, 2: global
 * RED.ADD.F32 Gop/s on random addresses of a 64 MiB array.  `scratch`: device buffer of >= 64 MiB. */
LSX_API int lsx_microbench(int kind, float* scratch, size_t scratch_bytes, double* result, void* stream);

LSX_API const char* lsx_last_error(void);
LSX_API int lsx_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LSX_RASTERIZER_H_INCLUDED */

// common.cuh — shared device helpers for the B200-native LangSurf rasterizer.
// Everything here is written for sm_100a only (no fallback paths).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <stdint.h>
#include <stddef.h>

namespace lsx {

#ifndef LSX_PRE_BWD_THREADS
#define LSX_PRE_BWD_THREADS 128  // Gaussians per block of the preprocess backward kernel (also sizes its pose-partial rows)
#endif
constexpr int TILE_X = 16;  // tile geometry is part of the binning contract (config.h:19-20 in the reference)
constexpr int TILE_Y = 16;
constexpr int TILE_PIXELS = TILE_X * TILE_Y;

// ---- host-side bookkeeping -------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define LSX_CUDA_OK(expr)                                                                        \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            lsx::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                  \
                           cudaGetErrorString(_e));                                              \
            return -2;                                                                           \
        }                                                                                        \
    } while (0)

// checks the launch itself; with debug also synchronises the stream (auxiliary.h:166-173 semantics)
#define LSX_KERNEL_OK(stream, debug)                                                             \
    do {                                                                                         \
        lsx::count_launch();                                                                     \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e == cudaSuccess && (debug)) _e = cudaStreamSynchronize(stream);                    \
        if (_e != cudaSuccess) {                                                                 \
            lsx::set_error("kernel failed at %s:%d: %s", __FILE__, __LINE__,                     \
                           cudaGetErrorString(_e));                                              \
            return -3;                                                                           \
        }                                                                                        \
    } while (0)

// ---- programmatic dependent launch (sm_90+) ---------------------------------------------------------------------------
// A kernel launched with launch_pdl() may be scheduled while its predecessor in the stream is still running: its blocks
// become resident as the predecessor's last wave drains instead of after the grid has been retired (measured on the
// radix passes: ~1.2 us per kernel boundary, profiles/r8n_pdl_radix_only.log).  Such a kernel executes pdl_wait() before it
// touches global memory: it returns once every prerequisite grid has completed and its writes are visible — ordinary stream
// semantics from there on — and lets the NEXT kernel of the stream be scheduled in the same way.
// Used for the chains of SHORT kernels only (radix passes, scans, tile ranges: 14 of a forward call's launches).  On the
// multi-wave kernels (footprint, forward render, preprocess_bwd behind the backward render) the stage timers improve too,
// but the real back-to-back step gets 1 % slower at C3 (profiles/r8q_pdl_bench_ab.log, r8r_pdl_bench_ab2.log): removed.
#ifndef LSX_PDL
#define LSX_PDL 1
#endif
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() {
#if LSX_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
template <typename... KArgs, typename... Args>
static inline void launch_pdl(void (*kernel)(KArgs...), long long grid, int block, size_t smem, cudaStream_t stream,
                              Args... args) {
#if LSX_PDL
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface in LSX_KERNEL_OK
#else
    kernel<<<(unsigned)grid, block, smem, stream>>>(static_cast<KArgs>(args)...);
#endif
}
#endif

// ---- LSX_BOUNDS_CHECK build (the pool refuses compute-sanitizer: profiles/r6d_compute_sanitizer_refused.txt) ------
// -DLSX_BOUNDS_CHECK=1 turns every data-dependent index of the binning / list / record structures into a checked access:
// an out-of-range index prints its site and traps (the launch then fails and the C ABI returns an error).  The default build
// compiles the checks away.  tools/ab_variants.py builds the checked library; the GPU test tier is run through it once per
// round (profiles/r6i_bounds_check_run.log).
#ifndef LSX_BOUNDS_CHECK
#define LSX_BOUNDS_CHECK 0
#endif
#if LSX_BOUNDS_CHECK
#define LSX_CHECK_INDEX(i, n, what)                                                                              \
    do {                                                                                                         \
        if (!((unsigned long long)(i) < (unsigned long long)(n))) {                                              \
            printf("LSX_BOUNDS_CHECK: %s index %lld outside [0, %lld) at %s:%d (block %d thread %d)\n", what,    \
                   (long long)(i), (long long)(n), __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x);       \
            __trap();                                                                                            \
        }                                                                                                        \
    } while (0)
#else
#define LSX_CHECK_INDEX(i, n, what) ((void)0)
#endif

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- blend-record geometry --------------------------------------------------------------------
// One packed record per Gaussian: head {x, y, conic.x, conic.y, conic.z, opacity, 0, 0} followed by
// the blended channels {rgb(3), language(F), instance(Fi), all_map(5)} zero-padded to a multiple of 4.
// The stride is a multiple of 8 floats so that a record is a whole number of 32-B DRAM/L2 sectors.
constexpr int REC_HEAD = 8;
__host__ __device__ static inline int round_up4(int c) { return (c + 3) & ~3; }
__host__ __device__ static inline int record_stride(int n_blend_channels) {
    return (REC_HEAD + round_up4(n_blend_channels) + 7) & ~7;
}

// ---- small device math (expression trees match the reference's scalar order on purpose:
//      radii / tile rects / depth keys are part of the bit-exact contract) ----------------------
struct Mat3 {  // column-major like the reference's math library: m[c][r]
    float m[3][3];
};

__device__ __forceinline__ Mat3 mat3_mul(const Mat3& a, const Mat3& b) {
    // result[c][r] = a[0][r]*b[c][0] + a[1][r]*b[c][1] + a[2][r]*b[c][2]   (left-to-right sums)
    Mat3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int row = 0; row < 3; ++row)
            r.m[c][row] = a.m[0][row] * b.m[c][0] + a.m[1][row] * b.m[c][1] + a.m[2][row] * b.m[c][2];
    return r;
}

__device__ __forceinline__ Mat3 mat3_transpose(const Mat3& a) {
    Mat3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int row = 0; row < 3; ++row) r.m[c][row] = a.m[row][c];
    return r;
}

__device__ __forceinline__ float3 xform_point_4x3(const float3& p, const float* __restrict__ m) {
    return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
                       m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}

__device__ __forceinline__ float4 xform_point_4x4(const float3& p, const float* __restrict__ m) {
    return make_float4(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
                       m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
                       m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]);
}

// pixel centre from NDC; evaluated in double exactly like auxiliary.h:41-44
__device__ __forceinline__ float ndc_to_pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

// tile rectangle of a splat (auxiliary.h:46-56): float divide by the tile size, C truncation, clamp
__device__ __forceinline__ void tile_rect(const float2 p, int max_radius, uint2& rect_min, uint2& rect_max,
                                          const uint32_t grid_x, const uint32_t grid_y) {
    rect_min.x = min(grid_x, (uint32_t)max((int)0, (int)((p.x - max_radius) / TILE_X)));
    rect_min.y = min(grid_y, (uint32_t)max((int)0, (int)((p.y - max_radius) / TILE_Y)));
    rect_max.x = min(grid_x, (uint32_t)max((int)0, (int)((p.x + max_radius + TILE_X - 1) / TILE_X)));
    rect_max.y = min(grid_y, (uint32_t)max((int)0, (int)((p.y + max_radius + TILE_Y - 1) / TILE_Y)));
}

// Gaussian falloff exponent and opacity, with the exact operation sequence the reference's kernels execute
// (SASS of forward.cu:363 / backward.cu:564 compiled by nvcc 12.9 for sm_100a):
//   qz = (dy*cz)*dy ; qx = dx*cx ; cross = (dx*cy)*dy ; sum = fma(dx, qx, qz) ; power = fma(sum, -0.5, -cross)
// Written with explicit round-to-nearest intrinsics so that the compiler cannot pick another contraction:
// alpha / T thresholds (1/255, 1e-4, 0.5) then flip for exactly the same (pixel, Gaussian) pairs, which makes
// n_contrib, final_T and out_observe reproducible bit for bit.
__device__ __forceinline__ float splat_power(const float cx, const float cy, const float cz, const float dx,
                                             const float dy) {
    const float qz = __fmul_rn(__fmul_rn(dy, cz), dy);
    const float qx = __fmul_rn(dx, cx);
    const float cross = __fmul_rn(__fmul_rn(dx, cy), dy);
    const float sum = __fmaf_rn(dx, qx, qz);
    return __fmaf_rn(sum, -0.5f, -cross);
}
__device__ __forceinline__ float splat_alpha(const float opacity, const float gauss) {
    return fminf(0.99f, __fmul_rn(opacity, gauss));
}

// real spherical-harmonics constants, degrees 0..3 (auxiliary.h:22-39)
__device__ constexpr float kSH0 = 0.28209479177387814f;
__device__ constexpr float kSH1 = 0.4886025119029199f;
__device__ constexpr float kSH2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                      -1.0925484305920792f, 0.5462742152960396f};
__device__ constexpr float kSH3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                      0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                      -0.5900435899266435f};

}  // namespace lsx

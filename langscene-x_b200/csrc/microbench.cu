// microbench.cu — device peaks MEASURED_PEAKS.json does not carry: FP32 FFMA issue rate,
// MUFU.EX2 rate and global fp32 reduction (RED.ADD.F32) rate.  bench.py reports roofline fractions of the
// render kernels against these measured numbers.
#include "../../include/lsx_rasterizer.h"
#include "kernels.cuh"

namespace lsx {
namespace {

__global__ void __launch_bounds__(256) ffma_kernel(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f,
          a7 = a0 + 7.f;
    const float b = 1.000001f, c = 1e-7f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a0 = __fmaf_rn(a0, b, c); a1 = __fmaf_rn(a1, b, c); a2 = __fmaf_rn(a2, b, c); a3 = __fmaf_rn(a3, b, c);
            a4 = __fmaf_rn(a4, b, c); a5 = __fmaf_rn(a5, b, c); a6 = __fmaf_rn(a6, b, c); a7 = __fmaf_rn(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) ex2_kernel(float* out, int iters) {
    float a0 = -threadIdx.x * 1e-3f, a1 = a0 - .1f, a2 = a0 - .2f, a3 = a0 - .3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
            a0 -= 1.f; a1 -= 1.f; a2 -= 1.f; a3 -= 1.f;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}

__global__ void __launch_bounds__(256) red_kernel(float* buf, uint32_t mask, int iters) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        atomicAdd(buf + ((x >> 4) & mask), 1.0f);  // result unused -> RED.E.ADD.F32
    }
}

template <typename F>
double best_ms(F&& launch, cudaStream_t stream) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 1e30;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0, stream);
        launch();
        cudaEventRecord(e1, stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

}  // namespace
}  // namespace lsx

using namespace lsx;

// kind 0: FP32 FFMA TFLOP/s (2 flops per FFMA); 1: MUFU.EX2 Gop/s; 2: global RED.ADD.F32 Gop/s over a
// 64 MiB (L2-resident) float array with pseudo-random addresses.  `scratch` must hold >= 64 MiB.
extern "C" int lsx_microbench(int kind, float* scratch, size_t scratch_bytes, double* result, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!scratch || !result || scratch_bytes < (64u << 20)) {
        set_error("lsx_microbench: scratch of >= 64 MiB required");
        return -1;
    }
    int dev = 0, sms = 0;
    LSX_CUDA_OK(cudaGetDevice(&dev));
    LSX_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8;
    if (kind == 0) {
        const int iters = 4096;
        const double ms = best_ms([&] { ffma_kernel<<<blocks, 256, 0, stream>>>(scratch, iters); count_launch(); }, stream);
        *result = 2.0 * 64.0 * iters * blocks * 256.0 / (ms * 1e-3) / 1e12;
    } else if (kind == 1) {
        const int iters = 2048;
        const double ms = best_ms([&] { ex2_kernel<<<blocks, 256, 0, stream>>>(scratch, iters); count_launch(); }, stream);
        *result = 32.0 * iters * blocks * 256.0 / (ms * 1e-3) / 1e9;
    } else if (kind == 2) {
        const int iters = 256;
        LSX_CUDA_OK(cudaMemsetAsync(scratch, 0, 64u << 20, stream));
        const uint32_t mask = (64u << 20) / 4 - 1;
        const double ms = best_ms([&] { red_kernel<<<blocks, 256, 0, stream>>>(scratch, mask, iters); count_launch(); }, stream);
        *result = (double)iters * blocks * 256.0 / (ms * 1e-3) / 1e9;
    } else {
        set_error("lsx_microbench: unknown kind %d", kind);
        return -1;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("lsx_microbench: %s", cudaGetErrorString(e));
        return -3;
    }
    return 0;
}

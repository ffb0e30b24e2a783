// async_copy.cuh — asynchronous global -> shared copy helpers (cp.async -> SASS LDGSTS / LDGDEPBAR / DEPBAR), sm_100a.
// Used by the tile renderers to stage batches of packed Gaussian records into shared memory while the previous
// batch is being blended.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lsx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- per-thread asynchronous 16-B copies (cp.async -> SASS LDGSTS): unlike the bulk copy, whose operands live in
// uniform registers (one elected lane per copy), every lane supplies its own source / destination.
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace lsx

// FP64 tensor-core (mma.sync.m8n8k4.f64) and cp.async helpers shared by the mat-vec / S-build kernels (vmc.cu) and the
// angle GEMM (psi.cu).  Fragment layout (PTX ISA, m8n8k4 .f64): A[row = lane>>2][k = lane&3], B[k = lane&3][col = lane>>2],
// C/D[row = lane>>2][col = 2*(lane&3) + {0,1}]  =>  each lane ends up with ONE complex number per 8x8 tile when a complex
// matrix is viewed as a real one with (re, im) adjacent.
#pragma once
#include "common.cuh"

namespace angpu {

#ifdef __CUDACC__
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double bit_sign(const uint64_t* cw, unsigned i) { return ((cw[i >> 6] >> (i & 63u)) & 1ull) ? 1.0 : -1.0; }

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int N_> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

#endif

} // namespace angpu

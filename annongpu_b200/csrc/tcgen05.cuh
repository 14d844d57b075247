// tcgen05 / TMA / mbarrier primitives shared by the tensor-core kernels (sbuild_tc.cu: S build; sv_tc.cu: factorised S.v), sm_100a.
// K-major fp32 (TF32) operand tiles of 128 rows x 16 elements (64-byte rows, SWIZZLE_64B), 128 x 128 fp32 accumulators in TMEM.
#pragma once
#include "runtime.hpp"
#include <cuda.h>

namespace angpu {
namespace tc {

constexpr int BLOCK_MN = 128;       // output tile
constexpr int BLOCK_K = 16;         // K elements per stage: 16 fp32 = 64 B rows (SWIZZLE_64B)
constexpr int UMMA_K = 8;           // tf32: 32 bytes per MMA
constexpr int TILE_BYTES = BLOCK_MN * BLOCK_K * 4;            // 8 KB

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while(!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// shared-memory matrix descriptor, K-major operand, 64-byte swizzle: 8-row groups are 512 B apart (SBO), LBO unused
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);            // start address
    d |= (uint64_t)0 << 16;                                   // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(512u >> 4) << 32;                         // stride byte offset
    d |= (uint64_t)1 << 46;                                   // descriptor version (sm_100)
    d |= (uint64_t)4 << 61;                                   // layout type: SWIZZLE_64B
    return d;
}
// instruction descriptor, kind::tf32, fp32 accumulate, K-major A and B, M = N = 128
__device__ __forceinline__ uint32_t umma_idesc_tf32(bool negate_a) {
    uint32_t d = 0;
    d |= 1u << 4;                            // D format: F32
    d |= 2u << 7;                            // A format: TF32
    d |= 2u << 10;                           // B format: TF32
    d |= (negate_a ? 1u : 0u) << 13;         // negate A
    d |= (uint32_t)(BLOCK_MN >> 3) << 17;    // N
    d |= (uint32_t)(BLOCK_MN >> 4) << 24;    // M
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
#endif

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline encode_fn get_encode() {
    static encode_fn fn = nullptr;
    if(!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        ANGPU_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        ANGPU_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<encode_fn>(p);
    }
    return fn;
}
// plane: `rows` rows of `kext` fp32 (row stride `kstride` elements, a multiple of 4); box = 16 x 128, 64-byte swizzle, zeros out of bounds
inline void make_map(CUtensorMap* m, float* plane, size_t rows, size_t kext, size_t kstride) {
    const cuuint64_t gdim[2] = {(cuuint64_t)kext, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)kstride * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)BLOCK_MN};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, plane, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if(r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
}

} // namespace tc
} // namespace angpu

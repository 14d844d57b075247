// S-matrix build on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
//   S_rc = sum_s conj(A_sr) A_sc,   A_sk = sqrt(w_s) (O_sk - <O_k>)        (TDVP.cu.template:216-302, centred form)
//        = [Re_r.Re_c + Im_r.Im_c] + i [Re_r.Im_c - Im_r.Re_c]            (dot products over the samples)
//
// The reference builds S with ns*P^2 global fp64 atomics (TDVP.cu.template:260-263).  Here the one genuinely dense
// contraction of the path runs as a TF32 tensor-core GEMM with fp32 accumulation in TMEM, made ~fp32-accurate by the
// standard 3xTF32 split  x = hi + lo (both TF32-exact):  a.b ~= hi.hi + hi.lo + lo.hi.   Centring and sqrt(w) scaling are
// applied in fp64 BEFORE the split, so no large mean term is cancelled in reduced precision.
// This is the opt-in FAST path (tolerance ~1e-5 relative to ||S||, BASELINE.json's fp32 tolerance); the default
// TDVP::build_S (vmc.cu: k_zherk) stays exact fp64.
//
// Data:   k_pack_planes  transposes O [ns][P] (complex fp64) into four K-major fp32 planes [P][Kpad]:
//         Re_hi, Re_lo, Im_hi, Im_lo (row = parameter, contiguous over samples, zero padded to Kpad).
// Kernel: k_sbuild_tf32  one CTA per upper-triangular 128x128 tile pair (tr <= tc), 192 threads:
//           warp 0      TMA producer   8 tiles / stage: {Re,Im} x {hi,lo} for the row block and the column block
//                                      (box 16 fp32 x 128 rows, 64B swizzle), 3 stages x 64 KB
//           warp 1      MMA issuer     12 x (BLOCK_K/8) tcgen05.mma.kind::tf32 per stage into two 128x128 fp32
//                                      accumulators in TMEM (Re S, Im S; Im uses the negate-A bit), tcgen05.commit -> mbarrier
//           warps 2-17  epilogue       every 128 samples: tcgen05.ld 32x32b of the finished accumulator set, added with
//                                      round-to-nearest into fp32 registers (the tensor core's own fp32 accumulation
//                                      truncates); at the end fp64 -> S[r][c] and its Hermitian mirror S[c][r]
#include "vmc.hpp"
#include "tcgen05.cuh"
#include <cmath>

namespace angpu {

namespace tc {

constexpr int STAGES = 3;
constexpr int TILES_PER_STAGE = 8;
constexpr int STAGE_BYTES = TILES_PER_STAGE * TILE_BYTES;     // 64 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
// barriers: full[STAGES] | empty[STAGES] | tmem_full[2] | tmem_empty[2] | tmem base slot
constexpr int TMEM_COLS = 512;      // 2 ping-pong sets x (Re, Im) 128-column fp32 accumulators
constexpr int CHUNK_KB = 8;         // k-blocks (x16 samples) accumulated in TMEM before the epilogue drains the set
constexpr int EPI_WARPS = 16;       // 4 lane quarters x 4 column groups of 32
constexpr int THREADS = 64 + 32 * EPI_WARPS;

struct Maps { CUtensorMap re_hi, re_lo, im_hi, im_lo; };

__global__ void __launch_bounds__(THREADS, 1)
k_sbuild_tf32(const __grid_constant__ Maps maps, unsigned P, unsigned num_kb, cplx* __restrict__ S) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned tile area, then the barriers
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    const uint32_t bars = tiles + STAGES * STAGE_BYTES;
    const uint32_t bar_full = bars, bar_empty = bars + 8u * STAGES, bar_tfull = bars + 16u * STAGES, bar_tempty = bar_tfull + 16u,
                   tmem_slot = bar_tempty + 16u;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

    // upper-triangular tile pair
    const unsigned nt = (P + BLOCK_MN - 1) / BLOCK_MN;
    unsigned t = blockIdx.x, tr = 0;
    while(t >= nt - tr) { t -= nt - tr; tr++; }
    const unsigned tc_ = tr + t;
    const int row0 = (int)(tr * BLOCK_MN), col0 = (int)(tc_ * BLOCK_MN);

    if(threadIdx.x == 0) {
        for(int s = 0; s < STAGES; s++) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
        for(int q = 0; q < 2; q++) { mbar_init(bar_tfull + 8u * q, 1); mbar_init(bar_tempty + 8u * q, EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if(warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if(warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if(lane == 0) {
            const CUtensorMap* m[4] = {&maps.re_hi, &maps.re_lo, &maps.im_hi, &maps.im_lo};
            for(unsigned kb = 0; kb < num_kb; kb++) {
                const unsigned s = kb % STAGES, ph = (kb / STAGES) & 1u;
                mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                mbar_arrive_expect_tx(bar_full + 8u * s, (uint32_t)STAGE_BYTES);
                const uint32_t base = tiles + s * STAGE_BYTES;
                const int k0 = (int)(kb * BLOCK_K);
                #pragma unroll
                for(int q = 0; q < 4; q++) {
                    tma_load_2d(base + (uint32_t)q * TILE_BYTES, m[q], k0, row0, bar_full + 8u * s);         // row block planes
                    tma_load_2d(base + (uint32_t)(4 + q) * TILE_BYTES, m[q], k0, col0, bar_full + 8u * s);   // column block planes
                }
            }
        }
    } else if(warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if(lane == 0) {
            const uint32_t id_pos = umma_idesc_tf32(false), id_neg = umma_idesc_tf32(true);
            for(unsigned kb = 0; kb < num_kb; kb++) {
                const unsigned s = kb % STAGES, ph = (kb / STAGES) & 1u;
                const unsigned chunk = kb / CHUNK_KB, set = chunk & 1u, in_chunk = kb % CHUNK_KB;
                const uint32_t acc_re = tmem_base + set * 2u * BLOCK_MN, acc_im = acc_re + (uint32_t)BLOCK_MN;
                if(in_chunk == 0) {
                    // the epilogue must have drained this accumulator set (first use of each set passes immediately)
                    mbar_wait(bar_tempty + 8u * set, ((chunk >> 1) & 1u) ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(bar_full + 8u * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t base = tiles + s * STAGE_BYTES;
                // tile order in a stage: 0 Re_r_hi, 1 Re_r_lo, 2 Im_r_hi, 3 Im_r_lo, 4 Re_c_hi, 5 Re_c_lo, 6 Im_c_hi, 7 Im_c_lo
                #pragma unroll
                for(int k = 0; k < BLOCK_K / UMMA_K; k++) {
                    const uint32_t koff = (uint32_t)(k * UMMA_K * 4);
                    uint64_t d[8];
                    #pragma unroll
                    for(int q = 0; q < 8; q++) d[q] = umma_desc_sw64(base + (uint32_t)q * TILE_BYTES + koff);
                    const uint32_t first = (in_chunk == 0 && k == 0) ? 0u : 1u;
                    // Re S += Re_r.Re_c + Im_r.Im_c      (hi.hi + hi.lo + lo.hi each)
                    umma_tf32(acc_re, d[0], d[4], id_pos, first);
                    umma_tf32(acc_re, d[0], d[5], id_pos, 1u);
                    umma_tf32(acc_re, d[1], d[4], id_pos, 1u);
                    umma_tf32(acc_re, d[2], d[6], id_pos, 1u);
                    umma_tf32(acc_re, d[2], d[7], id_pos, 1u);
                    umma_tf32(acc_re, d[3], d[6], id_pos, 1u);
                    // Im S += Re_r.Im_c - Im_r.Re_c
                    umma_tf32(acc_im, d[0], d[6], id_pos, first);
                    umma_tf32(acc_im, d[0], d[7], id_pos, 1u);
                    umma_tf32(acc_im, d[1], d[6], id_pos, 1u);
                    umma_tf32(acc_im, d[2], d[4], id_neg, 1u);
                    umma_tf32(acc_im, d[2], d[5], id_neg, 1u);
                    umma_tf32(acc_im, d[3], d[4], id_neg, 1u);
                }
                umma_commit(bar_empty + 8u * s);            // frees the stage when these MMAs have read it
                if(in_chunk == CHUNK_KB - 1 || kb == num_kb - 1) umma_commit(bar_tfull + 8u * set);   // chunk accumulated
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..17)
        // The tensor core accumulates in fp32 with truncation, so a long sum over the samples drifts (~n * 2^-24): every
        // CHUNK_KB k-blocks the accumulator set is drained and added, with round-to-nearest, into fp32 registers while the
        // MMA warp fills the other set.  Thread = one row (TMEM lane), 32 of the 128 columns, Re and Im.
        const unsigned quarter = warp & 3u;                 // TMEM lane quarter this warp may access
        const unsigned part = (warp - 2u) >> 2;             // 32-column group
        float accR[32], accI[32];
        #pragma unroll
        for(int j = 0; j < 32; j++) { accR[j] = 0.0f; accI[j] = 0.0f; }
        const unsigned num_chunks = (num_kb + CHUNK_KB - 1) / CHUNK_KB;
        for(unsigned chunk = 0; chunk < num_chunks; chunk++) {
            const unsigned set = chunk & 1u;
            mbar_wait(bar_tfull + 8u * set, (chunk >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tre = tmem_base + ((quarter * 32u) << 16) + set * 2u * BLOCK_MN + part * 32u;
            #pragma unroll
            for(int c16 = 0; c16 < 2; c16++) {
                uint32_t v[16];
                tmem_ld16(tre + (uint32_t)(c16 * 16), v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                #pragma unroll
                for(int j = 0; j < 16; j++) accR[c16 * 16 + j] += __uint_as_float(v[j]);
                tmem_ld16(tre + (uint32_t)BLOCK_MN + (uint32_t)(c16 * 16), v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                #pragma unroll
                for(int j = 0; j < 16; j++) accI[c16 * 16 + j] += __uint_as_float(v[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if(lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tempty + 8u * set) : "memory");
        }
        const unsigned r = (unsigned)row0 + quarter * 32u + lane;
        if(r < P) {
            #pragma unroll
            for(int j = 0; j < 32; j++) {
                const unsigned c = (unsigned)col0 + part * 32u + (unsigned)j;
                if(c < P) {
                    const cplx v((double)accR[j], (double)accI[j]);
                    if(tr != tc_) { S[(size_t)r * P + c] = v; S[(size_t)c * P + r] = conj(v); }
                    else if(c > r) { S[(size_t)r * P + c] = v; S[(size_t)c * P + r] = conj(v); }
                    else if(c == r) S[(size_t)r * P + c] = cplx(v.re, 0.0);      // diagonal of a Hermitian matrix: exactly real
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if(warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// O [ns][P] complex fp64 -> planes [P][Kpad] fp32: A = sqrt(w) (O - Obar), split hi/lo in TF32. 32x32 smem transpose.
__global__ void __launch_bounds__(256) k_pack_planes(const cplx* __restrict__ O, const double* __restrict__ w, const cplx* __restrict__ Obar,
                                                     double inv_W, size_t ns, unsigned P, size_t Kpad, float* __restrict__ re_hi, float* __restrict__ re_lo,
                                                     float* __restrict__ im_hi, float* __restrict__ im_lo) {
    __shared__ double tre[32][33], tim[32][33];
    const unsigned tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;      // 32 x 8
    const size_t s0 = (size_t)blockIdx.y * 32u;
    const unsigned k0 = blockIdx.x * 32u;
    for(unsigned j = ty; j < 32u; j += 8u) {
        const size_t s = s0 + j; const unsigned k = k0 + tx;
        double a = 0.0, b = 0.0;
        if(s < ns && k < P) {
            const double sw = sqrt(w[s]);
            const cplx o = O[s * P + k], m = inv_W * Obar[k];
            a = sw * (o.re - m.re); b = sw * (o.im - m.im);
        }
        tre[j][tx] = a; tim[j][tx] = b;
    }
    __syncthreads();
    for(unsigned j = ty; j < 32u; j += 8u) {
        const unsigned k = k0 + j; const size_t s = s0 + tx;
        if(k < P && s < Kpad) {
            const double a = tre[tx][j], b = tim[tx][j];
            const float ah = to_tf32((float)a), bh = to_tf32((float)b);
            const float al = to_tf32((float)(a - (double)ah)), bl = to_tf32((float)(b - (double)bh));
            const size_t idx = (size_t)k * Kpad + s;
            re_hi[idx] = ah; re_lo[idx] = al; im_hi[idx] = bh; im_lo[idx] = bl;
        }
    }
}

// S_rc += corr * conj(Obar_r) Obar_c  — restores the reference's un-normalised-weights convention
// S = sum w O*O - <O>*<O> when W = sum w != 1 (centring with <O>/W gives sum w O*O - <O>*<O>/W)
__global__ void k_rank1_add(cplx* __restrict__ S, const cplx* __restrict__ Obar, double corr, unsigned P) {
    const size_t total = (size_t)P * P;
    for(size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t r = e / P, c = e % P;
        S[e] += corr * (conj(Obar[r]) * Obar[c]);
    }
}

} // namespace tc

// S = (1/1) sum_s conj(A_sr) A_sc on the tensor cores.  Single-process only in this round (the centred form needs the
// global <O_k>, which eval() has already all-reduced; the partial S of each rank is all-reduced afterwards).
void TDVP::build_S_tensorcore() {
    set_reduce(sharded);
    ANGPU_REQUIRE(evaluated, "TDVP: call eval first");
    ensure_dense_O(last_psi);
    mark(5);
    const size_t ns = S.ns;
    const size_t Kpad = std::max<size_t>(tc::BLOCK_K, (ns + tc::BLOCK_K - 1) / tc::BLOCK_K * tc::BLOCK_K);
    Smat.resize((size_t)P * P);
    tc_planes.resize((size_t)4 * P * Kpad);
    float* re_hi = tc_planes.p; float* re_lo = re_hi + (size_t)P * Kpad; float* im_hi = re_lo + (size_t)P * Kpad; float* im_lo = im_hi + (size_t)P * Kpad;
    const double W = total_weight > 0.0 ? total_weight : 1.0;
    tc::k_pack_planes<<<dim3(ceil_div(P, 32), ceil_div(Kpad, 32)), 256, 0, stream()>>>(O.p, S.weight.p, Ok_dev(), 1.0 / W, ns, P, Kpad, re_hi, re_lo, im_hi, im_lo);
    ANGPU_CHECK_LAUNCH(); count_launch();
    tc::Maps maps;
    tc::make_map(&maps.re_hi, re_hi, P, Kpad, Kpad); tc::make_map(&maps.re_lo, re_lo, P, Kpad, Kpad);
    tc::make_map(&maps.im_hi, im_hi, P, Kpad, Kpad); tc::make_map(&maps.im_lo, im_lo, P, Kpad, Kpad);
    const unsigned nt = (P + tc::BLOCK_MN - 1) / tc::BLOCK_MN, tiles = nt * (nt + 1) / 2;
    ANGPU_CUDA(cudaFuncSetAttribute(tc::k_sbuild_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    tc::k_sbuild_tf32<<<tiles, tc::THREADS, tc::SMEM_BYTES, stream()>>>(maps, P, (unsigned)(Kpad / tc::BLOCK_K), Smat.p);
    ANGPU_CHECK_LAUNCH(); count_launch();
    allreduce_sum(reinterpret_cast<double*>(Smat.p), 2 * (size_t)P * P);
    if(std::fabs(1.0 / W - 1.0) > 1e-14) {
        const size_t total = (size_t)P * P;
        tc::k_rank1_add<<<(unsigned)std::min<size_t>((total + 255) / 256, (size_t)ctx().num_sms * 32), 256, 0, stream()>>>(Smat.p, Ok_dev(), 1.0 / W - 1.0, P);
        ANGPU_CHECK_LAUNCH(); count_launch();
    }
    have_S = true;
    mark(6);
    if(profile) { ANGPU_CUDA(cudaEventSynchronize(ev[6])); ANGPU_CUDA(cudaEventElapsedTime(&phase_ms[4], ev[5], ev[6])); }
}

} // namespace angpu

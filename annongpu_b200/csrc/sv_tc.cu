// Factorised S.v of a PsiRBM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// The matrix-free CG product is two GEMMs against the +-1 spin matrix sigma [ns][N] (vmc.cu: k_rowdot_dmma / k_colreduce_dmma on
// the FP64 tensor cores, exact):
//     U = sigma V          a_s  = sum_j T_sj U_sj                      V = the CG vector as [N][M]
//     Y = sigma^T Z        Z_sj = w_s a_s conj(T_sj)                   Y = sum_s w_s a_s conj(O_s)  as [N][M]
// sigma is EXACT in TF32; V and Z are split into TF32 hi + lo planes, so each product  sigma . (hi + lo)  loses nothing but the
// fp32 accumulation in TMEM (drained into round-to-nearest fp32 registers every 128 K-elements, like the S build).  The result
// is a product of ~1e-6 relative accuracy: the CG uses it for the search directions and refreshes the true residual with the exact
// FP64-tensor-core product every few iterations (TDVP::solve_cg), so the solution still meets its fp64 tolerance.
//
// One kernel, two shapes (operands K-major, 128-row x 16-element tiles, SWIZZLE_64B; one A tile + four B tiles per stage):
//   ROWDOT     A = sigma2 [ns][Kpad(N)]        B = V planes  [M][Kpad(N)]   K = sites      epilogue: a_part[col tile][s] = sum_j T_sj U_sj
//   COLREDUCE  A = sigma1 [N][Kpad(ns)]        B = Z planes  [M][Kpad(ns)]  K = samples    epilogue: partial Y[k-split][i][j]
//   warp 0: TMA producer; warp 1: MMA issuer (acc_re += A.B_re_hi + A.B_re_lo, acc_im likewise: 8 tcgen05.mma per 16 K-elements);
//   warps 2-17: epilogue (TMEM lane quarter x 32-column group).
#include "vmc.hpp"
#include "tcgen05.cuh"
#include <cmath>

namespace angpu {

namespace tc {

constexpr int SV_STAGES = 4;
constexpr int SV_TILES = 5;                                     // A | B_re_hi | B_re_lo | B_im_hi | B_im_lo
constexpr int SV_STAGE_BYTES = SV_TILES * TILE_BYTES;           // 40 KB
constexpr int SV_SMEM = SV_STAGES * SV_STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + 4 * BLOCK_MN * (int)sizeof(cplx) /*rowdot reduction*/;
constexpr int SV_TMEM_COLS = 512;                               // 2 ping-pong sets x (re, im) 128-column fp32 accumulators
constexpr int SV_CHUNK_KB = 8;                                  // k-blocks accumulated in TMEM between drains (128 K-elements)
constexpr int SV_EPI_WARPS = 16;
constexpr int SV_THREADS = 64 + 32 * SV_EPI_WARPS;

struct SvMaps { CUtensorMap a, re_hi, re_lo, im_hi, im_lo; };

// MODE 0: COLREDUCE  rows = sites (N), cols = hidden units (M), out = part_x[blockIdx.z][i * M + j]
// MODE 1: ROWDOT     rows = samples (ns), cols = hidden units (M), out = a_part[blockIdx.y][s] = sum_j T[s][j] (re + i im)
template<int MODE>
__global__ void __launch_bounds__(SV_THREADS, 1)
k_sv_tf32(const __grid_constant__ SvMaps maps, unsigned rows, unsigned cols, unsigned num_kb, unsigned kb_per_split, unsigned tiles_per_cta,
          const cplx* __restrict__ T, cplx* __restrict__ out, size_t out_stride) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    const uint32_t bars = tiles + SV_STAGES * SV_STAGE_BYTES;
    const uint32_t bar_full = bars, bar_empty = bars + 8u * SV_STAGES, bar_tfull = bars + 16u * SV_STAGES, bar_tempty = bar_tfull + 16u,
                   tmem_slot = bar_tempty + 16u;
    cplx* red = reinterpret_cast<cplx*>(smem_raw + ((tiles - raw) + SV_STAGES * SV_STAGE_BYTES + 256));     // [4][128] (ROWDOT)
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    // COLREDUCE: one column tile (blockIdx.y), the K range [kb0, kb1) of this k-split (blockIdx.z).
    // ROWDOT:    the whole K range, `tiles_per_cta` consecutive column tiles starting at blockIdx.y * tiles_per_cta: the accumulator
    //            sets ping-pong over (tile, drain chunk) units, so the epilogue of one tile overlaps the MMAs of the next.
    const int row0 = (int)(blockIdx.x * BLOCK_MN);
    const unsigned nct = (cols + BLOCK_MN - 1) / BLOCK_MN;
    const unsigned tile0 = (MODE == 1) ? blockIdx.y * tiles_per_cta : blockIdx.y;
    const unsigned ntile = (MODE == 1) ? (tile0 < nct ? min(tiles_per_cta, nct - tile0) : 0u) : 1u;
    const unsigned kb0 = blockIdx.z * kb_per_split, kb1 = min(num_kb, kb0 + kb_per_split), nkb = kb1 > kb0 ? kb1 - kb0 : 0u;
    const unsigned upt = (nkb + SV_CHUNK_KB - 1) / SV_CHUNK_KB;            // drain units per tile
    const unsigned total_it = ntile * nkb;

    if(threadIdx.x == 0) {
        for(int s = 0; s < SV_STAGES; s++) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
        for(int q = 0; q < 2; q++) { mbar_init(bar_tfull + 8u * q, 1); mbar_init(bar_tempty + 8u * q, SV_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if(warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)SV_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if(warp == 0) {
        if(lane == 0) {
            const CUtensorMap* mb[4] = {&maps.re_hi, &maps.re_lo, &maps.im_hi, &maps.im_lo};
            for(unsigned it = 0; it < total_it; it++) {
                const unsigned s = it % SV_STAGES, ph = (it / SV_STAGES) & 1u;
                mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                mbar_arrive_expect_tx(bar_full + 8u * s, (uint32_t)SV_STAGE_BYTES);
                const uint32_t base = tiles + s * SV_STAGE_BYTES;
                const int k0 = (int)((kb0 + it % nkb) * BLOCK_K), col0 = (int)((tile0 + it / nkb) * BLOCK_MN);
                tma_load_2d(base, &maps.a, k0, row0, bar_full + 8u * s);
                #pragma unroll
                for(int q = 0; q < 4; q++) tma_load_2d(base + (uint32_t)(1 + q) * TILE_BYTES, mb[q], k0, col0, bar_full + 8u * s);
            }
        }
    } else if(warp == 1) {
        if(lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(false);
            for(unsigned it = 0; it < total_it; it++) {
                const unsigned s = it % SV_STAGES, ph = (it / SV_STAGES) & 1u;
                const unsigned kb = it % nkb, unit = (it / nkb) * upt + kb / SV_CHUNK_KB, set = unit & 1u, in_chunk = kb % SV_CHUNK_KB;
                const uint32_t acc_re = tmem_base + set * 2u * BLOCK_MN, acc_im = acc_re + (uint32_t)BLOCK_MN;
                if(in_chunk == 0) {
                    mbar_wait(bar_tempty + 8u * set, ((unit >> 1) & 1u) ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(bar_full + 8u * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t base = tiles + s * SV_STAGE_BYTES;
                #pragma unroll
                for(int k = 0; k < BLOCK_K / UMMA_K; k++) {
                    const uint32_t koff = (uint32_t)(k * UMMA_K * 4);
                    uint64_t d[5];
                    #pragma unroll
                    for(int q = 0; q < 5; q++) d[q] = umma_desc_sw64(base + (uint32_t)q * TILE_BYTES + koff);
                    const uint32_t first = (in_chunk == 0 && k == 0) ? 0u : 1u;
                    umma_tf32(acc_re, d[0], d[1], idesc, first);
                    umma_tf32(acc_re, d[0], d[2], idesc, 1u);
                    umma_tf32(acc_im, d[0], d[3], idesc, first);
                    umma_tf32(acc_im, d[0], d[4], idesc, 1u);
                }
                umma_commit(bar_empty + 8u * s);
                if(in_chunk == SV_CHUNK_KB - 1 || kb == nkb - 1) umma_commit(bar_tfull + 8u * set);
            }
        }
    } else {
        const unsigned quarter = warp & 3u;                 // TMEM lane quarter this warp may access
        const unsigned part = (warp - 2u) >> 2;             // 32-column group
        const unsigned r = (unsigned)row0 + quarter * 32u + lane;
        cplx t(0.0, 0.0);                                   // ROWDOT: this row's sum over the column tiles of the CTA
        for(unsigned tile = 0; tile < ntile; tile++) {
            float accR[32], accI[32];
            #pragma unroll
            for(int j = 0; j < 32; j++) { accR[j] = 0.0f; accI[j] = 0.0f; }
            for(unsigned chunk = 0; chunk < upt; chunk++) {
                const unsigned unit = tile * upt + chunk, set = unit & 1u;
                mbar_wait(bar_tfull + 8u * set, (unit >> 1) & 1u);
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
            const unsigned col0 = (tile0 + tile) * BLOCK_MN;
            if(MODE == 0) {
                if(r < rows) {
                    cplx* o = out + (size_t)blockIdx.z * out_stride + (size_t)r * cols;
                    #pragma unroll
                    for(int j = 0; j < 32; j++) {
                        const unsigned c = col0 + part * 32u + (unsigned)j;
                        if(c < cols) o[c] = cplx((double)accR[j], (double)accI[j]);
                    }
                }
            } else if(r < rows) {
                const cplx* __restrict__ Tr = T + (size_t)r * cols;
                #pragma unroll
                for(int j = 0; j < 32; j++) {
                    const unsigned c = col0 + part * 32u + (unsigned)j;
                    if(c < cols) cfma(t, Tr[c], cplx((double)accR[j], (double)accI[j]));
                }
            }
        }
        if(MODE == 1) {
            red[part * BLOCK_MN + quarter * 32u + lane] = t;
            asm volatile("bar.sync 1, %0;" ::"r"(32 * SV_EPI_WARPS) : "memory");        // the 16 epilogue warps
            if(part == 0 && r < rows) {
                const unsigned q = quarter * 32u + lane;
                out[(size_t)blockIdx.y * out_stride + r] = (red[q] + red[BLOCK_MN + q]) + (red[2 * BLOCK_MN + q] + red[3 * BLOCK_MN + q]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if(warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)SV_TMEM_COLS) : "memory");
    }
}

// sigma as TF32 planes: sig1 [N][K1 = pad16(ns)] (row = site, COLREDUCE's A) and sig2 [ns][K2 = pad16(N)] (row = sample, ROWDOT's A)
__global__ void __launch_bounds__(256) k_pack_sigma(const uint64_t* __restrict__ conf, size_t ns, unsigned N, unsigned words, size_t K1, size_t K2,
                                                    float* __restrict__ sig1, float* __restrict__ sig2) {
    __shared__ float t[32][33];
    const unsigned tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;
    const size_t s0 = (size_t)blockIdx.y * 32u;
    const unsigned i0 = blockIdx.x * 32u;
    for(unsigned j = ty; j < 32u; j += 8u) {                            // (sample s0 + j, site i0 + tx)
        const size_t s = s0 + j; const unsigned i = i0 + tx;
        float v = 0.0f;
        if(s < ns && i < N) v = ((conf[s * words + (i >> 6)] >> (i & 63u)) & 1ull) ? 1.0f : -1.0f;
        t[j][tx] = v;
        if(s < ns && i < K2) sig2[s * K2 + i] = v;
    }
    __syncthreads();
    for(unsigned j = ty; j < 32u; j += 8u) {                            // (site i0 + j, sample s0 + tx)
        const unsigned i = i0 + j; const size_t s = s0 + tx;
        if(i < N && s < K1) sig1[(size_t)i * K1 + s] = t[tx][j];
    }
}
// V [N][M] complex fp64 -> planes [M][K2] fp32 (row = hidden unit, contiguous over sites), TF32 hi / lo of re and im
__global__ void __launch_bounds__(256) k_pack_v(const cplx* __restrict__ v, unsigned N, unsigned M, size_t K2, float* __restrict__ re_hi, float* __restrict__ re_lo,
                                                float* __restrict__ im_hi, float* __restrict__ im_lo) {
    __shared__ double tre[32][33], tim[32][33];
    const unsigned tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;
    const unsigned i0 = blockIdx.y * 32u, j0 = blockIdx.x * 32u;
    for(unsigned q = ty; q < 32u; q += 8u) {                            // (site i0 + q, unit j0 + tx)
        const unsigned i = i0 + q, j = j0 + tx;
        cplx z(0.0, 0.0);
        if(i < N && j < M) z = v[(size_t)i * M + j];
        tre[q][tx] = z.re; tim[q][tx] = z.im;
    }
    __syncthreads();
    for(unsigned q = ty; q < 32u; q += 8u) {                            // (unit j0 + q, site i0 + tx)
        const unsigned j = j0 + q, i = i0 + tx;
        if(j < M && i < K2) {
            const double a = tre[tx][q], b = tim[tx][q];
            const float ah = to_tf32((float)a), bh = to_tf32((float)b);
            const size_t idx = (size_t)j * K2 + i;
            re_hi[idx] = ah; re_lo[idx] = to_tf32((float)(a - (double)ah)); im_hi[idx] = bh; im_lo[idx] = to_tf32((float)(b - (double)bh));
        }
    }
}
// Z_sj = w_s (X_s - xbar) conj(T_sj) -> planes [M][K1] fp32 (row = hidden unit, contiguous over samples).  xbar = sum of `nbar`
// partials: with X_s = O_s . v and xbar = Obar . v the column sums ARE S v (sum_s w_s conj(O_s) (a_s - abar), for any total weight):
// the cancellation against conj(Obar) (Obar . v) happens here in fp64, before the TF32 split, instead of between two rounded sums.
__global__ void __launch_bounds__(256) k_pack_z(const cplx* __restrict__ T, const double* __restrict__ w, const cplx* __restrict__ X,
                                                const cplx* __restrict__ xbar_parts, unsigned nbar, size_t ns, unsigned M, size_t K1,
                                                float* __restrict__ re_hi, float* __restrict__ re_lo, float* __restrict__ im_hi, float* __restrict__ im_lo) {
    __shared__ double tre[32][33], tim[32][33];
    __shared__ cplx xbar_sh;
    const unsigned tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;
    const size_t s0 = (size_t)blockIdx.y * 32u;
    const unsigned j0 = blockIdx.x * 32u;
    if(threadIdx.x == 0) { cplx t(0.0, 0.0); for(unsigned q = 0; q < nbar; q++) t += xbar_parts[q]; xbar_sh = t; }
    __syncthreads();
    const cplx xbar = xbar_sh;
    for(unsigned q = ty; q < 32u; q += 8u) {                            // (sample s0 + q, unit j0 + tx)
        const size_t s = s0 + q; const unsigned j = j0 + tx;
        cplx z(0.0, 0.0);
        if(s < ns && j < M) z = (w[s] * (X[s] - xbar)) * conj(T[s * M + j]);
        tre[q][tx] = z.re; tim[q][tx] = z.im;
    }
    __syncthreads();
    for(unsigned q = ty; q < 32u; q += 8u) {                            // (unit j0 + q, sample s0 + tx)
        const unsigned j = j0 + q; const size_t s = s0 + tx;
        if(j < M && s < K1) {
            const double a = tre[tx][q], b = tim[tx][q];
            const float ah = to_tf32((float)a), bh = to_tf32((float)b);
            const size_t idx = (size_t)j * K1 + s;
            re_hi[idx] = ah; re_lo[idx] = to_tf32((float)(a - (double)ah)); im_hi[idx] = bh; im_lo[idx] = to_tf32((float)(b - (double)bh));
        }
    }
}
// row_a[s] = sum over the column tiles of a_part (fixed order)
__global__ void k_sum_a_parts(const cplx* __restrict__ a_part, unsigned parts, size_t ns, cplx* __restrict__ a) {
    for(size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += (size_t)gridDim.x * blockDim.x) {
        cplx t(0.0, 0.0);
        for(unsigned q = 0; q < parts; q++) t += a_part[(size_t)q * ns + s];
        a[s] = t;
    }
}

} // namespace tc

static size_t pad16(size_t n) { return std::max<size_t>(16, (n + 15) / 16 * 16); }

bool TDVP::tc_available() const { return factorised && S.ns >= 128 && rbm_M >= 32; }

// sigma planes of the samples of the last eval (once per eval: the configurations do not change during a solve)
void TDVP::tc_prepare() {
    if(tc_ready) return;
    (void)tc::get_encode();                       // throws when the driver lacks cuTensorMapEncodeTiled: solve_cg (auto) then stays exact
    const size_t ns = S.ns, K1 = pad16(ns), K2 = pad16(rbm_N);
    tc_sig.resize((size_t)rbm_N * K1 + ns * K2);
    float* sig1 = tc_sig.p; float* sig2 = sig1 + (size_t)rbm_N * K1;
    tc::k_pack_sigma<<<dim3(ceil_div(std::max<size_t>(rbm_N, K2), 32), ceil_div(std::max(ns, K1), 32)), 256, 0, stream()>>>(S.conf.p, ns, rbm_N, words, K1, K2, sig1, sig2);
    ANGPU_CHECK_LAUNCH(); count_launch();
    static bool attr = false;
    if(!attr) {
        ANGPU_CUDA(cudaFuncSetAttribute(tc::k_sv_tf32<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SV_SMEM));
        ANGPU_CUDA(cudaFuncSetAttribute(tc::k_sv_tf32<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SV_SMEM));
        attr = true;
    }
    tc_ready = true;
}
// row_a[s] = O_s . v with the sigma V product on the tcgen05 tensor cores
void TDVP::tc_rowdot(const cplx* v_dev) {
    tc_prepare();
    const size_t ns = S.ns, K1 = pad16(ns), K2 = pad16(rbm_N);
    const unsigned M = rbm_M, N = rbm_N;
    float* sig2 = tc_sig.p + (size_t)N * K1;
    tc_planes.resize(std::max(tc_planes.n, (size_t)4 * M * std::max(K1, K2)));
    float* p0 = tc_planes.p; float* p1 = p0 + (size_t)M * K2; float* p2 = p1 + (size_t)M * K2; float* p3 = p2 + (size_t)M * K2;
    tc::k_pack_v<<<dim3(ceil_div(M, 32), ceil_div(K2, 32)), 256, 0, stream()>>>(v_dev, N, M, K2, p0, p1, p2, p3);
    // tensor maps are cached per (buffers, shape): the buffers are grow-only, so a CG solve encodes them once
    static thread_local struct { const void* a = nullptr; const void* b = nullptr; size_t ns = 0, K = 0; unsigned M = 0; tc::SvMaps maps; } cache;
    if(cache.a != sig2 || cache.b != p0 || cache.ns != ns || cache.K != K2 || cache.M != M) {
        tc::make_map(&cache.maps.a, sig2, ns, K2, K2);
        tc::make_map(&cache.maps.re_hi, p0, M, K2, K2); tc::make_map(&cache.maps.re_lo, p1, M, K2, K2);
        tc::make_map(&cache.maps.im_hi, p2, M, K2, K2); tc::make_map(&cache.maps.im_lo, p3, M, K2, K2);
        cache.a = sig2; cache.b = p0; cache.ns = ns; cache.K = K2; cache.M = M;
    }
    const tc::SvMaps& maps = cache.maps;
    const unsigned nct = ceil_div(M, tc::BLOCK_MN), nst = ceil_div(ns, tc::BLOCK_MN), num_kb = (unsigned)(K2 / tc::BLOCK_K);
    // column tiles per CTA: as many as keep ~2 CTAs per SM (the tiles of a CTA pipeline through the two accumulator sets)
    unsigned groups = std::max(1u, std::min(nct, ((unsigned)ctx().num_sms * 2u + nst - 1u) / nst));
    const unsigned tpc = (nct + groups - 1u) / groups;
    groups = (nct + tpc - 1u) / tpc;
    const unsigned ncb = groups;
    tc_apart.resize((size_t)ncb * ns);
    row_a.resize(std::max<size_t>(1, ns));
    tc::k_sv_tf32<1><<<dim3(nst, groups, 1), tc::SV_THREADS, tc::SV_SMEM, stream()>>>(maps, (unsigned)ns, M, num_kb, num_kb, tpc, T.p, tc_apart.p, ns);
    tc::k_sum_a_parts<<<(unsigned)std::min<size_t>((ns + 255) / 256, (size_t)ctx().num_sms * 8), 256, 0, stream()>>>(tc_apart.p, ncb, ns, row_a.p);
    ANGPU_CHECK_LAUNCH(); count_launch(3);
}
// per-split partial sums of x_k = sum_s w_s (X_s - xbar) conj(O_sk), xbar = sum of the nbar partials, with the sigma^T Z product on the
// tcgen05 tensor cores (X = O v, xbar = Obar . v: the sums are S v with the mean already removed);
// returns the number of partials written to chunk_buf (layout of col_reduce_partials: [splits][P])
unsigned TDVP::tc_col_partials(const cplx* X, const cplx* xbar_parts, unsigned nbar, cplx** px_out) {
    tc_prepare();
    const size_t ns = S.ns, K1 = pad16(ns);
    const unsigned M = rbm_M, N = rbm_N;
    float* sig1 = tc_sig.p;
    tc_planes.resize(std::max(tc_planes.n, (size_t)4 * M * std::max(K1, pad16(N))));
    float* p0 = tc_planes.p; float* p1 = p0 + (size_t)M * K1; float* p2 = p1 + (size_t)M * K1; float* p3 = p2 + (size_t)M * K1;
    tc::k_pack_z<<<dim3(ceil_div(M, 32), ceil_div(K1, 32)), 256, 0, stream()>>>(T.p, S.weight.p, X, xbar_parts, nbar, ns, M, K1, p0, p1, p2, p3);
    static thread_local struct { const void* a = nullptr; const void* b = nullptr; size_t K = 0; unsigned N = 0, M = 0; tc::SvMaps maps; } cache;
    if(cache.a != sig1 || cache.b != p0 || cache.K != K1 || cache.N != N || cache.M != M) {
        tc::make_map(&cache.maps.a, sig1, N, K1, K1);
        tc::make_map(&cache.maps.re_hi, p0, M, K1, K1); tc::make_map(&cache.maps.re_lo, p1, M, K1, K1);
        tc::make_map(&cache.maps.im_hi, p2, M, K1, K1); tc::make_map(&cache.maps.im_lo, p3, M, K1, K1);
        cache.a = sig1; cache.b = p0; cache.K = K1; cache.N = N; cache.M = M;
    }
    const tc::SvMaps& maps = cache.maps;
    const unsigned nrt = ceil_div(N, tc::BLOCK_MN), nct = ceil_div(M, tc::BLOCK_MN), num_kb = (unsigned)(K1 / tc::BLOCK_K);
    // k-splits: enough CTAs for ~2 per SM, whole drain chunks (8 k-blocks = 128 samples) per split, <= 64 partials
    unsigned splits = std::max(1u, std::min(64u, ((unsigned)ctx().num_sms * 2u + nrt * nct - 1u) / (nrt * nct)));
    unsigned kb_per = (num_kb + splits - 1u) / splits;
    kb_per = (kb_per + tc::SV_CHUNK_KB - 1u) / tc::SV_CHUNK_KB * tc::SV_CHUNK_KB;
    splits = (num_kb + kb_per - 1u) / kb_per;
    chunk_buf.resize(std::max(chunk_buf.n, (size_t)2 * splits * P));
    cplx* px = chunk_buf.p + (size_t)splits * P;
    tc::k_sv_tf32<0><<<dim3(nrt, nct, splits), tc::SV_THREADS, tc::SV_SMEM, stream()>>>(maps, N, M, num_kb, kb_per, 1u, nullptr, px, (size_t)P);
    ANGPU_CHECK_LAUNCH(); count_launch(2);
    *px_out = px;
    return splits;
}

} // namespace angpu

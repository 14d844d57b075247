// Hand-written dense Hermitian positive-definite solve for the SR / TDVP system (the reference has no solver):
// blocked Cholesky A = U^dagger U on the ROW-MAJOR upper triangle + two triangular solves, all fp64.
//
//   for each block row k (CH_NB rows):
//     k_chol_diag        factor the CH_NB x CH_NB diagonal block in shared memory (packed upper triangle, 132 KB)
//     k_tri_inv          T_k = U_kk^{-1}, kept for the triangular solves
//     k_chol_panel_gemm  block row  U_k,>k = T_k^dagger A_k,>k  as a register-tiled product (no substitution chain)
//     k_zherk_dmma       trailing update  A_>k,>k -= U_k,>k^dagger U_k,>k  on the FP64 tensor cores (mma.sync.m8n8k4.f64): the block row
//                        just computed IS the operand, in place (zherk_dmma.cuh, the kernel of the exact S build in SUB mode);
//                        4/3 P^3 flops in total -- everything else is O(P^2 CH_NB)
//   U^dagger y = b   per block row: y_k = T_k^dagger b_k, then the axpy update of the columns to the right (rows of U contiguous)
//   U x = y          per block row: row dots over the columns to the right, then x_k = T_k t_k
// Row-major + upper triangle makes every access of the O(P^3) part contiguous; nothing is transposed or copied.
#include <cstdlib>
#include "vmc.hpp"
#include "zherk_dmma.cuh"

namespace angpu {

constexpr int CH_GROUP = 4;                                  // block rows per trailing update (ANGPU_CHOL_GROUP overrides)
constexpr int CH_NB = 128;                                  // block size: packed upper triangle of a block = 132 KB of shared memory
constexpr int CH_PACK = CH_NB * (CH_NB + 1) / 2;
__device__ __forceinline__ int ch_off(int r, int nb) { return r * nb - (r * (r - 1)) / 2 - r; }      // packed index of (r, c): ch_off(r) + c

// Cholesky of the nb x nb diagonal block at A (leading dimension lda), upper triangle, in shared memory (packed), blocked by 32:
//   (a) the 32 ROWS of a sub-block, over all columns to their right: right-looking elimination, one thread per (row, column mod 32);
//       the pivot row is used UNSCALED (a_rc -= conj(a_jr) a_jc / a_jj) and scaled once at the end, so a step needs ONE barrier;
//   (c) the rows below: rank-32 update, one thread per entry of every 32 x 32 tile.
// info: first non-positive pivot (1-based global index) or unchanged.
constexpr int CD_B = 32;
__global__ void __launch_bounds__(1024) k_chol_diag(cplx* __restrict__ A, size_t lda, int nb, int k0, int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char ch_smem[];
    cplx* U = reinterpret_cast<cplx*>(ch_smem);
    __shared__ double piv[CD_B];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    for(int r = ty; r < nb; r += 32)
        for(int c = r + tx; c < nb; c += 32) U[ch_off(r, nb) + c] = A[(size_t)r * lda + c];
    __syncthreads();
    for(int j0 = 0; j0 < nb; j0 += CD_B) {
        const int jb = min(CD_B, nb - j0);
        // (a)
        for(int s = 0; s < jb; s++) {
            const int j = j0 + s;
            double d = U[ch_off(j, nb) + j].re;
            if(!(d > 0.0)) { if(tid == 0 && *info == 0) *info = k0 + j + 1; d = 1.0; }
            if(tid == 0) piv[s] = d;
            if(ty > s && ty < jb) {
                const int r = j0 + ty;
                const cplx ur = U[ch_off(j, nb) + r];
                const double di = 1.0 / d;
                const double fr = ur.re * di, fi = ur.im * di;
                for(int c = j0 + tx; c < nb; c += 32) {
                    if(c >= r) {
                        const cplx uc = U[ch_off(j, nb) + c];
                        cplx v = U[ch_off(r, nb) + c];
                        v.re -= fr * uc.re + fi * uc.im; v.im -= fr * uc.im - fi * uc.re;
                        U[ch_off(r, nb) + c] = v;
                    }
                }
            }
            __syncthreads();
        }
        // scale the rows of the sub-block: U_jc = a_jc / sqrt(a_jj) for every column right of the pivot
        if(ty < jb) {
            const int r = j0 + ty;
            const double sd = sqrt(piv[ty]), di = 1.0 / sd;
            for(int c = j0 + tx; c < nb; c += 32) {
                if(c == r) U[ch_off(r, nb) + c] = cplx(sd, 0.0);
                else if(c > r) { cplx v = U[ch_off(r, nb) + c]; v.re *= di; v.im *= di; U[ch_off(r, nb) + c] = v; }
            }
        }
        __syncthreads();
        const int c1 = j0 + jb;                                  // first row / column below the sub-block
        if(c1 >= nb) break;
        // (c) A_rc -= sum_m conj(U_{j0+m, r}) U_{j0+m, c},  c1 <= r <= c < nb
        for(int R0 = c1; R0 < nb; R0 += 32) {
            const int r = R0 + ty;
            for(int C0 = R0; C0 < nb; C0 += 32) {
                const int c = C0 + tx;
                if(r < nb && c < nb && r <= c) {
                    cplx v = U[ch_off(r, nb) + c];
                    #pragma unroll 8
                    for(int m = 0; m < jb; m++) {
                        const cplx ur = U[ch_off(j0 + m, nb) + r], uc = U[ch_off(j0 + m, nb) + c];
                        v.re -= ur.re * uc.re + ur.im * uc.im; v.im -= ur.re * uc.im - ur.im * uc.re;
                    }
                    U[ch_off(r, nb) + c] = v;
                }
            }
        }
        __syncthreads();
    }
    for(int r = ty; r < nb; r += 32)
        for(int c = r + tx; c < nb; c += 32) A[(size_t)r * lda + c] = U[ch_off(r, nb) + c];
}

// T = U_kk^{-1} (upper triangular) of the factored nb x nb diagonal block, written as a full CH_NB x CH_NB row-major tile (zeros below
// the diagonal and beyond nb).  With the explicit inverse the block-row solve and the diagonal steps of the two triangular solves
// become matrix products -- full parallelism instead of 128 dependent substitution steps on the critical path of every block row.
// In place in shared memory (packed upper triangle), blocked by 32 like LAPACK's trtri:
//   1. the 32 x 32 diagonal sub-blocks, one warp each, column by column (trti2): T_jj = 1/U_jj, T_{<j, j} = -T_jj * (T_lead U_{<j, j});
//   2. block columns left to right:  T_{lead, b} = -(T_lead U_{lead, b}) T_bb  -- two small products on all 1024 threads, the
//      intermediate V = T_lead U_{lead, b} in a 48 KB scratch tile.
constexpr int TI_T = 1024;
constexpr size_t TI_SMEM = (size_t)CH_PACK * sizeof(cplx) + (size_t)(CH_NB - CD_B) * CD_B * sizeof(cplx);
__global__ void __launch_bounds__(TI_T) k_tri_inv(const cplx* __restrict__ Ukk, size_t lda, int nb, cplx* __restrict__ T) {
    extern __shared__ __align__(16) unsigned char ch_smem[];
    cplx* X = reinterpret_cast<cplx*>(ch_smem);                     // packed upper triangle: U on entry, U^{-1} at the end
    cplx* V = X + CH_PACK;                                          // [<= 96][32]
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    for(int r = ty; r < nb; r += 32)
        for(int c = r + tx; c < nb; c += 32) X[ch_off(r, nb) + c] = Ukk[(size_t)r * lda + c];
    __syncthreads();
    // 1. diagonal sub-blocks: warp w inverts block w; lane = row r of the sub-block
    if(ty * CD_B < nb) {
        const int j0 = ty * CD_B, jb = min(CD_B, nb - j0);
        for(int j = 0; j < jb; j++) {
            const double tjj = 1.0 / X[ch_off(j0 + j, nb) + j0 + j].re;
            cplx v(0.0, 0.0);
            if(tx < j) {                                            // (T_lead u)_r = sum_{r <= k < j} T_rk U_kj   (T_lead already inverted)
                cplx v1(0.0, 0.0);
                int k = tx;
                for(; k + 1 < j; k += 2) {
                    cfma(v, X[ch_off(j0 + tx, nb) + j0 + k], X[ch_off(j0 + k, nb) + j0 + j]);
                    cfma(v1, X[ch_off(j0 + tx, nb) + j0 + k + 1], X[ch_off(j0 + k + 1, nb) + j0 + j]);
                }
                if(k < j) cfma(v, X[ch_off(j0 + tx, nb) + j0 + k], X[ch_off(j0 + k, nb) + j0 + j]);
                v += v1;
            }
            __syncwarp();                                           // every lane has read column j before it is overwritten
            if(tx < j) X[ch_off(j0 + tx, nb) + j0 + j] = cplx(-tjj * v.re, -tjj * v.im);
            else if(tx == j) X[ch_off(j0 + j, nb) + j0 + j] = cplx(tjj, 0.0);
            __syncwarp();
        }
    }
    __syncthreads();
    // 2. off-diagonal block columns
    for(int j0 = CD_B; j0 < nb; j0 += CD_B) {
        const int jb = min(CD_B, nb - j0);
        // V[r][c] = sum_{r <= k < j0} T_rk U_{k, j0+c}
        for(int r = ty; r < j0; r += 32) {
            cplx v(0.0, 0.0), v1(0.0, 0.0);
            if(tx < jb) {
                int k = r;
                for(; k + 1 < j0; k += 2) {
                    cfma(v, X[ch_off(r, nb) + k], X[ch_off(k, nb) + j0 + tx]);
                    cfma(v1, X[ch_off(r, nb) + k + 1], X[ch_off(k + 1, nb) + j0 + tx]);
                }
                if(k < j0) cfma(v, X[ch_off(r, nb) + k], X[ch_off(k, nb) + j0 + tx]);
                v += v1;
            }
            V[r * CD_B + tx] = v;
        }
        __syncthreads();
        // T_{r, j0+c} = -sum_{k <= c} V[r][k] T_{j0+k, j0+c}
        for(int r = ty; r < j0; r += 32) {
            if(tx < jb) {
                cplx v(0.0, 0.0), v1(0.0, 0.0);
                int k = 0;
                for(; k + 1 <= tx; k += 2) {
                    cfma(v, V[r * CD_B + k], X[ch_off(j0 + k, nb) + j0 + tx]);
                    cfma(v1, V[r * CD_B + k + 1], X[ch_off(j0 + k + 1, nb) + j0 + tx]);
                }
                if(k <= tx) cfma(v, V[r * CD_B + k], X[ch_off(j0 + k, nb) + j0 + tx]);
                X[ch_off(r, nb) + j0 + tx] = cplx(-(v.re + v1.re), -(v.im + v1.im));
            }
        }
        __syncthreads();
    }
    for(int e = tid; e < CH_NB * CH_NB; e += TI_T) {
        const int r = e / CH_NB, c = e % CH_NB;
        T[e] = (r <= c && c < nb) ? X[ch_off(r, nb) + c] : cplx(0.0, 0.0);
    }
}

// Block row  U_k,>k = U_kk^{-dagger} A_k,>k = T^dagger A_k,>k  as a register-tiled product, in place:
//   X[r][c] = sum_{m <= r} conj(T[m][r]) A[m][c]
// Block = all CH_NB rows x PG_C columns, 256 threads as 16 row groups (8 rows) x 16 column groups (PG_C/16 columns); the m range
// streams through shared memory in chunks of PG_K rows of T and A.  Every input row of the block's columns is read before the
// first output is written (outputs stay in registers until the end), so the update is safe in place.
constexpr int PG_C = 64, PG_K = 16, PG_T = 256;
__global__ void __launch_bounds__(PG_T) k_chol_panel_gemm(const cplx* __restrict__ T, cplx* __restrict__ B, size_t lda, int nb, size_t cols) {
    __shared__ cplx Ts[PG_K][CH_NB];
    __shared__ cplx As[PG_K][PG_C];
    const int tid = threadIdx.x, rg = tid >> 4, cg = tid & 15;
    const size_t c0 = (size_t)blockIdx.x * PG_C;
    cplx acc[8][4];
    #pragma unroll
    for(int a = 0; a < 8; a++)
        #pragma unroll
        for(int b = 0; b < 4; b++) acc[a][b] = cplx(0.0, 0.0);
    const int rmax_warp = (((tid >> 5) * 2 + 1) * 8 + 7);               // largest row held by this warp (two row groups)
    for(int m0 = 0; m0 < nb; m0 += PG_K) {
        for(int e = tid; e < PG_K * CH_NB; e += PG_T) { const int kk = e / CH_NB, r = e % CH_NB; Ts[kk][r] = (m0 + kk < nb) ? T[(size_t)(m0 + kk) * CH_NB + r] : cplx(0.0, 0.0); }
        for(int e = tid; e < PG_K * PG_C; e += PG_T) {
            const int kk = e / PG_C, c = e % PG_C;
            As[kk][c] = (m0 + kk < nb && c0 + c < cols) ? B[(size_t)(m0 + kk) * lda + c0 + c] : cplx(0.0, 0.0);
        }
        __syncthreads();
        if(m0 <= rmax_warp) {                                            // T is upper triangular: rows r < m contribute nothing
            #pragma unroll 4
            for(int kk = 0; kk < PG_K; kk++) {
                cplx t[8], a[4];
                #pragma unroll
                for(int q = 0; q < 8; q++) t[q] = Ts[kk][rg * 8 + q];
                #pragma unroll
                for(int q = 0; q < 4; q++) a[q] = As[kk][cg + 16 * q];
                #pragma unroll
                for(int x = 0; x < 8; x++)
                    #pragma unroll
                    for(int y = 0; y < 4; y++) {                         // conj(t) * a
                        acc[x][y].re = fma(t[x].re, a[y].re, fma(t[x].im, a[y].im, acc[x][y].re));
                        acc[x][y].im = fma(t[x].re, a[y].im, fma(-t[x].im, a[y].re, acc[x][y].im));
                    }
            }
        }
        __syncthreads();
    }
    #pragma unroll
    for(int x = 0; x < 8; x++) {
        const int r = rg * 8 + x;
        if(r >= nb) continue;
        #pragma unroll
        for(int y = 0; y < 4; y++) { const size_t c = c0 + cg + 16 * y; if(c < cols) B[(size_t)r * lda + c] = acc[x][y]; }
    }
}

// ---- triangular solves with the factor (row-major upper U, leading dimension lda) and the inverted diagonal blocks T_k
// forward step of block k:  y_k = T_k^dagger b_k  (every block recomputes it: 128 x 128 / 2 MACs; block 0 stores it), then
// b[c] -= sum_m conj(U[m][c]) y_m for the columns right of the block: 64 columns per block, the rows dealt to 4 thread groups whose
// partial sums are added in a fixed order.  y is a separate vector (b_k is read by all blocks).
constexpr int TF_C = 64;
__global__ void __launch_bounds__(256) k_trsv_fwd(const cplx* __restrict__ T, const cplx* __restrict__ Urow, size_t lda, int nb, const cplx* __restrict__ bk,
                                                  cplx* __restrict__ yk, cplx* __restrict__ b, size_t cols) {
    __shared__ cplx bs[CH_NB], ys[CH_NB], part[4][TF_C];
    const int tid = threadIdx.x;
    if(tid < CH_NB) bs[tid] = (tid < nb) ? bk[tid] : cplx(0.0, 0.0);
    __syncthreads();
    if(tid < CH_NB) {
        cplx v(0.0, 0.0), v1(0.0, 0.0);
        if(tid < nb) {
            int m = 0;
            #pragma unroll 4
            for(; m + 1 <= tid; m += 2) {
                const cplx t = T[(size_t)m * CH_NB + tid], x = bs[m], t1 = T[(size_t)(m + 1) * CH_NB + tid], x1 = bs[m + 1];
                v.re += t.re * x.re + t.im * x.im; v.im += t.re * x.im - t.im * x.re;
                v1.re += t1.re * x1.re + t1.im * x1.im; v1.im += t1.re * x1.im - t1.im * x1.re;
            }
            if(m <= tid) { const cplx t = T[(size_t)m * CH_NB + tid], x = bs[m]; v.re += t.re * x.re + t.im * x.im; v.im += t.re * x.im - t.im * x.re; }
            v += v1;
            if(blockIdx.x == 0) yk[tid] = v;
        }
        ys[tid] = v;
    }
    __syncthreads();
    const int cx = tid & (TF_C - 1), g = tid / TF_C;
    const size_t c = (size_t)blockIdx.x * TF_C + cx;
    cplx v(0.0, 0.0);
    if(c < cols) {
        const int mlo = g * (CH_NB / 4), mhi = min(nb, mlo + CH_NB / 4);
        #pragma unroll 8
        for(int m = mlo; m < mhi; m++) {
            const cplx u = Urow[(size_t)m * lda + c], yv = ys[m];
            v.re += u.re * yv.re + u.im * yv.im; v.im += u.re * yv.im - u.im * yv.re;
        }
    }
    part[g][cx] = v;
    __syncthreads();
    if(g == 0 && c < cols) b[c] -= (part[0][cx] + part[1][cx]) + (part[2][cx] + part[3][cx]);
}
// backward step of block k, first half:  t_m = y_m - sum_c U[m][c] x_c over the columns right of the block (one block per row,
// fixed-order reduction)
__global__ void __launch_bounds__(256) k_trsv_bwd_dot(const cplx* __restrict__ Urow, size_t lda, const cplx* __restrict__ x, size_t cols,
                                                      const cplx* __restrict__ yk, cplx* __restrict__ tk) {
    __shared__ cplx part[8];
    const int m = blockIdx.x;
    cplx t(0.0, 0.0);
    for(size_t c = threadIdx.x; c < cols; c += 256) cfma(t, Urow[(size_t)m * lda + c], x[c]);
    t = warp_sum(t);
    if((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = t;
    __syncthreads();
    if(threadIdx.x == 0) {
        cplx s(0.0, 0.0);
        for(int q = 0; q < 8; q++) s += part[q];
        tk[m] = yk[m] - s;
    }
}
// second half:  x_k = T_k t  (one warp per row of T, lanes along the row)
__global__ void __launch_bounds__(1024) k_trsv_bwd_diag(const cplx* __restrict__ T, int nb, const cplx* __restrict__ tk, cplx* __restrict__ xk) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for(int r = warp; r < nb; r += 32) {
        cplx v(0.0, 0.0);
        for(int m = r + lane; m < nb; m += 32) cfma(v, T[(size_t)r * CH_NB + m], tk[m]);
        v = warp_sum(v);
        if(lane == 0) xk[r] = v;
    }
}

// A: P x P row-major Hermitian positive definite (upper triangle referenced, overwritten by U); b: right-hand side, overwritten by
// the solution.  info_dev[0] = 0 on success, else the 1-based index of the first non-positive pivot.
void cholesky_solve(cplx* A, cplx* b, unsigned P, int* info_dev, DevBuf<cplx>& work) {
    static bool attr = false;
    if(!attr) {
        ANGPU_CUDA(cudaFuncSetAttribute(k_chol_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CH_PACK * sizeof(cplx))));
        ANGPU_CUDA(cudaFuncSetAttribute(k_tri_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TI_SMEM));
        ANGPU_CUDA(cudaFuncSetAttribute(k_zherk_dmma<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZD_SMEM));
        attr = true;
    }
    ANGPU_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), stream()));
    const size_t lda = P;
    const unsigned nblk = (P + CH_NB - 1) / CH_NB;
    // workspace: the inverted diagonal blocks T_k (CH_NB x CH_NB each) + the intermediate vectors y and t
    work.resize((size_t)nblk * CH_NB * CH_NB + 2 * (size_t)nblk * CH_NB);
    cplx* Tall = work.p;
    cplx* yv = work.p + (size_t)nblk * CH_NB * CH_NB;
    cplx* tv = yv + (size_t)nblk * CH_NB;
    // Two-level blocking: diagonal blocks and block rows of CH_NB rows, trailing updates of a GROUP of G block rows at a time (1/G of the
    // passes over the trailing matrix, G times the work per tile of k_zherk_dmma).  Inside a group only the NEXT block row is brought
    // up to date after each block row -- the first two tile rows (128 complex rows) of the update, which k_zherk_dmma enumerates first.
    auto herk = [&](cudaStream_t st, unsigned r0, unsigned nrows, unsigned c0, unsigned tile_rows) {
        // A[c0:, c0:] -= U[r0:r0+nrows, c0:]^dagger U[r0:r0+nrows, c0:], restricted to the first `tile_rows` tile rows (0 = all)
        const unsigned cols = P - c0, nt = (cols + 63u) / 64u;
        unsigned tiles = nt * (nt + 1u) / 2u;
        if(tile_rows && tile_rows < nt) tiles = tile_rows * nt - tile_rows * (tile_rows - 1u) / 2u;
        k_zherk_dmma<4, true><<<dim3(tiles, 1), 512, ZD_SMEM, st>>>(reinterpret_cast<const double*>(A + (size_t)r0 * lda + c0), 2 * lda, nullptr,
                                                                      (size_t)nrows, cols, (size_t)nrows, A + (size_t)c0 * lda + c0, lda, 0);
        count_launch();
    };
    static const unsigned G = [] { const char* e = getenv("ANGPU_CHOL_GROUP"); const int g = e ? atoi(e) : CH_GROUP; return (unsigned)std::max(1, std::min(g, 16)); }();
    // Look-ahead: the trailing update of a group is issued in two parts -- first the rows of the NEXT group (the tile-row prefix of the
    // update), then everything below -- and the next group's chain of small kernels (diagonal factor, inverse, block row, in-group
    // updates: single-SM latency) runs on a second, high-priority stream while the second part streams on the main one.  The chain
    // touches only the next group's rows, the second part only the rows below them.  ANGPU_CHOL_LOOKAHEAD=0 serialises everything.
    static const bool lookahead = [] { const char* e = getenv("ANGPU_CHOL_LOOKAHEAD"); return !(e && atoi(e) == 0); }();
    static cudaStream_t side = nullptr;
    static cudaEvent_t ev_part_a = nullptr, ev_chain = nullptr;
    if(lookahead && !side) {
        int lo = 0, hi = 0;
        ANGPU_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        ANGPU_CUDA(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, hi));
        ANGPU_CUDA(cudaEventCreateWithFlags(&ev_part_a, cudaEventDisableTiming));
        ANGPU_CUDA(cudaEventCreateWithFlags(&ev_chain, cudaEventDisableTiming));
    }
    const cudaStream_t S0 = stream();
    bool chain_on_side = false;
    for(unsigned K0 = 0, kb = 0; K0 < P; K0 += G * CH_NB) {
        const cudaStream_t cs = chain_on_side ? side : S0;
        if(chain_on_side) ANGPU_CUDA(cudaStreamWaitEvent(side, ev_part_a, 0));      // this group's rows are up to date
        bool joined = !chain_on_side;
        for(unsigned sub = 0; sub < G; sub++, kb++) {
            const unsigned k0 = K0 + sub * CH_NB;
            if(k0 >= P) break;
            const int nb = (int)std::min<unsigned>(CH_NB, P - k0);
            const unsigned k1 = k0 + (unsigned)nb;
            cplx* Akk = A + (size_t)k0 * lda + k0;
            cplx* Tk = Tall + (size_t)kb * CH_NB * CH_NB;
            const size_t pack = (size_t)nb * (nb + 1) / 2 * sizeof(cplx);
            k_chol_diag<<<1, 1024, pack, cs>>>(Akk, lda, nb, (int)k0, info_dev);
            k_tri_inv<<<1, TI_T, TI_SMEM, cs>>>(Akk, lda, nb, Tk);
            count_launch(2);
            if(k1 < P) {
                k_chol_panel_gemm<<<ceil_div(P - k1, PG_C), PG_T, 0, cs>>>(Tk, Akk + nb, lda, nb, P - k1);
                count_launch();
                if(sub + 1 < G) herk(cs, K0, k1 - K0, k1, CH_NB / 64);               // the group's rows so far onto the next block row only
                else {
                    // the whole group onto everything below: the next group's rows first, then the rest
                    if(!joined) { ANGPU_CUDA(cudaEventRecord(ev_chain, side)); ANGPU_CUDA(cudaStreamWaitEvent(S0, ev_chain, 0)); joined = true; }
                    const unsigned c_mid = std::min(P, k1 + G * (unsigned)CH_NB);
                    if(lookahead && c_mid < P) {
                        herk(S0, K0, k1 - K0, k1, (c_mid - k1) / 64u);
                        ANGPU_CUDA(cudaEventRecord(ev_part_a, S0));
                        herk(S0, K0, k1 - K0, c_mid, 0);
                        chain_on_side = true;
                    } else {
                        herk(S0, K0, k1 - K0, k1, 0);
                        chain_on_side = false;
                    }
                }
            }
            ANGPU_CHECK_LAUNCH();
        }
        if(!joined) { ANGPU_CUDA(cudaEventRecord(ev_chain, side)); ANGPU_CUDA(cudaStreamWaitEvent(S0, ev_chain, 0)); }   // the factor ends inside this group
    }
    // U^dagger y = b
    for(unsigned k0 = 0, kb = 0; k0 < P; k0 += CH_NB, kb++) {
        const int nb = (int)std::min<unsigned>(CH_NB, P - k0);
        const unsigned k1 = k0 + (unsigned)nb;
        const cplx* Akk = A + (size_t)k0 * lda + k0;
        const size_t cols = P - k1;
        k_trsv_fwd<<<std::max(1u, ceil_div(cols, TF_C)), 256, 0, stream()>>>(Tall + (size_t)kb * CH_NB * CH_NB, Akk + nb, lda, nb, b + k0, yv + k0, b + k1, cols);
        count_launch();
    }
    // U x = y  (x overwrites b)
    for(unsigned kb = nblk; kb-- > 0;) {
        const unsigned k0 = kb * CH_NB;
        const int nb = (int)std::min<unsigned>(CH_NB, P - k0);
        const unsigned k1 = k0 + (unsigned)nb;
        const cplx* Akk = A + (size_t)k0 * lda + k0;
        const cplx* rhs = yv + k0;
        if(k1 < P) { k_trsv_bwd_dot<<<nb, 256, 0, stream()>>>(Akk + nb, lda, b + k1, P - k1, yv + k0, tv + k0); count_launch(); rhs = tv + k0; }
        k_trsv_bwd_diag<<<1, 1024, 0, stream()>>>(Tall + (size_t)kb * CH_NB * CH_NB, nb, rhs, b + k0);
        count_launch();
    }
    ANGPU_CHECK_LAUNCH();
}

} // namespace angpu

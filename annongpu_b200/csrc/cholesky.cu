// Hand-written dense Hermitian positive-definite solve for the SR / TDVP system (the reference has no solver):
// blocked Cholesky A = U^dagger U on the ROW-MAJOR upper triangle + two triangular solves, all fp64.
//
//   for each block row k (CH_NB rows):
//     k_chol_diag    factor the CH_NB x CH_NB diagonal block in shared memory (packed upper triangle, 132 KB)
//     k_chol_panel   block row  U_k,>k = U_kk^{-dagger} A_k,>k : one thread per column (coalesced along the row-major rows),
//                    16-row register sub-blocks
//     k_zherk_dmma   trailing update  A_>k,>k -= U_k,>k^dagger U_k,>k  on the FP64 tensor cores (mma.sync.m8n8k4.f64): the block row
//                    just computed IS the operand, in place (zherk_dmma.cuh, the kernel of the exact S build in SUB mode);
//                    4/3 P^3 flops in total -- everything else is O(P^2 CH_NB)
//   U^dagger y = b   column-sweep (axpy) form: rows of U are read contiguously
//   U x = y          dot form, again along rows
// Row-major + upper triangle makes every access of the O(P^3) part contiguous; nothing is transposed or copied.
#include "vmc.hpp"
#include "zherk_dmma.cuh"

namespace angpu {

constexpr int CH_NB = 128;                                  // block size: packed upper triangle of a block = 132 KB of shared memory
constexpr int CH_PACK = CH_NB * (CH_NB + 1) / 2;
__device__ __forceinline__ int ch_off(int r, int nb) { return r * nb - (r * (r - 1)) / 2 - r; }      // packed index of (r, c): ch_off(r) + c

// Unblocked right-looking Cholesky of the nb x nb diagonal block at A (leading dimension lda), upper triangle, in shared memory.
// info: first non-positive pivot (1-based global index) or unchanged.
__global__ void __launch_bounds__(1024) k_chol_diag(cplx* __restrict__ A, size_t lda, int nb, int k0, int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char ch_smem[];
    cplx* U = reinterpret_cast<cplx*>(ch_smem);
    __shared__ double dinv;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    for(int r = ty; r < nb; r += 32)
        for(int c = r + tx; c < nb; c += 32) U[ch_off(r, nb) + c] = A[(size_t)r * lda + c];
    __syncthreads();
    for(int j = 0; j < nb; j++) {
        if(tid == 0) {
            double d = U[ch_off(j, nb) + j].re;
            if(!(d > 0.0)) { if(*info == 0) *info = k0 + j + 1; d = 1.0; }
            d = sqrt(d);
            U[ch_off(j, nb) + j] = cplx(d, 0.0);
            dinv = 1.0 / d;
        }
        __syncthreads();
        const double di = dinv;
        for(int c = j + 1 + tid; c < nb; c += 1024) U[ch_off(j, nb) + c] = di * U[ch_off(j, nb) + c];
        __syncthreads();
        // A[r][c] -= conj(U[j][r]) U[j][c]  for j < r <= c
        for(int r = j + 1 + ty; r < nb; r += 32) {
            const cplx ur = conj(U[ch_off(j, nb) + r]);
            for(int c = r + tx; c < nb; c += 32) {
                cplx v = U[ch_off(r, nb) + c];
                const cplx uc = U[ch_off(j, nb) + c];
                v.re -= ur.re * uc.re - ur.im * uc.im; v.im -= ur.re * uc.im + ur.im * uc.re;
                U[ch_off(r, nb) + c] = v;
            }
        }
        __syncthreads();
    }
    for(int r = ty; r < nb; r += 32)
        for(int c = r + tx; c < nb; c += 32) A[(size_t)r * lda + c] = U[ch_off(r, nb) + c];
}

// Block row: for every column c of `cols` columns right of the diagonal block, solve U_kk^dagger x = a (forward substitution with the
// lower-triangular U_kk^dagger) in place.  Ukk: the factored diagonal block (row-major, lda); B: first row of the block row at the
// first column right of the block (same lda).  One thread per column; 16 rows of the column in registers at a time.
constexpr int CH_SUB = 16;
__global__ void __launch_bounds__(128) k_chol_panel(const cplx* __restrict__ Ukk, cplx* __restrict__ B, size_t lda, int nb, size_t cols) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(c >= cols) return;
    for(int r0 = 0; r0 < nb; r0 += CH_SUB) {
        cplx a[CH_SUB];
        #pragma unroll
        for(int q = 0; q < CH_SUB; q++) a[q] = (r0 + q < nb) ? B[(size_t)(r0 + q) * lda + c] : cplx(0.0, 0.0);
        for(int m = 0; m < r0; m++) {                          // rows solved before (read back; coalesced across the columns)
            const cplx x = B[(size_t)m * lda + c];
            const cplx* __restrict__ urow = Ukk + (size_t)m * lda + r0;
            #pragma unroll
            for(int q = 0; q < CH_SUB; q++) {
                const cplx u = (r0 + q < nb) ? ldg(&urow[q]) : cplx(0.0, 0.0);     // conj(U[m][r0+q]) x
                a[q].re -= u.re * x.re + u.im * x.im; a[q].im -= u.re * x.im - u.im * x.re;
            }
        }
        #pragma unroll
        for(int q = 0; q < CH_SUB; q++) {
            if(r0 + q < nb) {
                const cplx* __restrict__ urow = Ukk + (size_t)(r0 + q) * lda + r0;
                const double di = 1.0 / ldg(&urow[q]).re;
                a[q].re *= di; a[q].im *= di;
                #pragma unroll
                for(int p = q + 1; p < CH_SUB; p++) {
                    if(r0 + p < nb) {
                        const cplx u = ldg(&urow[p]);
                        a[p].re -= u.re * a[q].re + u.im * a[q].im; a[p].im -= u.re * a[q].im - u.im * a[q].re;
                    }
                }
                B[(size_t)(r0 + q) * lda + c] = a[q];
            }
        }
    }
}

// ---- triangular solves with the factor (row-major upper U, leading dimension lda), right-hand side b in place
// forward, diagonal block: U_kk^dagger y = b  (axpy form: y_m is final once all rows above are done)
__global__ void __launch_bounds__(CH_NB) k_trsv_fwd_diag(const cplx* __restrict__ Ukk, size_t lda, int nb, cplx* __restrict__ b) {
    __shared__ cplx ym;
    const int q = threadIdx.x;
    cplx v = (q < nb) ? b[q] : cplx(0.0, 0.0);
    for(int m = 0; m < nb; m++) {
        if(q == m) { const double di = 1.0 / Ukk[(size_t)m * lda + m].re; v.re *= di; v.im *= di; ym = v; }
        __syncthreads();
        if(q > m && q < nb) {
            const cplx u = Ukk[(size_t)m * lda + q], y = ym;
            v.re -= u.re * y.re + u.im * y.im; v.im -= u.re * y.im - u.im * y.re;
        }
        __syncthreads();
    }
    if(q < nb) b[q] = v;
}
// forward, the columns right of the block: b[c] -= sum_m conj(U[m][c]) y_m
__global__ void __launch_bounds__(256) k_trsv_fwd_update(const cplx* __restrict__ Urow, size_t lda, int nb, const cplx* __restrict__ y,
                                                         cplx* __restrict__ b, size_t cols) {
    __shared__ cplx ys[CH_NB];
    for(int m = threadIdx.x; m < nb; m += blockDim.x) ys[m] = y[m];
    __syncthreads();
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(c >= cols) return;
    cplx v = b[c];
    #pragma unroll 4
    for(int m = 0; m < nb; m++) {
        const cplx u = Urow[(size_t)m * lda + c], yv = ys[m];
        v.re -= u.re * yv.re + u.im * yv.im; v.im -= u.re * yv.im - u.im * yv.re;
    }
    b[c] = v;
}
// backward: t_m = sum_{c} U[m][c] x_c over the `cols` columns right of the block (one block per row, fixed-order reduction),
// subtracted from b_m
__global__ void __launch_bounds__(256) k_trsv_bwd_dot(const cplx* __restrict__ Urow, size_t lda, const cplx* __restrict__ x, size_t cols,
                                                      cplx* __restrict__ b) {
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
        b[m] -= s;
    }
}
// backward, diagonal block: U_kk x = b  (axpy form from the last row up; the column reads stay inside the L2-resident block)
__global__ void __launch_bounds__(CH_NB) k_trsv_bwd_diag(const cplx* __restrict__ Ukk, size_t lda, int nb, cplx* __restrict__ b) {
    __shared__ cplx xm;
    const int q = threadIdx.x;
    cplx v = (q < nb) ? b[q] : cplx(0.0, 0.0);
    for(int m = nb - 1; m >= 0; m--) {
        if(q == m) { const double di = 1.0 / Ukk[(size_t)m * lda + m].re; v.re *= di; v.im *= di; xm = v; }
        __syncthreads();
        if(q < m) { const cplx u = Ukk[(size_t)q * lda + m], x = xm; v.re -= u.re * x.re - u.im * x.im; v.im -= u.re * x.im + u.im * x.re; }
        __syncthreads();
    }
    if(q < nb) b[q] = v;
}

// A: P x P row-major Hermitian positive definite (upper triangle referenced, overwritten by U); b: right-hand side, overwritten by
// the solution.  info_dev[0] = 0 on success, else the 1-based index of the first non-positive pivot.
void cholesky_solve(cplx* A, cplx* b, unsigned P, int* info_dev) {
    static bool attr = false;
    if(!attr) {
        ANGPU_CUDA(cudaFuncSetAttribute(k_chol_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CH_PACK * sizeof(cplx))));
        ANGPU_CUDA(cudaFuncSetAttribute(k_zherk_dmma<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZD_SMEM));
        attr = true;
    }
    ANGPU_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), stream()));
    const size_t lda = P;
    for(unsigned k0 = 0; k0 < P; k0 += CH_NB) {
        const int nb = (int)std::min<unsigned>(CH_NB, P - k0);
        const unsigned k1 = k0 + (unsigned)nb;
        cplx* Akk = A + (size_t)k0 * lda + k0;
        k_chol_diag<<<1, 1024, (size_t)nb * (nb + 1) / 2 * sizeof(cplx), stream()>>>(Akk, lda, nb, (int)k0, info_dev);
        count_launch();
        if(k1 < P) {
            const size_t cols = P - k1;
            k_chol_panel<<<ceil_div(cols, 128), 128, 0, stream()>>>(Akk, Akk + nb, lda, nb, cols);
            const unsigned nt = (unsigned)((cols + 63) / 64), tiles = nt * (nt + 1) / 2;
            k_zherk_dmma<4, true><<<dim3(tiles, 1), 512, ZD_SMEM, stream()>>>(reinterpret_cast<const double*>(Akk + nb), 2 * lda, nullptr, (size_t)nb,
                                                                                 (unsigned)cols, (size_t)nb, A + (size_t)k1 * lda + k1, lda, 0);
            count_launch(2);
        }
        ANGPU_CHECK_LAUNCH();
    }
    // U^dagger y = b
    for(unsigned k0 = 0; k0 < P; k0 += CH_NB) {
        const int nb = (int)std::min<unsigned>(CH_NB, P - k0);
        const unsigned k1 = k0 + (unsigned)nb;
        const cplx* Akk = A + (size_t)k0 * lda + k0;
        k_trsv_fwd_diag<<<1, CH_NB, 0, stream()>>>(Akk, lda, nb, b + k0);
        count_launch();
        if(k1 < P) { k_trsv_fwd_update<<<ceil_div(P - k1, 256), 256, 0, stream()>>>(Akk + nb, lda, nb, b + k0, b + k1, P - k1); count_launch(); }
    }
    // U x = y
    for(unsigned kb = (P + CH_NB - 1) / CH_NB; kb-- > 0;) {
        const unsigned k0 = kb * CH_NB;
        const int nb = (int)std::min<unsigned>(CH_NB, P - k0);
        const unsigned k1 = k0 + (unsigned)nb;
        const cplx* Akk = A + (size_t)k0 * lda + k0;
        if(k1 < P) { k_trsv_bwd_dot<<<nb, 256, 0, stream()>>>(Akk + nb, lda, b + k1, P - k1, b + k0); count_launch(); }
        k_trsv_bwd_diag<<<1, CH_NB, 0, stream()>>>(Akk, lda, nb, b + k0);
        count_launch();
    }
    ANGPU_CHECK_LAUNCH();
}

} // namespace angpu

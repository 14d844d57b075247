// Hermitian rank-k update on the FP64 tensor cores, shared by the exact S build (vmc.cu) and the blocked Cholesky (cholesky.cu).
#pragma once
#include "dmma.cuh"

namespace angpu {

#ifdef __CUDACC__
// ---- exact fp64 S build on the FP64 tensor cores.  S'[k][k'] = sum_s w_s conj(O_sk) O_sk' is read off the REAL Gram matrix
// G = X^T diag(w) X of X = O viewed as [ns][2P] doubles ((re, im) adjacent):
//     Re S'[k][k'] = G[2k][2k'] + G[2k+1][2k'+1],   Im S'[k][k'] = G[2k][2k'+1] - G[2k+1][2k']
// (4 ns P^2 flops over the upper-triangular tiles, the same count as the complex Hermitian update).  Block = one 128 x 128
// real tile (64 x 64 complex), 8 warps as 4 x 2, warp tile 32 x 64 = 4 x 8 mma.m8n8k4.f64 tiles with accumulators in
// registers; both operand tiles stream through a cp.async double buffer of 32 samples; 12 LDS.64 feed 32 DMMAs per k-step.
constexpr int ZD_T = 128, ZD_KT = 32, ZD_PAD = 8, ZD_STRIDE = ZD_T + ZD_PAD;
constexpr int ZD_TILE_DOUBLES = ZD_KT * ZD_STRIDE;
constexpr size_t ZD_SMEM = (size_t)(4 * ZD_TILE_DOUBLES + 2 * ZD_KT) * sizeof(double);      // 2 stages x (A, B tile) + weights
// Generalised for the blocked Cholesky (cholesky.cu): X has a row stride `ldx` (doubles), the output a leading dimension `ldc`
// (complex), w == nullptr means unit weights, and SUB = true SUBTRACTS the upper-triangular entries from C instead of
// writing them with their mirror (the trailing update A_22 -= U_12^dagger U_12 with X = the block row U_12 in place).
template<int NWN, bool SUB>                               // warps along n: 2 (8 warps, warp tile 32 x 64) or 4 (16 warps, 32 x 32)
__global__ void __launch_bounds__(128 * NWN) k_zherk_dmma(const double* __restrict__ X, size_t ldx, const double* __restrict__ w, size_t ns, unsigned P,
                                                    size_t chunk, cplx* __restrict__ Sout, size_t ldc, size_t split_stride) {
    extern __shared__ __align__(16) double zd_smem[];
    const unsigned nt = (P + 63u) / 64u;
    unsigned t = blockIdx.x, tr = 0;
    while(t >= nt - tr) { t -= nt - tr; tr++; }
    const unsigned tc = tr + t;
    const size_t s0 = (size_t)blockIdx.y * chunk, s1 = min(ns, s0 + chunk);
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, row = lane >> 2, kq = lane & 3u;
    constexpr int NB = 8 / (NWN / 2);                          // n-tiles per warp: 8 or 4
    constexpr unsigned NT = 128u * NWN;                        // threads per block
    const unsigned m0 = (warp / NWN) * 32u, n0 = (warp % NWN) * (8u * NB);
    const unsigned ncol = 2u * P, ca = tr * ZD_T, cb = tc * ZD_T;
    const bool diag = (tr == tc);
    double* wsm = zd_smem + 4 * ZD_TILE_DOUBLES;
    auto issue = [&](size_t sb, unsigned buf) {
        double* Ta = zd_smem + (size_t)buf * 2 * ZD_TILE_DOUBLES;
        double* Tb = Ta + ZD_TILE_DOUBLES;
        for(unsigned e = threadIdx.x; e < ZD_KT * (ZD_T / 2); e += NT) {
            const unsigned kk = e / (ZD_T / 2), c = (e % (ZD_T / 2)) * 2u;
            const size_t sidx = sb + kk;
            const bool oka = sidx < s1 && ca + c < ncol, okb = sidx < s1 && cb + c < ncol;
            cp_async16_zfill(Ta + kk * ZD_STRIDE + c, oka ? (const void*)(X + sidx * ldx + ca + c) : (const void*)X, oka);
            if(!diag) cp_async16_zfill(Tb + kk * ZD_STRIDE + c, okb ? (const void*)(X + sidx * ldx + cb + c) : (const void*)X, okb);
        }
        if(threadIdx.x < ZD_KT) { const size_t sidx = sb + threadIdx.x; wsm[buf * ZD_KT + threadIdx.x] = sidx < s1 ? (w ? w[sidx] : 1.0) : 0.0; }
        cp_async_commit();
    };
    double acc[4][NB][2];
    #pragma unroll
    for(int a = 0; a < 4; a++)
        #pragma unroll
        for(int b = 0; b < NB; b++) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
    unsigned buf = 0;
    if(s0 < s1) issue(s0, 0);
    if(SUB) {
        // C - G: the accumulators of the even-row lanes start at -C (their 2 x 2 real block folds to the complex entry below), so the
        // read of C overlaps the first operand stage and the epilogue is store-only
        #pragma unroll
        for(int a = 0; a < 4; a++)
            #pragma unroll
            for(int b = 0; b < NB; b++) {
                const unsigned k = tr * 64u + (m0 + 8u * a + row) / 2u, kp = tc * 64u + (n0 + 8u * b) / 2u + kq;
                if((row & 1u) == 0u && k < P && kp < P && (!diag || kp >= k)) {
                    const cplx c = Sout[(size_t)k * ldc + kp];
                    acc[a][b][0] = -c.re; acc[a][b][1] = -c.im;
                }
            }
    }
    for(size_t sb = s0; sb < s1; sb += ZD_KT, buf ^= 1u) {
        if(sb + ZD_KT < s1) { issue(sb + ZD_KT, buf ^ 1u); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();
        const double* Ta = zd_smem + (size_t)buf * 2 * ZD_TILE_DOUBLES;
        const double* Tb = diag ? Ta : Ta + ZD_TILE_DOUBLES;
        const double* wk = wsm + buf * ZD_KT;
        #pragma unroll 2
        for(unsigned k4 = 0; k4 < ZD_KT / 4; k4++) {
            const unsigned kk = k4 * 4u + kq;
            const double wv = wk[kk];
            const double* ra = Ta + kk * ZD_STRIDE + m0 + row;
            const double* rb = Tb + kk * ZD_STRIDE + n0 + row;
            double af[4], bf[NB];
            #pragma unroll
            for(int a = 0; a < 4; a++) af[a] = wv * ra[a * 8];
            #pragma unroll
            for(int b = 0; b < NB; b++) bf[b] = rb[b * 8];
            #pragma unroll
            for(int a = 0; a < 4; a++)
                #pragma unroll
                for(int b = 0; b < NB; b++) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
        __syncthreads();                                   // the buffer is refilled by the next iteration's issue
    }
    // epilogue: lane (row r, kq) holds G[R][C], G[R][C+1] with R = m0 + 8a + r, C = n0 + 8b + 2 kq; the odd row R+1 lives in
    // lane ^ 4.  Even-row lanes combine the 2 x 2 real block into one complex entry and write it (and its mirror).
    cplx* Sp = Sout + (size_t)blockIdx.y * split_stride;
    #pragma unroll
    for(int a = 0; a < 4; a++)
        #pragma unroll
        for(int b = 0; b < NB; b++) {
            const double p0 = __shfl_xor_sync(FULL, acc[a][b][0], 4), p1 = __shfl_xor_sync(FULL, acc[a][b][1], 4);
            if((row & 1u) == 0u) {
                const unsigned k = tr * 64u + (m0 + 8u * a + row) / 2u, kp = tc * 64u + (n0 + 8u * b) / 2u + kq;
                if(k < P && kp < P && (!diag || kp >= k)) {        // diagonal tiles: upper part + mirror, exactly Hermitian
                    cplx v(acc[a][b][0] + p1, acc[a][b][1] - p0);
                    if(k == kp) v.im = 0.0;
                    if(SUB) {
                        Sp[(size_t)k * ldc + kp] = cplx(-v.re, -v.im);          // v = G - C
                    } else {
                        Sp[(size_t)k * ldc + kp] = v;
                        if(k != kp) Sp[(size_t)kp * ldc + k] = conj(v);
                    }
                }
            }
        }
}

#endif

} // namespace angpu

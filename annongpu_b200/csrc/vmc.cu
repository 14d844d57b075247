// Ensembles, reductions over samples, ExpectationValue / TDVP, S.v, CG and dense solve.
#include "vmc.hpp"
#include "pauli_basis.cuh"
#include "dmma.cuh"
#include "zherk_dmma.cuh"
#include <string>
#include <vector>
#include <algorithm>
#include <cmath>

namespace angpu {

// ============================================================================================ small kernels

__global__ void k_fill(double* p, double v, size_t n) {
    for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
// Spins::enumerate (include/basis/Spins.h:65-77, 291-296): basis index == bitmask in word 0
__global__ void k_enumerate(uint64_t* conf, size_t begin, size_t n, unsigned words) {
    for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        conf[i * words] = (uint64_t)(begin + i);
        for(unsigned w = 1; w < words; w++) conf[i * words + w] = 0ull;
    }
}

constexpr int RED_T = 1024;
// deterministic block reduction of NV doubles per thread (fixed tree order)
template<int NV>
__device__ void block_reduce(double (&v)[NV], double* out) {
    __shared__ double sm[NV][RED_T / 32];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    #pragma unroll
    for(int i = 0; i < NV; i++) { v[i] = warp_sum(v[i]); if(lane == 0) sm[i][warp] = v[i]; }
    __syncthreads();
    if(warp == 0) {
        #pragma unroll
        for(int i = 0; i < NV; i++) {
            double x = (lane < blockDim.x / 32) ? sm[i][lane] : 0.0;
            x = warp_sum(x);
            if(lane == 0) out[i] = x;
        }
    }
    __syncthreads();
}

// out4 = {Re sum w E, Im sum w E, sum w |E|^2, sum w};  lp2 (optional) = sum w log psi
__global__ void __launch_bounds__(RED_T) k_scalar_sums(const double* __restrict__ w, const cplx* __restrict__ eloc,
                                                       const cplx* __restrict__ log_psi, size_t ns, double* out4, double* lp2) {
    double v[6] = {0, 0, 0, 0, 0, 0};
    for(size_t s = threadIdx.x; s < ns; s += RED_T) {
        const double ws = w[s];
        if(eloc) { const cplx e = eloc[s]; v[0] += ws * e.re; v[1] += ws * e.im; v[2] += ws * abs2(e); }
        v[3] += ws;
        if(log_psi) { const cplx l = log_psi[s]; v[4] += ws * l.re; v[5] += ws * l.im; }
    }
    double r[6];
    block_reduce<6>(v, r);
    if(threadIdx.x == 0) {
        if(out4) { out4[0] = r[0]; out4[1] = r[1]; out4[2] = r[2]; out4[3] = r[3]; }
        if(lp2) { lp2[0] = r[4]; lp2[1] = r[5]; }
    }
}

// out = sum_k op(a_k) b_k, op = conj if CONJ_A  (single block, deterministic)
template<bool CONJ_A>
__global__ void __launch_bounds__(RED_T) k_dot(const cplx* __restrict__ a, const cplx* __restrict__ b, size_t n, cplx* out) {
    double v[2] = {0, 0};
    for(size_t k = threadIdx.x; k < n; k += RED_T) {
        cplx x = a[k]; if(CONJ_A) x = conj(x);
        const cplx p = x * b[k];
        v[0] += p.re; v[1] += p.im;
    }
    double r[2];
    block_reduce<2>(v, r);
    if(threadIdx.x == 0) *out = cplx(r[0], r[1]);
}

// ============================================================================================ column reductions
// mean_k = sum_s w_s O_sk ;  x_k = sum_s w_s X_s conj(O_sk)   over the samples of one chunk (blockIdx.y)

__global__ void __launch_bounds__(128) k_col_reduce_dense(const cplx* __restrict__ O, const double* __restrict__ w,
        const cplx* __restrict__ X, size_t ns, unsigned P, size_t chunk, cplx* __restrict__ part_mean, cplx* __restrict__ part_x) {
    const unsigned k = blockIdx.x * 128u + threadIdx.x;
    const size_t s0 = (size_t)blockIdx.y * chunk, s1 = min(ns, s0 + chunk);
    if(k >= P) return;
    cplx m(0.0, 0.0), x(0.0, 0.0);
    for(size_t s = s0; s < s1; s++) {
        const cplx o = O[s * P + k];
        const double ws = w[s];
        m.re = fma(ws, o.re, m.re); m.im = fma(ws, o.im, m.im);
        const cplx wx = ws * X[s];
        cfma(x, wx, conj(o));
    }
    if(part_mean) part_mean[(size_t)blockIdx.y * P + k] = m;
    part_x[(size_t)blockIdx.y * P + k] = x;
}

// PsiRBM factorised rows O_s,(i,j) = sigma_si T_sj.  Thread = column j with RBM_IT sites in registers; the signs of a
// 32-sample tile are staged in shared memory as +-1.0 doubles (warp-broadcast LDS.128), so the inner loop is
// 1 LDG (T) + RBM_IT/2 LDS + 2*RBM_IT DFMA per sample (4*RBM_IT with the mean).  Samples are split in chunks
// (blockIdx.z) whose partial sums are added in fixed order afterwards (deterministic, no atomics).
constexpr int RBM_IT = 16, RBM_TS = 32;
template<bool WANT_MEAN>
__global__ void __launch_bounds__(128) k_col_reduce_rbm(const uint64_t* __restrict__ conf, const cplx* __restrict__ T,
        const double* __restrict__ w, const cplx* __restrict__ X, size_t ns, unsigned N, unsigned M, unsigned words, size_t chunk,
        cplx* __restrict__ part_mean, cplx* __restrict__ part_x) {
    __shared__ __align__(16) double sgn[RBM_TS][RBM_IT];
    __shared__ cplx swx[RBM_TS];
    __shared__ double sw[RBM_TS];
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;       // blockDim.x = 32 .. 128 (narrow models: no idle threads)
    const unsigned i0 = blockIdx.y * RBM_IT;
    const size_t s0 = (size_t)blockIdx.z * chunk, s1 = min(ns, s0 + chunk);
    const bool jok = j < M;
    cplx m[WANT_MEAN ? RBM_IT : 1], x[RBM_IT];
    #pragma unroll
    for(int ii = 0; ii < RBM_IT; ii++) { x[ii] = cplx(0.0, 0.0); if(WANT_MEAN) m[ii] = cplx(0.0, 0.0); }
    const unsigned word = i0 >> 6, shift = i0 & 63u;     // RBM_IT divides 64: the sites of a tile share one word
    for(size_t sb = s0; sb < s1; sb += RBM_TS) {
        __syncthreads();
        for(unsigned e = threadIdx.x; e < RBM_TS * RBM_IT; e += blockDim.x) {
            const unsigned st = e / RBM_IT, ii = e % RBM_IT;
            const size_t s = sb + st;
            double sg = 0.0;
            if(s < s1) sg = ((conf[s * words + word] >> (shift + ii)) & 1ull) ? 1.0 : -1.0;
            sgn[st][ii] = sg;
        }
        if(threadIdx.x < RBM_TS) {
            const size_t s = sb + threadIdx.x;
            const double ws = (s < s1) ? w[s] : 0.0;
            sw[threadIdx.x] = ws;
            swx[threadIdx.x] = (s < s1) ? ws * X[s] : cplx(0.0, 0.0);
        }
        __syncthreads();
        const unsigned nst = (unsigned)min((size_t)RBM_TS, s1 - sb);
        if(jok) {
            for(unsigned st = 0; st < nst; st++) {
                const cplx t = T[(sb + st) * M + j];
                const cplx b = swx[st] * conj(t);
                cplx a(0.0, 0.0);
                if(WANT_MEAN) a = sw[st] * t;
                #pragma unroll
                for(int ii = 0; ii < RBM_IT; ii += 2) {
                    const double2 sg = *reinterpret_cast<const double2*>(&sgn[st][ii]);
                    x[ii].re = fma(sg.x, b.re, x[ii].re); x[ii].im = fma(sg.x, b.im, x[ii].im);
                    x[ii + 1].re = fma(sg.y, b.re, x[ii + 1].re); x[ii + 1].im = fma(sg.y, b.im, x[ii + 1].im);
                    if(WANT_MEAN) {
                        m[ii].re = fma(sg.x, a.re, m[ii].re); m[ii].im = fma(sg.x, a.im, m[ii].im);
                        m[ii + 1].re = fma(sg.y, a.re, m[ii + 1].re); m[ii + 1].im = fma(sg.y, a.im, m[ii + 1].im);
                    }
                }
            }
        }
    }
    if(!jok) return;
    const size_t P = (size_t)N * M;
    #pragma unroll
    for(int ii = 0; ii < RBM_IT; ii++) {
        const unsigned i = i0 + ii;
        if(i < N) {
            if(WANT_MEAN) part_mean[(size_t)blockIdx.z * P + (size_t)i * M + j] = m[ii];
            part_x[(size_t)blockIdx.z * P + (size_t)i * M + j] = x[ii];
        }
    }
}

// out[k] = sum_ch part[ch][k]  (fixed order)
__global__ void k_sum_chunks(const cplx* __restrict__ part, unsigned chunks, size_t P, cplx* __restrict__ out, bool conj_out = false) {
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < P; k += (size_t)gridDim.x * blockDim.x) {
        cplx a(0.0, 0.0);
        for(unsigned c = 0; c < chunks; c++) a += part[(size_t)c * P + k];
        out[k] = conj_out ? conj(a) : a;
    }
}
// The same sums for MANY chunks of a SHORT vector (C1: 512 chunks of P = 512): 32 columns x 32 chunk groups per block; group
// g adds the chunks g, g + 32, ... and the 32 group sums are added in fixed order.  Up to two vectors per launch.
__global__ void __launch_bounds__(1024) k_sum_chunks_wide(const cplx* __restrict__ part_a, const cplx* __restrict__ part_b, unsigned chunks, size_t P,
                                                          cplx* __restrict__ out_a, cplx* __restrict__ out_b, bool conj_b) {
    __shared__ cplx sm[32][33];
    const unsigned col = threadIdx.x & 31u, g = threadIdx.x >> 5;
    const size_t k = (size_t)blockIdx.x * 32u + col;
    const cplx* __restrict__ part = blockIdx.y ? part_b : part_a;
    cplx a(0.0, 0.0);
    if(k < P) for(unsigned c = g; c < chunks; c += 32u) a += part[(size_t)c * P + k];
    sm[g][col] = a;
    __syncthreads();
    if(g == 0 && k < P) {
        cplx t(0.0, 0.0);
        #pragma unroll 8
        for(int q = 0; q < 32; q++) t += sm[q][col];
        if(blockIdx.y) out_b[k] = conj_b ? conj(t) : t; else out_a[k] = t;
    }
}
// out_a (and out_b) = chunk sums, picking the launch shape by the number of chunks
static void sum_chunks(const cplx* part_a, cplx* out_a, const cplx* part_b, cplx* out_b, unsigned chunks, size_t P, bool conj_b = false);

__global__ void k_fill_cplx(cplx* p, cplx v, size_t n) {
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) p[k] = v;
}

// a_s = O_s . v   (dense rows; one block per sample)
__global__ void __launch_bounds__(256) k_rowdot_dense(const cplx* __restrict__ O, const cplx* __restrict__ v, unsigned P, cplx* __restrict__ a) {
    const size_t s = blockIdx.x;
    cplx acc(0.0, 0.0);
    for(unsigned k = threadIdx.x; k < P; k += 256u) cfma(acc, O[s * P + k], v[k]);
    __shared__ cplx sm[8];
    acc = warp_sum(acc);
    if((threadIdx.x & 31u) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if(threadIdx.x == 0) { cplx r(0.0, 0.0); for(int i = 0; i < 8; i++) r += sm[i]; a[s] = r; }
}

// out = (S + shift) v with the dense, already centred and rank-reduced S (row-major, P x P): one block per row, 16-byte
// coalesced loads; S is streamed once per product (P^2 x 16 B), v stays in L2.  HBM-bound.
__global__ void __launch_bounds__(256) k_smat_vec(const cplx* __restrict__ Sm, const cplx* __restrict__ v, const double* __restrict__ diag,
                                                  double shift_abs, double shift_rel, unsigned P, cplx* __restrict__ out) {
    const size_t i = blockIdx.x;
    const cplx* row = Sm + i * P;
    cplx a0(0.0, 0.0), a1(0.0, 0.0);
    unsigned k = threadIdx.x;
    for(; k + 256u < P; k += 512u) { cfma(a0, ldg(row + k), v[k]); cfma(a1, ldg(row + k + 256u), v[k + 256u]); }
    if(k < P) cfma(a0, ldg(row + k), v[k]);
    __shared__ cplx sm[8];
    cplx acc = warp_sum(a0 + a1);
    if((threadIdx.x & 31u) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if(threadIdx.x == 0) {
        cplx r(0.0, 0.0);
        for(int w = 0; w < 8; w++) r += sm[w];
        if(diag) r += (shift_abs + shift_rel * diag[i]) * v[i];
        out[i] = r;
    }
}

// a_s = sum_j T_sj (sum_i sigma_si v_ij)   (PsiRBM factorised rows).  Block = 32 samples x all j; thread = 8 samples x 2
// columns (j, j+64) per 128-column block; signs staged in shared memory as +-1.0 doubles [site][sample] so that the 8
// samples of a thread are 4 broadcast LDS.128; v is read once per block (L1/L2): ns/32 re-reads instead of ns/8.
constexpr int RBM_ST = 32, RBM_SI = 64;
__global__ void __launch_bounds__(256) k_rowdot_rbm(const uint64_t* __restrict__ conf, const cplx* __restrict__ T, const cplx* __restrict__ v,
        size_t ns, unsigned N, unsigned M, unsigned words, cplx* __restrict__ a) {
    __shared__ __align__(16) double sgn[RBM_SI][RBM_ST];
    __shared__ cplx red[RBM_ST][2];
    const size_t sb = (size_t)blockIdx.x * RBM_ST;
    const unsigned sg8 = (threadIdx.x >> 6) * 8u, jt = threadIdx.x & 63u;
    cplx tot[8];
    #pragma unroll
    for(int q = 0; q < 8; q++) tot[q] = cplx(0.0, 0.0);
    for(unsigned jb = 0; jb < M; jb += 128u) {
        const unsigned j0 = jb + jt, j1 = j0 + 64u;
        const bool ok0 = j0 < M, ok1 = j1 < M;
        cplx in0[8], in1[8];
        #pragma unroll
        for(int q = 0; q < 8; q++) { in0[q] = cplx(0.0, 0.0); in1[q] = cplx(0.0, 0.0); }
        for(unsigned ib = 0; ib < N; ib += RBM_SI) {
            __syncthreads();
            for(unsigned e = threadIdx.x; e < RBM_SI * RBM_ST; e += 256u) {
                const unsigned ii = e / RBM_ST, st = e % RBM_ST;
                const unsigned i = ib + ii;
                double sg = 0.0;
                if(sb + st < ns && i < N) sg = ((conf[(sb + st) * words + (i >> 6)] >> (i & 63u)) & 1ull) ? 1.0 : -1.0;
                sgn[ii][st] = sg;
            }
            __syncthreads();
            const unsigned ni = min((unsigned)RBM_SI, N - ib);
            for(unsigned ii = 0; ii < ni; ii++) {
                const size_t row = (size_t)(ib + ii) * M;
                const cplx v0 = ok0 ? v[row + j0] : cplx(0.0, 0.0);
                const cplx v1 = ok1 ? v[row + j1] : cplx(0.0, 0.0);
                #pragma unroll
                for(int q = 0; q < 8; q += 2) {
                    const double2 sg = *reinterpret_cast<const double2*>(&sgn[ii][sg8 + q]);
                    in0[q].re = fma(sg.x, v0.re, in0[q].re); in0[q].im = fma(sg.x, v0.im, in0[q].im);
                    in1[q].re = fma(sg.x, v1.re, in1[q].re); in1[q].im = fma(sg.x, v1.im, in1[q].im);
                    in0[q + 1].re = fma(sg.y, v0.re, in0[q + 1].re); in0[q + 1].im = fma(sg.y, v0.im, in0[q + 1].im);
                    in1[q + 1].re = fma(sg.y, v1.re, in1[q + 1].re); in1[q + 1].im = fma(sg.y, v1.im, in1[q + 1].im);
                }
            }
        }
        #pragma unroll
        for(int q = 0; q < 8; q++) {
            const size_t s = sb + sg8 + q;
            if(s < ns) {
                if(ok0) cfma(tot[q], T[s * M + j0], in0[q]);
                if(ok1) cfma(tot[q], T[s * M + j1], in1[q]);
            }
        }
    }
    // reduce over the 64 column-threads (2 warps) of each sample group
    #pragma unroll
    for(int q = 0; q < 8; q++) {
        const cplx r = warp_sum(tot[q]);
        if((threadIdx.x & 31u) == 0) red[sg8 + q][(threadIdx.x >> 5) & 1u] = r;
    }
    __syncthreads();
    if(threadIdx.x < RBM_ST && sb + threadIdx.x < ns) a[sb + threadIdx.x] = red[threadIdx.x][0] + red[threadIdx.x][1];
}

// ============================================================================================ FP64 tensor-core (DMMA) mat-vec
// The factorised S.v is two GEMMs against the +-1 spin matrix sigma [ns][N]:
//     U = sigma V            a_s = sum_j T_sj U_sj                     (k_rowdot_dmma)
//     Y = sigma^T Z          Z_sj = w_s a_s conj(T_sj)                 (k_colreduce_dmma, per-chunk partials)
// with complex matrices viewed as real ones with 2M columns ((re, im) adjacent).  Both run on the FP64 tensor cores
// (mma.sync.m8n8k4.f64: exact fp64 products and accumulation, so S.v keeps its 1e-10 parity), the +-1 operand is
// generated from the configuration bits in registers, the other operand is staged in shared memory.
// Fragment layout (PTX ISA, m8n8k4 .f64): A[row = lane>>2][k = lane&3], B[k = lane&3][col = lane>>2],
// C/D[row = lane>>2][col = 2*(lane&3) + {0,1}]  =>  each lane ends up with ONE complex number per 8x8 tile.
constexpr int RD_KC = 32, RD_PAD = 8;                                  // 8 samples per warp, RD_CB real columns per pass
constexpr size_t rd_smem(int cb) { return 2 * (size_t)RD_KC * (cb + RD_PAD) * sizeof(double); }   // double-buffered V tile (cp.async)

template<int RW, int RD_CB>                               // warps per block (8 samples each); real columns per pass (64 | 128)
__global__ void __launch_bounds__(RW * 32) k_rowdot_dmma(const uint64_t* __restrict__ conf, const cplx* __restrict__ T,
        const cplx* __restrict__ v, size_t ns, unsigned N, unsigned M, unsigned words, cplx* __restrict__ a_out) {
    extern __shared__ __align__(16) double rd_buf[];
    constexpr int RD_STRIDE = RD_CB + RD_PAD, RD_STAGE_DOUBLES = RD_KC * RD_STRIDE;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, row = lane >> 2, kq = lane & 3u;
    const size_t s = (size_t)blockIdx.x * (RW * 8) + warp * 8u + row;          // this lane's sample (A row / C row)
    uint64_t cw[MAXW] = {0ull, 0ull, 0ull, 0ull};
    #pragma unroll
    for(unsigned w = 0; w < (unsigned)MAXW; w++) if(s < ns && w < words) cw[w] = conf[s * words + w];
    const double* __restrict__ vr = reinterpret_cast<const double*>(v);
    const unsigned ncol = 2u * M;
    const unsigned nkc = (N + RD_KC - 1) / RD_KC, ncb = (ncol + RD_CB - 1) / RD_CB, nstage = nkc * ncb;
    // stage q = (column block q / nkc, site chunk q % nkc): 32 sites x 128 doubles, 16-byte cp.async chunks, zero-filled OOB
    auto issue = [&](unsigned q) {
        double* dst = rd_buf + (q & 1u) * RD_STAGE_DOUBLES;
        const unsigned cb = (q / nkc) * RD_CB, i0 = (q % nkc) * RD_KC;
        for(unsigned e = threadIdx.x; e < RD_KC * (RD_CB / 2); e += RW * 32) {
            const unsigned kk = e / (RD_CB / 2), c = (e % (RD_CB / 2)) * 2u;
            const bool ok = (i0 + kk < N) && (cb + c < ncol);
            cp_async16_zfill(dst + kk * RD_STRIDE + c, ok ? (const void*)(vr + (size_t)(i0 + kk) * ncol + cb + c) : (const void*)vr, ok);
        }
        cp_async_commit();
    };
    cplx tot(0.0, 0.0);
    double acc[RD_CB / 8][2];
    #pragma unroll
    for(int t = 0; t < RD_CB / 8; t++) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    issue(0);
    for(unsigned q = 0; q < nstage; q++) {
        if(q + 1u < nstage) { issue(q + 1u); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();
        const double* Vs = rd_buf + (q & 1u) * RD_STAGE_DOUBLES;
        const unsigned cb = (q / nkc) * RD_CB, i0 = (q % nkc) * RD_KC;
        #pragma unroll 2
        for(unsigned k4 = 0; k4 < RD_KC / 4; k4++) {
            const unsigned i = i0 + k4 * 4u + kq;
            const unsigned wd = i >> 6;
            uint64_t word = cw[0];
            if(wd == 1u) word = cw[1];
            if(wd == 2u) word = cw[2];
            if(wd == 3u) word = cw[3];
            const double asg = (s < ns && i < N) ? (((word >> (i & 63u)) & 1ull) ? 1.0 : -1.0) : 0.0;
            const double* vrow = Vs + (k4 * 4u + kq) * RD_STRIDE + row;
            #pragma unroll
            for(int t = 0; t < RD_CB / 8; t++) dmma(acc[t][0], acc[t][1], asg, vrow[t * 8]);
        }
        if(q % nkc == nkc - 1u) {                          // column block complete: a_s += sum_j T_sj U_sj, reset
            if(s < ns) {
                #pragma unroll
                for(int t = 0; t < RD_CB / 8; t++) {
                    const unsigned j = (cb >> 1) + (unsigned)t * 4u + kq;
                    if(j < M) cfma(tot, T[s * M + j], cplx(acc[t][0], acc[t][1]));
                }
            }
            #pragma unroll
            for(int t = 0; t < RD_CB / 8; t++) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
        }
        __syncthreads();                                   // the buffer is refilled by the next iteration's issue
    }
    tot.re += __shfl_xor_sync(FULL, tot.re, 1); tot.im += __shfl_xor_sync(FULL, tot.im, 1);
    tot.re += __shfl_xor_sync(FULL, tot.re, 2); tot.im += __shfl_xor_sync(FULL, tot.im, 2);
    if(kq == 0 && s < ns) a_out[s] = tot;
}

constexpr int CD_WARPS = 8, CD_KT = 32, CD_PAD = 8;   // tiles of 32 samples x CD_CB real columns
template<int ITW, int CD_CB>                            // ITW site tiles (of 8) per warp: N <= 64 * ITW; CD_CB in {64, 128}
__global__ void __launch_bounds__(CD_WARPS * 32) k_colreduce_dmma(const uint64_t* __restrict__ conf, const cplx* __restrict__ T,
        const double* __restrict__ w, const cplx* __restrict__ X, size_t ns, unsigned N, unsigned M, unsigned words, size_t chunk,
        cplx* __restrict__ part_x) {
    __shared__ __align__(16) double Zs[CD_KT][CD_CB + CD_PAD];
    __shared__ uint64_t sconf[CD_KT][MAXW];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, row = lane >> 2, kq = lane & 3u;
    const unsigned cb = blockIdx.x * CD_CB;                       // first real column of this block
    const unsigned j0 = cb >> 1;                                  // first complex column
    const size_t s0 = (size_t)blockIdx.y * chunk, s1 = min(ns, s0 + chunk);
    const unsigned ntile_i = (N + 7u) / 8u;
    double acc[ITW][CD_CB / 8][2];
    #pragma unroll
    for(int q = 0; q < ITW; q++)
        #pragma unroll
        for(int t = 0; t < CD_CB / 8; t++) { acc[q][t][0] = 0.0; acc[q][t][1] = 0.0; }
    // register prefetch of the next tile's z = w a conj(T): element e = threadIdx.x + 256 m -> (sample e / 64, column e % 64)
    constexpr int CD_ZPT = CD_KT * (CD_CB / 2) / (CD_WARPS * 32);      // z elements staged per thread
    cplx zreg[CD_ZPT];
    auto fetch = [&](size_t sb) {
        #pragma unroll
        for(int m = 0; m < CD_ZPT; m++) {
            const unsigned e = threadIdx.x + (unsigned)m * (CD_WARPS * 32);
            const unsigned st = e / (CD_CB / 2), jj = e % (CD_CB / 2);
            const size_t sidx = sb + st;
            cplx z(0.0, 0.0);
            if(sidx < s1 && j0 + jj < M) z = (w[sidx] * X[sidx]) * conj(T[sidx * M + j0 + jj]);
            zreg[m] = z;
        }
    };
    if(s0 < s1) fetch(s0);
    for(size_t sb = s0; sb < s1; sb += CD_KT) {
        __syncthreads();                                          // previous tile fully consumed
        #pragma unroll
        for(int m = 0; m < CD_ZPT; m++) {
            const unsigned e = threadIdx.x + (unsigned)m * (CD_WARPS * 32);
            const unsigned st = e / (CD_CB / 2), jj = e % (CD_CB / 2);
            *reinterpret_cast<double2*>(&Zs[st][2 * jj]) = make_double2(zreg[m].re, zreg[m].im);
        }
        if(threadIdx.x < CD_KT * MAXW) {
            const unsigned st = threadIdx.x / MAXW, wd = threadIdx.x % MAXW;
            const size_t sidx = sb + st;
            sconf[st][wd] = (sidx < s1 && wd < words) ? conf[sidx * words + wd] : 0ull;
        }
        __syncthreads();
        if(sb + CD_KT < s1) fetch(sb + CD_KT);                    // overlaps the global latency with the MMAs below
        #pragma unroll 2
        for(unsigned k4 = 0; k4 < CD_KT / 4; k4++) {
            const unsigned st = k4 * 4u + kq;
            double b[CD_CB / 8];
            #pragma unroll
            for(int t = 0; t < CD_CB / 8; t++) b[t] = Zs[st][t * 8 + row];
            const bool live = sb + st < s1;
            #pragma unroll
            for(int q = 0; q < ITW; q++) {
                const unsigned it = blockIdx.z * (CD_WARPS * ITW) + warp + (unsigned)q * CD_WARPS;
                if(it < ntile_i) {                                 // warp-uniform
                    const unsigned i = it * 8u + row;
                    const double asg = (live && i < N) ? bit_sign(sconf[st], i) : 0.0;
                    #pragma unroll
                    for(int t = 0; t < CD_CB / 8; t++) dmma(acc[q][t][0], acc[q][t][1], asg, b[t]);
                }
            }
        }
    }
    const size_t P = (size_t)N * M;
    #pragma unroll
    for(int q = 0; q < ITW; q++) {
        const unsigned it = blockIdx.z * (CD_WARPS * ITW) + warp + (unsigned)q * CD_WARPS;
        if(it < ntile_i) {
            const unsigned i = it * 8u + row;
            #pragma unroll
            for(int t = 0; t < CD_CB / 8; t++) {
                const unsigned j = j0 + (unsigned)t * 4u + kq;
                if(i < N && j < M) part_x[(size_t)blockIdx.y * P + (size_t)i * M + j] = cplx(acc[q][t][0], acc[q][t][1]);
            }
        }
    }
}

// F_k = F'_k - E conj(Obar_k)      (TDVP.cu.template:300, 331-333)
__global__ void k_finalize_F(const cplx* __restrict__ packed, unsigned P, cplx* __restrict__ F) {
    const cplx E = packed[0];
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < P; k += (size_t)gridDim.x * blockDim.x)
        F[k] = packed[2 + P + k] - E * conj(packed[2 + k]);
}
// out_k -= conj(Obar_k) * dot      (TDVP.cu.template:435-442)
__global__ void k_sv_correct(const cplx* __restrict__ Obar, const cplx* __restrict__ dot, unsigned P, cplx* __restrict__ out) {
    const cplx d = *dot;
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < P; k += (size_t)gridDim.x * blockDim.x)
        out[k] -= conj(Obar[k]) * d;
}

// ============================================================================================ S matrix
// S'[r][c] = sum_s w_s conj(O_sr) O_sc  — Hermitian rank-k update, upper-triangular 64x64 tiles, split over samples.
constexpr int ZT = 64, ZK = 8;
__global__ void __launch_bounds__(256) k_zherk(const cplx* __restrict__ O, const double* __restrict__ w, size_t ns, unsigned P,
                                               size_t chunk, cplx* __restrict__ Sout, size_t split_stride) {
    // decode the upper-triangular tile index
    const unsigned nt = (P + ZT - 1) / ZT;
    unsigned t = blockIdx.x, tr = 0;
    while(t >= nt - tr) { t -= nt - tr; tr++; }
    const unsigned tc = tr + t;
    const size_t s0 = (size_t)blockIdx.y * chunk, s1 = min(ns, s0 + chunk);
    __shared__ cplx A[ZK][ZT], B[ZK][ZT];
    const unsigned tx = threadIdx.x & 15u, ty = threadIdx.x >> 4;
    cplx acc[4][4];
    #pragma unroll
    for(int i = 0; i < 4; i++)
        #pragma unroll
        for(int j = 0; j < 4; j++) acc[i][j] = cplx(0.0, 0.0);
    for(size_t sb = s0; sb < s1; sb += ZK) {
        // 256 threads load 2 x (ZK x ZT) elements
        for(unsigned e = threadIdx.x; e < ZK * ZT; e += 256u) {
            const unsigned kk = e / ZT, col = e % ZT;
            const size_t s = sb + kk;
            cplx a(0.0, 0.0), b(0.0, 0.0);
            if(s < s1) {
                const unsigned r = tr * ZT + col, c = tc * ZT + col;
                if(r < P) a = w[s] * conj(O[s * P + r]);
                if(c < P) b = O[s * P + c];
            }
            A[kk][col] = a; B[kk][col] = b;
        }
        __syncthreads();
        #pragma unroll
        for(int kk = 0; kk < ZK; kk++) {
            cplx a[4], b[4];
            #pragma unroll
            for(int i = 0; i < 4; i++) { a[i] = A[kk][ty + 16 * i]; b[i] = B[kk][tx + 16 * i]; }
            #pragma unroll
            for(int i = 0; i < 4; i++)
                #pragma unroll
                for(int j = 0; j < 4; j++) cfma(acc[i][j], a[i], b[j]);
        }
        __syncthreads();
    }
    cplx* Sp = Sout + (size_t)blockIdx.y * split_stride;
    #pragma unroll
    for(int i = 0; i < 4; i++)
        #pragma unroll
        for(int j = 0; j < 4; j++) {
            const unsigned r = tr * ZT + ty + 16 * i, c = tc * ZT + tx + 16 * j;
            if(r < P && c < P) {
                Sp[(size_t)r * P + c] = acc[i][j];
                if(tr != tc) Sp[(size_t)c * P + r] = conj(acc[i][j]);
            }
        }
}
// S = sum_splits S' - conj(Obar_r) Obar_c     (TDVP.cu.template:293-299);  also adds nothing else
__global__ void k_S_finalize(const cplx* __restrict__ Spart, unsigned splits, size_t split_stride, const cplx* __restrict__ Obar,
                             unsigned P, cplx* __restrict__ S, bool subtract_mean) {
    const size_t total = (size_t)P * P;
    for(size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        cplx a(0.0, 0.0);
        for(unsigned c = 0; c < splits; c++) a += Spart[(size_t)c * split_stride + e];
        if(subtract_mean) { const size_t r = e / P, c = e % P; a -= conj(Obar[r]) * Obar[c]; }
        S[e] = a;
    }
}

// ============================================================================================ fused S.v tail and CG update
// out_k = sum_ch part[ch][k] - conj(Obar_k) * dot + (shift_abs + shift_rel * diag_k) * v_k      (diag may be null)
__global__ void k_sv_finish(const cplx* part, unsigned chunks, const cplx* __restrict__ Obar, const cplx* __restrict__ dot,
                            const cplx* __restrict__ v, const double* __restrict__ diag, double shift_abs, double shift_rel,
                            size_t P, cplx* out, bool correct) {   // part may alias out (chunks == 1)
    const cplx d = *dot;
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < P; k += (size_t)gridDim.x * blockDim.x) {
        cplx a(0.0, 0.0);
        for(unsigned c = 0; c < chunks; c++) a += part[(size_t)c * P + k];
        if(correct) a -= conj(Obar[k]) * d;
        if(diag) a += (shift_abs + shift_rel * diag[k]) * v[k];
        out[k] = a;
    }
}

// ============================================================================================ multi-block CG vector kernels
// (n > 64k, e.g. C5 with P = 320000): per-block partial sums in a fixed number of blocks, summed by every consumer in
// the same order (deterministic, no atomics).  scal: [0] rs, [1] pAp, [2] rs_new.
constexpr int VB_BLOCKS = 148, VB_T = 256;
template<int NV>
__device__ void block_reduce_small(double (&v)[NV], double* out) {      // VB_T threads
    __shared__ double sm[NV][VB_T / 32];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    #pragma unroll
    for(int i = 0; i < NV; i++) { v[i] = warp_sum(v[i]); if(lane == 0) sm[i][warp] = v[i]; }
    __syncthreads();
    if(warp == 0) {
        #pragma unroll
        for(int i = 0; i < NV; i++) {
            double x = (lane < VB_T / 32) ? sm[i][lane] : 0.0;
            x = warp_sum(x);
            if(lane == 0) out[i] = x;
        }
    }
    __syncthreads();
}
__device__ __forceinline__ cplx sum_partials(const cplx* part) {
    cplx t(0.0, 0.0);
    for(int b = 0; b < VB_BLOCKS; b++) t += part[b];
    return t;
}
// the same sum computed once per block by its first warp (fixed order: lane-strided, then the shuffle tree) and broadcast;
// every thread of the block must call it
__device__ __forceinline__ cplx sum_partials_block(const cplx* part) {
    __shared__ cplx bc;
    if(threadIdx.x < 32u) {
        cplx t(0.0, 0.0);
        for(int b = (int)threadIdx.x; b < VB_BLOCKS; b += 32) t += part[b];
        t = warp_sum(t);
        if(threadIdx.x == 0) bc = t;
    }
    __syncthreads();
    const cplx r = bc;
    __syncthreads();
    return r;
}
// part[b] = sum over the block's elements of op(a) b
template<bool CONJ_A>
__global__ void __launch_bounds__(VB_T) k_dot_part(const cplx* __restrict__ a, const cplx* __restrict__ b, size_t n, cplx* __restrict__ part,
                                                   cplx* scal_roll) {
    if(scal_roll && blockIdx.x == 0 && threadIdx.x == 0) scal_roll[0] = scal_roll[2];        // rs <- rs_new of the last iteration
    double v[2] = {0, 0}, r[2];
    for(size_t k = (size_t)blockIdx.x * VB_T + threadIdx.x; k < n; k += (size_t)VB_BLOCKS * VB_T) {
        cplx x = a[k]; if(CONJ_A) x = conj(x);
        const cplx p = x * b[k];
        v[0] += p.re; v[1] += p.im;
    }
    block_reduce_small<2>(v, r);
    if(threadIdx.x == 0) part[blockIdx.x] = cplx(r[0], r[1]);
}
__global__ void k_sum_part_to_scalar(const cplx* __restrict__ part, cplx* out) { if(threadIdx.x == 0) *out = sum_partials(part); }
// (Preconditioned) CG vector updates.  scal[0] = (r.z, |r|^2) of the current residual, z = minv * r (Jacobi) or z = r
// (minv == null); scal[2] receives the same pair for the updated residual and is rolled into scal[0] by k_dot_part.
// alpha = r.z / sum(part_pAp); x += alpha p; r -= alpha Ap; part_rr[b] = (sum minv |r|^2, sum |r|^2)
__global__ void __launch_bounds__(VB_T) k_cg_xr_mb(cplx* __restrict__ x, cplx* __restrict__ r, const cplx* __restrict__ p, const cplx* __restrict__ Ap,
                                                   const cplx* __restrict__ scal, const cplx* __restrict__ part_pAp, cplx* __restrict__ part_rr,
                                                   const double* __restrict__ minv, size_t n) {
    const double alpha = scal[0].re / sum_partials_block(part_pAp).re;
    double v[2] = {0, 0}, red[2];
    for(size_t k = (size_t)blockIdx.x * VB_T + threadIdx.x; k < n; k += (size_t)VB_BLOCKS * VB_T) {
        x[k] += alpha * p[k];
        const cplx rk = r[k] - alpha * Ap[k];
        r[k] = rk;
        const double a2 = abs2(rk);
        v[1] += a2;
        v[0] += minv ? minv[k] * a2 : a2;
    }
    block_reduce_small<2>(v, red);
    if(threadIdx.x == 0) part_rr[blockIdx.x] = cplx(red[0], red[1]);
}
// (rz_new, rr_new) = sum(part_rr); beta = rz_new / rz; p = z + beta p; scal[2] = (rz_new, rr_new);
// part_dot[b] (optional) = partial sums of Obar . p_new, the scalar the next S.v needs
__global__ void __launch_bounds__(VB_T) k_cg_p_mb(cplx* __restrict__ p, const cplx* __restrict__ r, cplx* __restrict__ scal,
                                                  const cplx* __restrict__ part_rr, const double* __restrict__ minv,
                                                  const cplx* __restrict__ Obar, cplx* __restrict__ part_dot, size_t n) {
    const cplx rs_new = sum_partials_block(part_rr);
    const double beta = rs_new.re / scal[0].re;
    double v[2] = {0, 0}, red[2];
    for(size_t k = (size_t)blockIdx.x * VB_T + threadIdx.x; k < n; k += (size_t)VB_BLOCKS * VB_T) {
        const cplx z = minv ? minv[k] * r[k] : r[k];
        const cplx pk = z + beta * p[k];
        p[k] = pk;
        if(part_dot) { const cplx d = Obar[k] * pk; v[0] += d.re; v[1] += d.im; }
    }
    if(part_dot) {
        block_reduce_small<2>(v, red);
        if(threadIdx.x == 0) part_dot[blockIdx.x] = cplx(red[0], red[1]);
    }
    if(blockIdx.x == 0 && threadIdx.x == 0) scal[2] = rs_new;
}
// residual refresh of the CG with tensor-core products: r = b - (A x computed exactly); part_rr as k_cg_xr_mb leaves it
__global__ void __launch_bounds__(VB_T) k_cg_refresh(cplx* __restrict__ r, const cplx* __restrict__ b, const cplx* __restrict__ Ax,
                                                     const double* __restrict__ minv, cplx* __restrict__ part_rr, size_t n) {
    double v[2] = {0, 0}, red[2];
    for(size_t k = (size_t)blockIdx.x * VB_T + threadIdx.x; k < n; k += (size_t)VB_BLOCKS * VB_T) {
        const cplx rk = b[k] - Ax[k];
        r[k] = rk;
        const double a2 = abs2(rk);
        v[1] += a2;
        v[0] += minv ? minv[k] * a2 : a2;
    }
    block_reduce_small<2>(v, red);
    if(threadIdx.x == 0) part_rr[blockIdx.x] = cplx(red[0], red[1]);
}
// S.v epilogue fused with the CG scalar work: Ap = sum of the column-reduction chunks - conj(Obar) (Obar . p) + shift p,
// part_pAp[b] = partial sums of conj(p) . Ap; rolls (r.z, |r|^2) of the previous iteration into scal[0]
__global__ void __launch_bounds__(VB_T) k_sv_finish_cg(const cplx* part, unsigned chunks, const cplx* __restrict__ Obar,
        const cplx* __restrict__ part_dot, const cplx* __restrict__ v, const double* __restrict__ diag, double shift_abs, double shift_rel,
        size_t P, cplx* out, cplx* __restrict__ part_pAp, cplx* scal_roll) {      // part may alias out (chunks == 1)
    if(scal_roll && blockIdx.x == 0 && threadIdx.x == 0) scal_roll[0] = scal_roll[2];
    const cplx d = sum_partials_block(part_dot);
    double acc[2] = {0, 0}, red[2];
    for(size_t k = (size_t)blockIdx.x * VB_T + threadIdx.x; k < P; k += (size_t)VB_BLOCKS * VB_T) {
        cplx a(0.0, 0.0);
        for(unsigned c = 0; c < chunks; c++) a += part[(size_t)c * P + k];
        a -= conj(Obar[k]) * d;
        const cplx vk = v[k];
        a += (shift_abs + shift_rel * diag[k]) * vk;
        out[k] = a;
        const cplx q = conj(vk) * a;
        acc[0] += q.re; acc[1] += q.im;
    }
    block_reduce_small<2>(acc, red);
    if(threadIdx.x == 0) part_pAp[blockIdx.x] = cplx(red[0], red[1]);
}
// start of the iteration: p = z = minv * r (or r); scal0 = (r.z, |r|^2).  One block (runs once per solve).
__global__ void __launch_bounds__(RED_T) k_cg_init(cplx* __restrict__ p, const cplx* __restrict__ r, const double* __restrict__ minv, size_t n, cplx* scal0,
                                                   const cplx* __restrict__ Obar, cplx* __restrict__ part_dot) {
    double v[4] = {0, 0, 0, 0}, red[4];
    for(size_t k = threadIdx.x; k < n; k += RED_T) {
        const cplx rk = r[k];
        const double a2 = abs2(rk);
        const cplx pk = minv ? minv[k] * rk : rk;
        p[k] = pk;
        v[1] += a2; v[0] += minv ? minv[k] * a2 : a2;
        const cplx d = Obar[k] * pk; v[2] += d.re; v[3] += d.im;
    }
    block_reduce<4>(v, red);
    if(threadIdx.x == 0) { *scal0 = cplx(red[0], red[1]); part_dot[0] = cplx(red[2], red[3]); }
    for(int b = 1 + (int)threadIdx.x; b < VB_BLOCKS; b += RED_T) part_dot[b] = cplx(0.0, 0.0);
}
// minv_k = 1 / (diag_k + shift_abs + shift_rel * diag_k): the inverse diagonal of the shifted S
__global__ void k_jacobi_inverse(const double* __restrict__ diag, double shift_abs, double shift_rel, size_t n, double* __restrict__ minv) {
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const double m = diag[k] + shift_abs + shift_rel * diag[k];
        minv[k] = m > 0.0 ? 1.0 / m : 1.0;
    }
}

// ============================================================================================ solver kernels
// scal layout (cplx): [0] rs_old, [1] pAp, [2] rs_new, [3] scratch
__global__ void k_cg_xr(cplx* x, cplx* r, const cplx* p, const cplx* Ap, const cplx* scal, size_t n) {
    const double alpha = scal[0].re / scal[1].re;
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        x[k] += alpha * p[k];
        r[k] -= alpha * Ap[k];
    }
}
__global__ void k_cg_p(cplx* p, const cplx* r, cplx* scal, size_t n) {
    const double beta = scal[2].re / scal[0].re;
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        p[k] = r[k] + beta * p[k];
}
__global__ void k_cg_roll(cplx* scal) { scal[0] = scal[2]; }
__global__ void k_add_shift(cplx* Ap, const cplx* p, const double* diag, double shift_abs, double shift_rel, size_t n) {
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        Ap[k] += (shift_abs + shift_rel * diag[k]) * p[k];
}
__global__ void k_scale_vec(const cplx* in, cplx phase, cplx* out, size_t n, bool conj_in) {
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        cplx v = in[k]; if(conj_in) v = conj(v);
        out[k] = phase * v;
    }
}
__global__ void k_conj_inplace(cplx* v, size_t n) {
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) v[k] = conj(v[k]);
}
// diag_k = sum_s w_s |O_sk|^2 - |Obar_k|^2 : per-chunk partials (blockIdx.y), summed in fixed order afterwards
__global__ void k_diag_dense(const cplx* __restrict__ O, const double* __restrict__ w, size_t ns, unsigned P, size_t chunk, double* __restrict__ part) {
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= P) return;
    const size_t s0 = (size_t)blockIdx.y * chunk, s1 = min(ns, s0 + chunk);
    double a = 0.0;
    for(size_t s = s0; s < s1; s++) a = fma(w[s], abs2(O[s * P + k]), a);
    part[(size_t)blockIdx.y * P + k] = a;
}
__global__ void k_diag_rbm(const cplx* __restrict__ T, const double* __restrict__ w, size_t ns, unsigned M, size_t chunk, double* __restrict__ part) {
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if(j >= M) return;
    const size_t s0 = (size_t)blockIdx.y * chunk, s1 = min(ns, s0 + chunk);
    double a = 0.0;
    for(size_t s = s0; s < s1; s++) a = fma(w[s], abs2(T[s * M + j]), a);
    part[(size_t)blockIdx.y * M + j] = a;
}
// d[k] = sum_ch part[ch][k % period]   (period = M broadcasts the RBM's per-j value over the N sites)
__global__ void k_diag_sum(const double* __restrict__ part, unsigned chunks, unsigned period, size_t n, double* __restrict__ d) {
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        double a = 0.0;
        for(unsigned c = 0; c < chunks; c++) a += part[(size_t)c * period + (k % period)];
        d[k] = a;
    }
}
__global__ void k_diag_finalize(double* d, const cplx* Obar, size_t n) {
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) d[k] -= abs2(Obar[k]);
}
__global__ void k_add_diag_shift(cplx* A, const double* diag, double shift_abs, double shift_rel, unsigned P) {
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < P; k += (size_t)gridDim.x * blockDim.x)
        A[k * (size_t)P + k].re += shift_abs + shift_rel * diag[k];
}
__global__ void k_exp_inplace(cplx* v, size_t n) {
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) v[k] = cexp(v[k]);
}
__global__ void k_mul_exp(const cplx* log_psi, const cplx* eloc, cplx* out, size_t n) {
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) out[k] = cexp(log_psi[k]) * eloc[k];
}
__global__ void k_sum_rows(const cplx* __restrict__ O, size_t ns, unsigned P, cplx* __restrict__ out) {
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= P) return;
    cplx a(0.0, 0.0);
    for(size_t s = 0; s < ns; s++) a += O[s * P + k];
    out[k] = a;
}

static inline unsigned grid_for(size_t n, unsigned block = 256) {
    return (unsigned)std::max<size_t>(1, std::min<size_t>((n + block - 1) / block, (size_t)ctx().num_sms * 32));
}
static void sum_chunks(const cplx* part_a, cplx* out_a, const cplx* part_b, cplx* out_b, unsigned chunks, size_t P, bool conj_b) {
    if(chunks >= 32u && (P + 255) / 256 < (size_t)ctx().num_sms) {
        k_sum_chunks_wide<<<dim3((unsigned)((P + 31) / 32), part_b ? 2 : 1), 1024, 0, stream()>>>(part_a, part_b, chunks, P, out_a, out_b, conj_b);
        ANGPU_CHECK_LAUNCH(); count_launch();
        return;
    }
    k_sum_chunks<<<grid_for(P), 256, 0, stream()>>>(part_a, chunks, P, out_a);
    if(part_b) k_sum_chunks<<<grid_for(P), 256, 0, stream()>>>(part_b, chunks, P, out_b, conj_b);
    ANGPU_CHECK_LAUNCH(); count_launch(part_b ? 2 : 1);
}

// ============================================================================================ Ensemble

void Ensemble::generate(Psi& psi, SampleSet& S) {
    // collectives of this evaluation run iff the ensemble is sharded (an unsharded ensemble is never summed over ranks;
    // a sharded one without a transport reports this shard's partial sums -- rank emulation in one process, tests)
    set_reduce(world > 1);
    // MonteCarloPaulis / ExactSummationPaulis (Pauli-string basis, pauli_basis.cuh): PsiDeep with N = 3 num_sites input units
    ANGPU_REQUIRE(paulis ? psi.pauli_sites != 0u : psi.pauli_sites == 0u,
                  paulis ? "Pauli-string ensembles need a PsiDeep with N = 3 num_sites input units" : "a PsiDeep on the Pauli-string basis needs MonteCarloPaulis / ExactSummationPaulis");
    S.pauli_sites = psi.pauli_sites;
    if(is_mc) {
        ANGPU_REQUIRE(num_chains >= 1 && num_samples >= 1, "MonteCarlo: num_samples and num_markov_chains must be positive");
        size_t c0, cn; shard(num_chains, c0, cn);
        McParams mc{};
        mc.num_samples = num_samples; mc.num_sweeps = num_sweeps; mc.num_therm = num_therm;
        mc.steps_per_chain = (unsigned)(num_samples / num_chains);
        mc.num_chains_local = (unsigned)cn; mc.chain0 = (unsigned)c0;
        mc.seed_lo = (unsigned)seed; mc.seed_hi = (unsigned)(seed >> 32); mc.call = call; mc.pauli_sites = psi.pauli_sites;
        S.resize((size_t)mc.steps_per_chain * cn, psi.words);
        d_acc_rej.resize(4); d_acc_rej.zero();
        psi.mc_sample(mc, S, d_acc_rej.p);
        if(S.ns) { k_fill<<<grid_for(S.ns), 256, 0, stream()>>>(S.weight.p, 1.0 / (double)num_samples, S.ns); ANGPU_CHECK_LAUNCH(); count_launch(); }
        call++;
    } else if(paulis) {
        ANGPU_REQUIRE(num_sites == psi.pauli_sites, "ExactSummationPaulis: num_sites differs from the wavefunction's");
        ANGPU_REQUIRE(num_sites <= 20u, "ExactSummationPaulis: too many sites (4^num_sites configurations)");
        size_t b, n; shard((size_t)1 << (2u * num_sites), b, n);
        S.resize(n, psi.words);
        if(n) { k_enumerate_paulis<<<grid_for(n), 256, 0, stream()>>>(S.conf.p, b, n, num_sites, psi.words); ANGPU_CHECK_LAUNCH(); count_launch(); }
        psi.log_psi(S, true);
    } else {
        ANGPU_REQUIRE(num_sites == psi.N, "ExactSummation: num_sites differs from the wavefunction's");
        ANGPU_REQUIRE(num_sites <= 40u, "ExactSummation: too many sites");
        size_t b, n; shard((size_t)1 << num_sites, b, n);
        S.resize(n, psi.words);
        if(n) { k_enumerate<<<grid_for(n), 256, 0, stream()>>>(S.conf.p, b, n, psi.words); ANGPU_CHECK_LAUNCH(); count_launch(); }
        psi.log_psi(S, true);
    }
}
__global__ void k_reweight(double* __restrict__ w, const cplx* __restrict__ lp, const cplx* __restrict__ lp_sampling, size_t n) {
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        w[k] *= exp(2.0 * (lp[k].re - lp_sampling[k].re));
}
void Ensemble::generate_reweighted(Psi& psi, Psi& psi_sampling, SampleSet& S) {
    ANGPU_REQUIRE(psi.N == psi_sampling.N, "reweighting: psi and psi_sampling act on different numbers of sites");
    generate(psi_sampling, S);
    if(S.ns == 0) return;
    lp_sampling.resize(S.ns);
    ANGPU_CUDA(cudaMemcpyAsync(lp_sampling.p, S.log_psi.p, sizeof(cplx) * S.ns, cudaMemcpyDeviceToDevice, stream()));
    S.has_angles = false;                       // any cached first-layer angles belong to psi_sampling
    psi.log_psi(S, false);
    k_reweight<<<grid_for(S.ns), 256, 0, stream()>>>(S.weight.p, S.log_psi.p, lp_sampling.p, S.ns);
    ANGPU_CHECK_LAUNCH(); count_launch();
}
void Ensemble::acceptance(unsigned long long out[2]) {
    out[0] = out[1] = 0;
    if(d_acc_rej.n >= 2) d_acc_rej.download(out, 2);
}

// ============================================================================================ ExpectationValue

static void scalar_sums(const SampleSet& S, bool with_eloc, bool with_lp, double* out4_dev, double* lp2_dev) {
    k_scalar_sums<<<1, RED_T, 0, stream()>>>(S.weight.p, with_eloc ? S.eloc.p : nullptr, with_lp ? S.log_psi.p : nullptr, S.ns, out4_dev, lp2_dev);
    ANGPU_CHECK_LAUNCH(); count_launch();
}

cplx ExpectationValue::value(const Operator& op, Psi& psi, Ensemble& ens) {
    double f; cplx m; fluctuation(op, psi, ens, f, m); return m;
}
void ExpectationValue::fluctuation(const Operator& op, Psi& psi, Ensemble& ens, double& fluct, cplx& mean) {
    ens.generate(psi, S);
    psi.eloc(op, S);
    d_scal.resize(8);
    scalar_sums(S, true, false, d_scal.p, nullptr);
    allreduce_sum(d_scal.p, 4);
    double h[4]; d_scal.download(h, 4);
    mean = cplx(h[0], h[1]);
    fluct = std::sqrt(h[2] - abs2(mean));
}

cplx ExpectationValue::value_reweighted(const Operator& op, Psi& psi, Psi& psi_sampling, Ensemble& ens) {
    ens.generate_reweighted(psi, psi_sampling, S);
    psi.eloc(op, S);
    d_scal.resize(8);
    scalar_sums(S, true, false, d_scal.p, nullptr);
    allreduce_sum(d_scal.p, 4);
    double h[4]; d_scal.download(h, 4);
    return cplx(h[0] / h[3], h[1] / h[3]);
}
// one warp per sample: eloc[s] = exp(sum over ALL strings of c_n <s|P_n|s'> coefficient)  (Operator.hpp:125-157)
__global__ void k_exp_fast_energy(const OpDev op, const uint64_t* __restrict__ confs, size_t ns, cplx* __restrict__ out) {
    const unsigned lane = threadIdx.x & 31u, wpb = blockDim.x >> 5;
    for(size_t s = (size_t)blockIdx.x * wpb + (threadIdx.x >> 5); s < ns; s += (size_t)gridDim.x * wpb) {
        const cplx e = fast_local_energy_warp(op, confs + s * op.words);
        if(lane == 0) out[s] = cexp(e);
    }
}
cplx ExpectationValue::exp_sigma_z(const Operator& op, Psi& psi, Ensemble& ens) {
    ANGPU_REQUIRE(!ens.paulis, "exp_sigma_z: spin-basis ensembles only (fast_local_energy of a diagonal operator on Spins)");
    require_operator_fits(op, psi.N, psi.words);
    ens.generate(psi, S);
    if(S.ns) {
        k_exp_fast_energy<<<grid_for(S.ns * 32), 256, 0, stream()>>>(op.dev, S.conf.p, S.ns, S.eloc.p);
        ANGPU_CHECK_LAUNCH(); count_launch();
    }
    d_scal.resize(8);
    scalar_sums(S, true, false, d_scal.p, nullptr);
    allreduce_sum(d_scal.p, 4);
    double h[4]; d_scal.download(h, 4);
    return cplx(h[0], h[1]);
}

// ============================================================================================ TDVP

// ANGPU_MATVEC=fma selects the plain-FMA factorised mat-vec kernels (A/B comparison); default: FP64 tensor cores
static bool use_dmma() {
    static int v = -1;
    if(v < 0) { const char* e = getenv("ANGPU_MATVEC"); v = (e && std::string(e) == "fma") ? 0 : 1; }
    return v == 1;
}

static unsigned pick_chunks(size_t ns, size_t col_blocks) {
    // enough blocks to fill the GPU twice, chunk >= 32 samples, <= 128 chunks
    const size_t want = ((size_t)ctx().num_sms * 8 + col_blocks - 1) / std::max<size_t>(1, col_blocks);
    size_t ch = std::max<size_t>(1, std::min<size_t>(want, 128));
    ch = std::min<size_t>(ch, std::max<size_t>(1, ns / 32));
    return (unsigned)ch;
}

// per-chunk partial sums of mean_k = sum_s w_s O_sk and x_k = sum_s w_s X_s conj(O_sk) into t.chunk_buf
struct ColPartials { unsigned chunks; cplx* mean; cplx* x; };
static ColPartials col_reduce_partials(TDVP& t, const cplx* X, bool want_mean) {
    const size_t ns = t.S.ns; const unsigned P = t.P;
    if(ns == 0) {
        t.chunk_buf.resize((size_t)2 * P); t.chunk_buf.zero();
        return ColPartials{1u, want_mean ? t.chunk_buf.p : nullptr, t.chunk_buf.p + P};
    }
    unsigned chunks; size_t chunk;
    if(t.factorised) {
        const unsigned jt = std::min(128u, (t.rbm_M + 31u) / 32u * 32u);      // threads per block of k_col_reduce_rbm
        const unsigned jb = ceil_div(t.rbm_M, jt), ib = ceil_div(t.rbm_N, RBM_IT);
        // chunks of whole 32-sample tiles: enough blocks to fill the GPU twice, but the partial-sum traffic
        // (chunks * P * 16 B written and read back) is kept below ~64 MB and chunks <= 64
        const size_t colblocks = (2 * (size_t)t.rbm_M + 63) / 64;     // k_colreduce_dmma blocks per chunk
        // measured: C2 (N = 64, 80-register variant) is best with ~2 blocks per SM, C5 (N = 200) with ~4
        const char* env_cf = getenv("ANGPU_CHUNK_FACTOR");
        const size_t cf = env_cf ? (size_t)atoi(env_cf) : (t.rbm_N <= 64u ? 2 : 4);
        size_t want_chunks = ((size_t)ctx().num_sms * cf + colblocks - 1) / colblocks;
        want_chunks = std::min<size_t>(want_chunks, std::max<size_t>(1, ((size_t)64 << 20) / ((size_t)P * sizeof(cplx))));
        size_t max_chunks = 64;
        if(want_mean) {
            // k_col_reduce_rbm<true> launches jb x ib blocks per chunk: enough chunks for ~4 blocks per SM (C1: one block per
            // chunk -> 512 chunks instead of 64, 635 -> ~80 us), bounded by the partial-sum traffic (2 x chunks x P x 16 B <= 64 MB)
            want_chunks = ((size_t)ctx().num_sms * 4 + (size_t)jb * ib - 1) / ((size_t)jb * ib);
            want_chunks = std::min<size_t>(want_chunks, std::max<size_t>(1, ((size_t)32 << 20) / ((size_t)P * sizeof(cplx))));
            want_chunks = std::max<size_t>(want_chunks, 1);
            max_chunks = 4096;
        }
        chunks = (unsigned)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(max_chunks, want_chunks), (ns + RBM_TS - 1) / RBM_TS));
        chunk = ((ns + chunks - 1) / chunks + RBM_TS - 1) / RBM_TS * RBM_TS; chunks = (unsigned)((ns + chunk - 1) / chunk);
        t.chunk_buf.resize((size_t)2 * chunks * P);
        cplx* pm = want_mean ? t.chunk_buf.p : nullptr; cplx* px = t.chunk_buf.p + (size_t)chunks * P;
        if(want_mean) k_col_reduce_rbm<true><<<dim3(jb, ib, chunks), jt, 0, stream()>>>(t.S.conf.p, t.T.p, t.S.weight.p, X, ns, t.rbm_N, t.rbm_M, t.words, chunk, pm, px);
        else if(use_dmma()) {
            const dim3 g128(ceil_div(2 * t.rbm_M, 128), chunks), g64(ceil_div(2 * t.rbm_M, 64), chunks);
            if(t.rbm_N <= 64u) k_colreduce_dmma<1, 64><<<g64, CD_WARPS * 32, 0, stream()>>>(t.S.conf.p, t.T.p, t.S.weight.p, X, ns, t.rbm_N, t.rbm_M, t.words, chunk, px);
            else if(t.rbm_N <= 128u) k_colreduce_dmma<2, 64><<<g64, CD_WARPS * 32, 0, stream()>>>(t.S.conf.p, t.T.p, t.S.weight.p, X, ns, t.rbm_N, t.rbm_M, t.words, chunk, px);
            else {
                // wider lattices: 16 site tiles (128 sites) per block along grid.z instead of 4 tiles per warp in one block
                // (183 registers, one block per SM): C5 (N = 200) runs 2 z-slices of the <2,64> build
                const dim3 g64z(g64.x, g64.y, ceil_div((t.rbm_N + 7u) / 8u, CD_WARPS * 2));
                k_colreduce_dmma<2, 64><<<g64z, CD_WARPS * 32, 0, stream()>>>(t.S.conf.p, t.T.p, t.S.weight.p, X, ns, t.rbm_N, t.rbm_M, t.words, chunk, px);
            }
        }
        else k_col_reduce_rbm<false><<<dim3(jb, ib, chunks), jt, 0, stream()>>>(t.S.conf.p, t.T.p, t.S.weight.p, X, ns, t.rbm_N, t.rbm_M, t.words, chunk, pm, px);
        ANGPU_CHECK_LAUNCH(); count_launch();
        return ColPartials{chunks, pm, px};
    }
    const unsigned cb = ceil_div(P, 128);
    chunks = pick_chunks(ns, cb); chunk = (ns + chunks - 1) / chunks; chunks = (unsigned)((ns + chunk - 1) / chunk);
    t.chunk_buf.resize((size_t)2 * chunks * P);
    cplx* pm = want_mean ? t.chunk_buf.p : nullptr; cplx* px = t.chunk_buf.p + (size_t)chunks * P;
    k_col_reduce_dense<<<dim3(cb, chunks), 128, 0, stream()>>>(t.O.p, t.S.weight.p, X, ns, P, chunk, pm, px);
    ANGPU_CHECK_LAUNCH(); count_launch();
    return ColPartials{chunks, pm, px};
}
// mean_out / x_out: [P] device; X: per-sample complex factor
static void col_reduce(TDVP& t, const cplx* X, cplx* mean_out, cplx* x_out) {
    // factorised rows with enough sites to fill the 8-row MMA tiles: both sums on the FP64 tensor cores, as two passes of
    // the same kernel (x with the given X; the mean as conj(sum_s w_s conj(O_sk)), i.e. X = 1)
    if(mean_out && t.factorised && use_dmma() && t.rbm_N >= 32u && t.S.ns >= 1024) {
        const ColPartials cx = col_reduce_partials(t, X, false);
        k_sum_chunks<<<grid_for(t.P), 256, 0, stream()>>>(cx.x, cx.chunks, t.P, x_out);
        if(t.ones.n < t.S.ns) {
            t.ones.resize(t.S.ns);
            k_fill_cplx<<<grid_for(t.S.ns), 256, 0, stream()>>>(t.ones.p, cplx(1.0, 0.0), t.S.ns);
            count_launch();
        }
        const ColPartials cm = col_reduce_partials(t, t.ones.p, false);
        k_sum_chunks<<<grid_for(t.P), 256, 0, stream()>>>(cm.x, cm.chunks, t.P, mean_out, true);
        ANGPU_CHECK_LAUNCH(); count_launch(2);
        return;
    }
    const ColPartials cp = col_reduce_partials(t, X, mean_out != nullptr);
    sum_chunks(cp.x, x_out, mean_out ? cp.mean : nullptr, mean_out, cp.chunks, t.P);
}

void TDVP::mark(int i) {
    if(!profile) return;
    if(!ev[i]) ANGPU_CUDA(cudaEventCreate(&ev[i]));
    ANGPU_CUDA(cudaEventRecord(ev[i], stream()));
}
TDVP::~TDVP() { for(auto& e : ev) if(e) cudaEventDestroy(e); }

void TDVP::eval(const Operator& op, Psi& psi, Ensemble& ens, bool want_S, Psi* psi_sampling) {
    ANGPU_REQUIRE(psi.P == P, "TDVP: num_params differs from the wavefunction's");
    last_psi = &psi; words = psi.words; num_steps_global = ens.num_steps();
    mark(0);
    if(psi_sampling) ens.generate_reweighted(psi, *psi_sampling, S);
    else ens.generate(psi, S);
    sharded = ens.world > 1;
    mark(1);
    psi.eloc(op, S);
    mark(2);
    packed.resize(2 + 2 * (size_t)P);
    scalar_sums(S, true, false, reinterpret_cast<double*>(packed.p), nullptr);
    prepare_rows(psi, want_S);
    col_reduce(*this, S.eloc.p, packed.p + 2, packed.p + 2 + P);
    allreduce_sum(reinterpret_cast<double*>(packed.p), 2 * (2 + 2 * (size_t)P));
    mark(3);
    cplx h[2]; packed.download(h, 2);
    E = h[0]; E2 = h[1].re; total_weight = h[1].im;
    F.resize(P);
    k_finalize_F<<<grid_for(P), 256, 0, stream()>>>(packed.p, P, F.p);
    ANGPU_CHECK_LAUNCH(); count_launch();
    evaluated = true;
    mark(4);
    if(profile) {
        ANGPU_CUDA(cudaEventSynchronize(ev[4]));
        for(int i = 0; i < 3; i++) ANGPU_CUDA(cudaEventElapsedTime(&phase_ms[i], ev[i], ev[i + 1]));
        ANGPU_CUDA(cudaEventElapsedTime(&phase_ms[3], ev[0], ev[4]));
    }
    if(want_S) build_S();
}

void TDVP::prepare_rows(Psi& psi, bool dense) {
    ANGPU_REQUIRE(psi.P == P, "TDVP: num_params differs from the wavefunction's");
    last_psi = &psi; words = psi.words;
    have_S = false; tc_ready = false;
    if(psi.kind == Psi::RBM && !dense) {
        PsiRBM& rbm = static_cast<PsiRBM&>(psi);
        rbm.compute_T(S, T);
        factorised = true; have_dense_O = false; rbm_N = rbm.N; rbm_M = rbm.M;
    } else {
        O.resize(S.ns * (size_t)P);
        psi.ok_rows(S, 0, S.ns, O.p);
        factorised = false; have_dense_O = true;
    }
}
void TDVP::weighted_conj_column_sums(const cplx* X, cplx* x_out) { col_reduce(*this, X, nullptr, x_out); }

void TDVP::ensure_dense_O(Psi* psi) {
    if(have_dense_O) return;
    ANGPU_REQUIRE(evaluated && psi, "TDVP: no samples (call eval first)");
    ANGPU_REQUIRE(Psi::is_live(psi), "TDVP: the psi of the last eval was destroyed before its O_k rows were materialised (keep it alive until get_S / get_O_k_samples / solve_dense / build_S_tensorcore, or call them before angpu_psi_destroy)");
    O.resize(S.ns * (size_t)P);
    psi->ok_rows(S, 0, S.ns, O.p);
    have_dense_O = true;
}

void TDVP::build_S() {
    set_reduce(sharded);
    ANGPU_REQUIRE(evaluated, "TDVP: call eval first");
    ensure_dense_O(last_psi);
    mark(5);
    Smat.resize((size_t)P * P);
    const unsigned nt = (P + ZT - 1) / ZT, tiles = nt * (nt + 1) / 2;
    const size_t ns = S.ns;
    // split the sample range when there are too few tiles to fill the GPU (bounded scratch: <= 1 GiB)
    unsigned splits = 1;
    if(ns > 0) {
        const size_t want = ((size_t)ctx().num_sms * 2 + tiles - 1) / tiles;
        const size_t mem_cap = std::max<size_t>(1, ((size_t)1 << 30) / (sizeof(cplx) * (size_t)P * P));
        splits = (unsigned)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(want, mem_cap), std::max<size_t>(1, ns / ZK)));
    }
    const size_t chunk = ns ? (ns + splits - 1) / splits : 1;
    splits = ns ? (unsigned)((ns + chunk - 1) / chunk) : 1;
    const size_t stride = (size_t)P * P;
    cplx* part = Smat.p;
    if(splits > 1) { cg_buf.resize(stride * splits); part = cg_buf.p; }
    if(ns) {
        // ANGPU_SBUILD=fma selects the plain-FMA tile kernel (A/B comparison); default: FP64 tensor cores
        const char* env = getenv("ANGPU_SBUILD");
        if(env && std::string(env) == "fma") k_zherk<<<dim3(tiles, splits), 256, 0, stream()>>>(O.p, S.weight.p, ns, P, chunk, part, stride);
        else {
            static bool attr_set = false;
            if(!attr_set) {
                ANGPU_CUDA(cudaFuncSetAttribute(k_zherk_dmma<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZD_SMEM));
                ANGPU_CUDA(cudaFuncSetAttribute(k_zherk_dmma<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZD_SMEM));
                attr_set = true;
            }
            if(env && std::string(env) == "dmma8") k_zherk_dmma<2, false><<<dim3(tiles, splits), 256, ZD_SMEM, stream()>>>(reinterpret_cast<const double*>(O.p), 2 * (size_t)P, S.weight.p, ns, P, chunk, part, P, stride);
            else k_zherk_dmma<4, false><<<dim3(tiles, splits), 512, ZD_SMEM, stream()>>>(reinterpret_cast<const double*>(O.p), 2 * (size_t)P, S.weight.p, ns, P, chunk, part, P, stride);
        }
        ANGPU_CHECK_LAUNCH(); count_launch();
    } else Smat.zero();
    if(splits > 1) {
        k_S_finalize<<<grid_for(stride), 256, 0, stream()>>>(part, splits, stride, Ok_dev(), P, Smat.p, false);
        ANGPU_CHECK_LAUNCH(); count_launch();
    }
    allreduce_sum(reinterpret_cast<double*>(Smat.p), 2 * stride);
    k_S_finalize<<<grid_for(stride), 256, 0, stream()>>>(Smat.p, 1, stride, Ok_dev(), P, Smat.p, true);
    ANGPU_CHECK_LAUNCH(); count_launch();
    have_S = true;
    mark(6);
    if(profile) { ANGPU_CUDA(cudaEventSynchronize(ev[6])); ANGPU_CUDA(cudaEventElapsedTime(&phase_ms[4], ev[5], ev[6])); }
}

// row_a[s] = O_s . v over the local samples
void TDVP::rowdot(const cplx* v_dev) {
    const size_t ns = S.ns;
    row_a.resize(std::max<size_t>(1, ns));
    if(!ns) return;
    if(factorised && use_dmma()) {
        static bool attr_set = false;
        if(!attr_set) {
            ANGPU_CUDA(cudaFuncSetAttribute(k_rowdot_dmma<8, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rd_smem(128)));
            ANGPU_CUDA(cudaFuncSetAttribute(k_rowdot_dmma<4, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rd_smem(128)));
            ANGPU_CUDA(cudaFuncSetAttribute(k_rowdot_dmma<8, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rd_smem(64)));
            ANGPU_CUDA(cudaFuncSetAttribute(k_rowdot_dmma<4, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rd_smem(64)));
            attr_set = true;
        }
        // 64 samples per block, or 32 when that would leave SMs without a block (C2: 8192 samples -> 256 blocks, not 128);
        // 128 real columns per pass for N <= 64 (C2), 64 (86 registers, more resident blocks) for wider lattices (C5: -6 %)
        const bool big = ns >= (size_t)ctx().num_sms * 2 * 64;
        if(rbm_N <= 64u) {
            if(big) k_rowdot_dmma<8, 128><<<ceil_div(ns, 64), 256, rd_smem(128), stream()>>>(S.conf.p, T.p, v_dev, ns, rbm_N, rbm_M, words, row_a.p);
            else k_rowdot_dmma<4, 128><<<ceil_div(ns, 32), 128, rd_smem(128), stream()>>>(S.conf.p, T.p, v_dev, ns, rbm_N, rbm_M, words, row_a.p);
        } else {
            if(big) k_rowdot_dmma<8, 64><<<ceil_div(ns, 64), 256, rd_smem(64), stream()>>>(S.conf.p, T.p, v_dev, ns, rbm_N, rbm_M, words, row_a.p);
            else k_rowdot_dmma<4, 64><<<ceil_div(ns, 32), 128, rd_smem(64), stream()>>>(S.conf.p, T.p, v_dev, ns, rbm_N, rbm_M, words, row_a.p);
        }
    }
    else if(factorised) k_rowdot_rbm<<<ceil_div(ns, RBM_ST), 256, 0, stream()>>>(S.conf.p, T.p, v_dev, ns, rbm_N, rbm_M, words, row_a.p);
    else k_rowdot_dense<<<(unsigned)ns, 256, 0, stream()>>>(O.p, v_dev, P, row_a.p);
    ANGPU_CHECK_LAUNCH(); count_launch();
}
// out = S v (+ (shift_abs + shift_rel diag) v when diag != null).  dot_dev: device scalar Obar . v if already known.
void TDVP::matvec(const cplx* v_dev, cplx* out_dev, const cplx* dot_dev, const double* diag, double shift_abs, double shift_rel, bool allow_S) {
    ANGPU_REQUIRE(evaluated, "TDVP: call eval / eval_F first");
    if(allow_S && have_S && !factorised) {      // S already built (eval): one pass over P^2 instead of two over Ns x P
        k_smat_vec<<<P, 256, 0, stream()>>>(Smat.p, v_dev, diag, shift_abs, shift_rel, P, out_dev);
        ANGPU_CHECK_LAUNCH(); count_launch();
        return;
    }
    // tc_products: the factorised product on the tcgen05 tensor cores (sv_tc.cu; ~1e-6 relative) instead of the exact DMMA kernels
    const bool tcp = tc_products > 0 && tc_available();       // S_dot_vector / probes: only on request (solve_cg: also by size)
    if(tcp) tc_rowdot(v_dev); else rowdot(v_dev);
    d_scal.resize(16);
    if(!dot_dev) {
        cplx* dot = reinterpret_cast<cplx*>(d_scal.p) + 4;
        if(P > 65536u) {
            vb_part.resize(3 * VB_BLOCKS);
            k_dot_part<false><<<VB_BLOCKS, VB_T, 0, stream()>>>(Ok_dev(), v_dev, P, vb_part.p + 2 * VB_BLOCKS, nullptr);
            k_sum_part_to_scalar<<<1, 32, 0, stream()>>>(vb_part.p + 2 * VB_BLOCKS, dot);
            count_launch();
        } else k_dot<false><<<1, RED_T, 0, stream()>>>(Ok_dev(), v_dev, P, dot);
        ANGPU_CHECK_LAUNCH(); count_launch();
        dot_dev = dot;
    }
    ColPartials cp;
    if(tcp) {
        // the mean Obar . v is removed from a_s before the second product (sv_tc.cu: k_pack_z): no rank-1 correction afterwards
        cplx* px = nullptr; const unsigned ch = tc_col_partials(row_a.p, dot_dev, 1u, &px); cp = ColPartials{ch, nullptr, px};
        tc_zero.resize(VB_BLOCKS); tc_zero.zero();
        dot_dev = tc_zero.p;
    }
    else cp = col_reduce_partials(*this, row_a.p, false);
    if(reduce_on()) {
        k_sum_chunks<<<grid_for(P), 256, 0, stream()>>>(cp.x, cp.chunks, P, out_dev);
        ANGPU_CHECK_LAUNCH(); count_launch();
        allreduce_sum(reinterpret_cast<double*>(out_dev), 2 * (size_t)P);
        k_sv_finish<<<grid_for(P), 256, 0, stream()>>>(out_dev, 1u, Ok_dev(), dot_dev, v_dev, diag, shift_abs, shift_rel, P, out_dev, true);
    } else {
        k_sv_finish<<<grid_for(P), 256, 0, stream()>>>(cp.x, cp.chunks, Ok_dev(), dot_dev, v_dev, diag, shift_abs, shift_rel, P, out_dev, true);
    }
    ANGPU_CHECK_LAUNCH(); count_launch();
}
int TDVP::tc_products_default() { static const int v = [] { const char* e = getenv("ANGPU_CG_TC"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }(); return v; }
// solve_cg: on request, or -- left at "auto" -- when the product is large enough for the tensor-core pipeline to win (measured: C2,
// ns N M = 1.3e8: 0.107 vs 0.096 ms per iteration, slower; C5 shard, 5.2e9: 1.00 vs 2.35 ms)
bool TDVP::tc_wanted() const { return tc_products > 0 || (tc_products < 0 && (double)S.ns * rbm_N * rbm_M >= 1e9); }
void TDVP::S_dot_vector_dev(const cplx* v_dev, cplx* out_dev) { matvec(v_dev, out_dev, nullptr, nullptr, 0.0, 0.0); }
void TDVP::S_dot_vector(const cplx* v_host, cplx* out_host) {
    set_reduce(sharded);
    vec_in.upload(v_host, P);
    vec_out.resize(P);
    S_dot_vector_dev(vec_in.p, vec_out.p);
    vec_out.download(out_host, P);
}

static void tdvp_diag(TDVP& t, DevBuf<double>& dbuf) {
    dbuf.resize(t.P);
    const size_t ns = t.S.ns;
    const unsigned period = t.factorised ? t.rbm_M : t.P;
    unsigned chunks = (unsigned)std::max<size_t>(1, std::min<size_t>(256, ns / 64));
    const size_t chunk = ns ? (ns + chunks - 1) / chunks : 1;
    chunks = ns ? (unsigned)((ns + chunk - 1) / chunk) : 1;
    DevBuf<double>& part = t.diag_part;               // grow-only member: no cudaMalloc / cudaFree / stream sync per call
    part.resize((size_t)chunks * period);
    if(ns == 0) part.zero();
    else if(t.factorised) k_diag_rbm<<<dim3(ceil_div(period, 128), chunks), 128, 0, stream()>>>(t.T.p, t.S.weight.p, ns, t.rbm_M, chunk, part.p);
    else k_diag_dense<<<dim3(ceil_div(period, 128), chunks), 128, 0, stream()>>>(t.O.p, t.S.weight.p, ns, t.P, chunk, part.p);
    k_diag_sum<<<grid_for(t.P), 256, 0, stream()>>>(part.p, chunks, period, t.P, dbuf.p);
    ANGPU_CHECK_LAUNCH(); count_launch(2);
    allreduce_sum(dbuf.p, t.P);
    k_diag_finalize<<<grid_for(t.P), 256, 0, stream()>>>(dbuf.p, t.Ok_dev(), t.P);
    ANGPU_CHECK_LAUNCH(); count_launch();
}

int TDVP::solve_cg(double tol, unsigned max_iter, double shift_abs, double shift_rel, cplx rhs_phase, cplx* x_host, double* rel_res_out) {
    ANGPU_REQUIRE(evaluated, "TDVP: call eval / eval_F first");
    set_reduce(sharded);
    // with a dense S at hand (eval) the products stream S; ANGPU_CG_MATRIX_FREE=1 forces the O-based products
    const char* env_free = getenv("ANGPU_CG_MATRIX_FREE");
    const bool force_free = env_free && env_free[0] == '1';
    const bool use_S = have_S && !factorised && !force_free;
    mark(5);
    const size_t n = P;
    // Jacobi preconditioning (M = diagonal of the shifted S) is the default; ANGPU_CG_PRECOND=none gives plain CG
    const char* env_pc = getenv("ANGPU_CG_PRECOND");
    const bool jacobi = !(env_pc && std::string(env_pc) == "none");
    DevBuf<double> dg, minv_buf;
    if(shift_rel != 0.0 || jacobi) tdvp_diag(*this, dg); else { dg.resize(n); dg.zero(); }
    const double* minv = nullptr;
    if(jacobi) {
        minv_buf.resize(n);
        k_jacobi_inverse<<<grid_for(n), 256, 0, stream()>>>(dg.p, shift_abs, shift_rel, n, minv_buf.p);
        ANGPU_CHECK_LAUNCH(); count_launch();
        minv = minv_buf.p;
    }
    cg_buf.resize(6 * n);                                     // x | r | p | Ap | b | A x (residual refresh)
    cplx *x = cg_buf.p, *r = x + n, *p = r + n, *Ap = p + n, *b = Ap + n, *Ax = b + n;
    d_scal.resize(16);
    cplx* scal = reinterpret_cast<cplx*>(d_scal.p);
    ANGPU_CUDA(cudaMemsetAsync(x, 0, sizeof(cplx) * n, stream()));
    k_scale_vec<<<grid_for(n), 256, 0, stream()>>>(F.p, rhs_phase, b, n, false);
    ANGPU_CUDA(cudaMemcpyAsync(r, b, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, stream()));
    vb_part.resize(3 * VB_BLOCKS);
    cplx* part_pAp = vb_part.p; cplx* part_rr = vb_part.p + VB_BLOCKS; cplx* part_dot = vb_part.p + 2 * VB_BLOCKS;
    k_cg_init<<<1, RED_T, 0, stream()>>>(p, r, minv, n, scal + 0, Ok_dev(), part_dot);
    ANGPU_CHECK_LAUNCH(); count_launch(2);
    // sample-based products: the S.v epilogue, p.Ap, |r|^2 and Obar.p ride in three fused vector kernels (5 launches per
    // iteration; with several ranks the per-rank column sums are all-reduced before the epilogue); products on the dense S
    // use the generic matvec + a separate dot product
    const bool fused = !use_S && (S.ns > 0 || reduce_on());
    // PsiRBM: the search-direction products on the tcgen05 tensor cores (sv_tc.cu, TF32 hi/lo planes, ~1e-6 relative), the TRUE
    // residual r = b - A x recomputed with the exact FP64-tensor-core product every 32 iterations; the last phase runs exact (below).
    bool use_tc = fused && tc_wanted() && tc_available();
    if(use_tc && tc_products < 0) {                       // auto mode never fails because of the fast path: fall back to exact products
        try { tc_prepare(); } catch(const Error&) { cudaGetLastError(); use_tc = false; }
    }
    if(reduce_on()) {
        // every rank must take the same path (the tensor-core path issues extra collectives for the residual refresh): all or none
        // (votes, ranks) summed over the ranks -- the rank count comes from the reduction itself, so this also holds for the callback
        // transport, which does not know the world size
        double* flag = d_scal.p + 14;
        const double mine[2] = {use_tc ? 1.0 : 0.0, 1.0};
        ANGPU_CUDA(cudaMemcpyAsync(flag, mine, 2 * sizeof(double), cudaMemcpyHostToDevice, stream()));
        allreduce_sum(flag, 2);
        double sum[2] = {0.0, 0.0};
        ANGPU_CUDA(cudaMemcpyAsync(sum, flag, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream()));
        ANGPU_CUDA(cudaStreamSynchronize(stream()));
        use_tc = sum[0] > sum[1] - 0.5;
    }
    // convergence is decided on values summed over ranks, so that every rank takes the same decision
    auto read_rs = [&](int slot) -> double {      // |r|^2 = the imaginary slot of the (r.z, |r|^2) pair
        double* chk = d_scal.p + 12;
        ANGPU_CUDA(cudaMemcpyAsync(chk, reinterpret_cast<double*>(scal + slot) + 1, sizeof(double), cudaMemcpyDeviceToDevice, stream()));
        allreduce_sum(chk, 1);
        double hv = 0.0;
        ANGPU_CUDA(cudaMemcpyAsync(&hv, chk, sizeof(double), cudaMemcpyDeviceToHost, stream()));
        ANGPU_CUDA(cudaStreamSynchronize(stream()));
        return hv;
    };
    cplx h(0.0, 0.0);
    // convergence is read on the host every `check_every` iterations (tensor-core products: 8, the measured-stable refresh period)
    static const unsigned check_exact = [] { const char* e = getenv("ANGPU_CG_CHECK"); const int v = e ? atoi(e) : 8; return (unsigned)std::max(1, v); }();
    const unsigned check_every = use_tc ? 8u : check_exact;
    const double b2 = read_rs(0);
    if(rel_res_out) *rel_res_out = 0.0;
    unsigned it = 0;
    // Tensor-core products (~1e-6 relative; sv_tc.cu): the residual recursion r -= alpha A p is REPLACED by r = b - A x with the exact
    // FP64-tensor-core product at every convergence check (every 8 iterations) -- measured: with a refresh every 16 or 32 iterations
    // the C2 solve diverges, every 8 it takes the iterations of the exact solve (C2 72 / 72, C5 192 / 184) -- so the residual that is
    // tested and reported is always an fp64 one.  Guard: if a refreshed residual is 10x above the best one seen, the rest of the
    // solve uses exact products.
    bool tc_phase = use_tc, refresh_now = false;
    double best_true = -1.0;
    if(use_tc) { tc_zero.resize(VB_BLOCKS); tc_zero.zero(); }
    if(b2 > 0.0) {
        for(it = 1; it <= max_iter; it++) {
            refresh_now = false;
            if(fused) {
                ColPartials cp;
                const cplx* dot_parts = part_dot;
                if(tc_phase) {
                    tc_rowdot(p);
                    cplx* px = nullptr; const unsigned ch = tc_col_partials(row_a.p, part_dot, VB_BLOCKS, &px); cp = ColPartials{ch, nullptr, px};
                    dot_parts = tc_zero.p;                    // the mean is already removed (k_pack_z)
                }
                else { rowdot(p); cp = col_reduce_partials(*this, row_a.p, false); }
                if(reduce_on()) {
                    k_sum_chunks<<<grid_for(P), 256, 0, stream()>>>(cp.x, cp.chunks, P, Ap);
                    ANGPU_CHECK_LAUNCH(); count_launch();
                    allreduce_sum(reinterpret_cast<double*>(Ap), 2 * n);
                    cp.x = Ap; cp.chunks = 1u;
                }
                k_sv_finish_cg<<<VB_BLOCKS, VB_T, 0, stream()>>>(cp.x, cp.chunks, Ok_dev(), dot_parts, p, dg.p, shift_abs, shift_rel, n, Ap,
                                                                  part_pAp, it > 1 ? scal : nullptr);
                k_cg_xr_mb<<<VB_BLOCKS, VB_T, 0, stream()>>>(x, r, p, Ap, scal, part_pAp, part_rr, minv, n);
                refresh_now = tc_phase && (it % check_every == 0 || it == max_iter);
                if(refresh_now) {
                    const int keep = tc_products; tc_products = 0;
                    matvec(x, Ax, nullptr, dg.p, shift_abs, shift_rel, false);
                    tc_products = keep;
                    k_cg_refresh<<<VB_BLOCKS, VB_T, 0, stream()>>>(r, b, Ax, minv, part_rr, n);
                    count_launch();
                }
                k_cg_p_mb<<<VB_BLOCKS, VB_T, 0, stream()>>>(p, r, scal, part_rr, minv, Ok_dev(), part_dot, n);
                ANGPU_CHECK_LAUNCH(); count_launch(3);
            } else {
                matvec(p, Ap, nullptr, dg.p, shift_abs, shift_rel, use_S);
                k_dot_part<true><<<VB_BLOCKS, VB_T, 0, stream()>>>(p, Ap, n, part_pAp, it > 1 ? scal : nullptr);
                k_cg_xr_mb<<<VB_BLOCKS, VB_T, 0, stream()>>>(x, r, p, Ap, scal, part_pAp, part_rr, minv, n);
                k_cg_p_mb<<<VB_BLOCKS, VB_T, 0, stream()>>>(p, r, scal, part_rr, minv, nullptr, nullptr, n);
                ANGPU_CHECK_LAUNCH(); count_launch(3);
            }
            if(it % check_every == 0 || it == max_iter) {
                h.re = read_rs(2);
                if(rel_res_out) *rel_res_out = std::sqrt(h.re / b2);
                if(h.re <= tol * tol * b2) break;
                if(tc_phase) {
                    if(best_true >= 0.0 && h.re > 100.0 * best_true) tc_phase = false;       // (squared norms: 10x in the residual)
                    if(best_true < 0.0 || h.re < best_true) best_true = h.re;
                }
            }
        }
        if(it > max_iter) it = max_iter;
    }
    mark(6);
    last_x = x;
    if(x_host) ANGPU_CUDA(cudaMemcpyAsync(x_host, x, sizeof(cplx) * n, cudaMemcpyDeviceToHost, stream()));
    ANGPU_CUDA(cudaStreamSynchronize(stream()));
    if(profile) ANGPU_CUDA(cudaEventElapsedTime(&phase_ms[5], ev[5], ev[6]));
    return (int)it;
}

void TDVP::solve_dense(double shift_abs, double shift_rel, cplx rhs_phase, cplx* x_host) {
    ANGPU_REQUIRE(evaluated, "TDVP: call eval first");
    set_reduce(sharded);
    if(!have_S) build_S();
    mark(5);
    const size_t n = P;
    DevBuf<double> dg;
    if(shift_rel != 0.0) tdvp_diag(*this, dg); else { dg.resize(n); dg.zero(); }
    // The factorisation works in place on a copy of S (16 P^2 bytes), a grow-only member: cudaMalloc/cudaFree of a GB-sized
    // buffer per call cost 0.1-0.8 s on some boxes.  Hand-written blocked Cholesky on the row-major upper triangle
    // (cholesky.cu): no library call on the SR path.
    DevBuf<cplx>& A = solve_A; DevBuf<cplx>& b = solve_b; DevBuf<int>& info = solve_info;
    b.resize(n); info.resize(2);
    A.copy_from(Smat);
    k_add_diag_shift<<<grid_for(n), 256, 0, stream()>>>(A.p, dg.p, shift_abs, shift_rel, P);
    k_scale_vec<<<grid_for(n), 256, 0, stream()>>>(F.p, rhs_phase, b.p, n, false);
    ANGPU_CHECK_LAUNCH(); count_launch(2);
    cholesky_solve(A.p, b.p, P, info.p, solve_work);
    mark(6);
    int hinfo = 0; info.download(&hinfo, 1);
    if(hinfo != 0) throw Error("dense solve: S + shift is not positive definite (pivot " + std::to_string(hinfo) + " is not positive); increase the diagonal shift");
    last_x = b.p;
    if(x_host) b.download(x_host, n); else ANGPU_CUDA(cudaStreamSynchronize(stream()));
    if(profile) ANGPU_CUDA(cudaEventElapsedTime(&phase_ms[5], ev[5], ev[6]));
}

// ============================================================================================ HilbertSpaceDistance
// per sample (HilbertSpaceDistance.cu.template:41-70): omega, probability ratio, next-state norm; block sums of
// w*omega (2), w*ratio, w*norm into out[4] by a single block (fixed order: deterministic)
__global__ void __launch_bounds__(RED_T) k_hsd_terms(const double* __restrict__ w, const cplx* __restrict__ lp, const cplx* __restrict__ lpp,
        const cplx* __restrict__ eloc, size_t ns, bool is_unitary, cplx* __restrict__ omega_out, cplx* __restrict__ ratio_out, double* __restrict__ out4) {
    double v[4] = {0.0, 0.0, 0.0, 0.0}, red[4];
    for(size_t s = threadIdx.x; s < ns; s += RED_T) {
        const cplx d = conj(lpp[s] - lp[s]), E = eloc[s];
        cplx om; double nsn;
        if(is_unitary) { om = cexp(d) * E; nsn = abs2(E); }
        else           { om = cexp(E + d); nsn = exp(2.0 * E.re); }
        const double r = exp(2.0 * (lpp[s].re - lp[s].re));
        omega_out[s] = om; ratio_out[s] = cplx(r, 0.0);
        v[0] += w[s] * om.re; v[1] += w[s] * om.im; v[2] += w[s] * r; v[3] += w[s] * nsn;
    }
    block_reduce<4>(v, red);
    if(threadIdx.x == 0) { out4[0] = red[0]; out4[1] = red[1]; out4[2] = red[2]; out4[3] = red[3]; }
}

void HilbertSpaceDistance::averages(Psi& psi, Psi& psi_prime, const Operator& op, bool is_unitary, Ensemble& ens, bool want_gradient, double h[5]) {
    ANGPU_REQUIRE(psi.N == psi_prime.N, "HilbertSpaceDistance: psi and psi_prime act on different numbers of sites");
    ANGPU_REQUIRE(psi_prime.P == P, "HilbertSpaceDistance: num_params differs from psi_prime's");
    ens.generate(psi, S);
    psi.eloc(op, S);
    SampleSet& Sp = rows.S;                            // the same configurations and weights, log psi' and psi' caches
    Sp.resize(S.ns, psi_prime.words);
    omega.resize(std::max<size_t>(1, S.ns)); ratio.resize(std::max<size_t>(1, S.ns));
    d_scal.resize(8);
    if(S.ns) {
        ANGPU_CUDA(cudaMemcpyAsync(Sp.conf.p, S.conf.p, sizeof(uint64_t) * S.ns * S.words, cudaMemcpyDeviceToDevice, stream()));
        ANGPU_CUDA(cudaMemcpyAsync(Sp.weight.p, S.weight.p, sizeof(double) * S.ns, cudaMemcpyDeviceToDevice, stream()));
        psi_prime.log_psi(Sp, false);
    }
    k_hsd_terms<<<1, RED_T, 0, stream()>>>(S.weight.p, S.log_psi.p, Sp.log_psi.p, S.eloc.p, S.ns, is_unitary, omega.p, ratio.p, d_scal.p);
    ANGPU_CHECK_LAUNCH(); count_launch();
    allreduce_sum(d_scal.p, 4);
    if(want_gradient) {
        g.resize(2 * (size_t)P);
        rows.evaluated = true;
        rows.prepare_rows(psi_prime, false);
        rows.weighted_conj_column_sums(omega.p, g.p);
        rows.weighted_conj_column_sums(ratio.p, g.p + P);
        allreduce_sum(reinterpret_cast<double*>(g.p), 4 * (size_t)P);
    }
    d_scal.download(h, 4);
}
double HilbertSpaceDistance::distance(Psi& psi, Psi& psi_prime, const Operator& op, bool is_unitary, Ensemble& ens) {
    double h[5]; averages(psi, psi_prime, op, is_unitary, ens, false, h);
    const double u = h[0] * h[0] + h[1] * h[1], v = h[3] * h[2];
    return std::sqrt(std::max(1.0 - u / v, 1e-8));
}
double HilbertSpaceDistance::gradient(cplx* result, Psi& psi, Psi& psi_prime, const Operator& op, bool is_unitary, Ensemble& ens, float nu) {
    double h[5]; averages(psi, psi_prime, op, is_unitary, ens, true, h);
    const cplx om(h[0], h[1]);
    const double u = abs2(om), v = h[3] * h[2];
    const double dist = std::sqrt(std::max(1.0 - u / v, 1e-8)), prefactor = std::pow(dist, (double)nu);
    std::vector<cplx> gh(2 * (size_t)P); g.download(gh.data(), gh.size());
    for(unsigned k = 0; k < P; k++) {                  // HilbertSpaceDistance.cu.template:160-168
        const cplx u_k = conj(om) * gh[k];
        const cplx v_k = h[3] * gh[P + k];
        const cplx num = u_k * v - u * v_k;
        result[k] = cplx(-num.re / (v * v) / prefactor, -num.im / (v * v) / prefactor);
    }
    return dist;
}

// ============================================================================================ KullbackLeibler
// per sample (KullbackLeibler.cu.template:41-60): weight = w' exp(2 (scale Re log psi - Re log psi')), deviation =
// log psi' - scale log psi - last_mean_deviation, masked to 0 where |deviation| <= threshold.  The weights are rewritten
// in place; out6 = {sum weight, Re/Im sum weight (log psi' - scale log psi), Re/Im sum_masked weight dev, sum_masked weight |dev|^2}
__global__ void __launch_bounds__(RED_T) k_kl_terms(double* __restrict__ w, const cplx* __restrict__ lpp, const cplx* __restrict__ lp, size_t ns,
        double scale, cplx last_md, double threshold2, cplx* __restrict__ dev_out, double* __restrict__ out6) {
    double v[6] = {0, 0, 0, 0, 0, 0}, red[6];
    for(size_t s = threadIdx.x; s < ns; s += RED_T) {
        const cplx l = scale * lp[s], diff = lpp[s] - l;
        const double wt = w[s] * exp(2.0 * (l.re - lpp[s].re));
        w[s] = wt;
        const cplx d = diff - last_md;
        const double d2 = abs2(d);
        const bool keep = d2 > threshold2;
        dev_out[s] = keep ? d : cplx(0.0, 0.0);
        v[0] += wt; v[1] += wt * diff.re; v[2] += wt * diff.im;
        if(keep) { v[3] += wt * d.re; v[4] += wt * d.im; v[5] += wt * d2; }
    }
    block_reduce<6>(v, red);
    if(threadIdx.x == 0) { for(int i = 0; i < 6; i++) out6[i] = red[i]; }
}
__global__ void k_kl_factor(const cplx* __restrict__ dev, size_t ns, int which, cplx* __restrict__ out) {     // 0: conj(dev), 1: |dev|^2
    for(size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += (size_t)gridDim.x * blockDim.x)
        out[s] = which == 0 ? conj(dev[s]) : cplx(abs2(dev[s]), 0.0);
}
// column sums with |O_sk|^2 over dense rows: out[0][k] = sum w |O|^2, out[1][k] = sum w |dev|^2 |O|^2, out[2..3][k] = Re/Im sum w dev |O|^2
__global__ void k_kl_abs2_cols(const cplx* __restrict__ O, const double* __restrict__ w, const cplx* __restrict__ dev, size_t ns, unsigned P,
                               double* __restrict__ out) {
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= P) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for(size_t s = 0; s < ns; s++) {
        const double o2 = abs2(O[s * P + k]), wt = w[s];
        const cplx d = dev[s];
        a0 += wt * o2; a1 += wt * abs2(d) * o2; a2 += wt * d.re * o2; a3 += wt * d.im * o2;
    }
    out[k] = a0; out[(size_t)P + k] = a1; out[2 * (size_t)P + k] = a2; out[3 * (size_t)P + k] = a3;
}

// mode 0: scalars only; 1: + O_k and dev conj(O_k) sums; 2: + the noise terms (dense rows)
void KullbackLeibler::averages(Psi& psi, Psi& psi_prime, Ensemble& ens, double threshold, int mode, double h[6]) {
    ANGPU_REQUIRE(psi.N == psi_prime.N, "KullbackLeibler: psi and psi_prime act on different numbers of sites");
    ANGPU_REQUIRE(psi_prime.P == P, "KullbackLeibler: num_params differs from psi_prime's");
    SampleSet& S = rows.S;
    ens.generate(psi_prime, S);                                   // samples of psi_prime: conf, log psi', w'
    Sp.resize(S.ns, psi.words);
    dev.resize(std::max<size_t>(1, S.ns));
    d_scal.resize(8);
    if(S.ns) {
        ANGPU_CUDA(cudaMemcpyAsync(Sp.conf.p, S.conf.p, sizeof(uint64_t) * S.ns * S.words, cudaMemcpyDeviceToDevice, stream()));
        psi.log_psi(Sp, false);
    }
    k_kl_terms<<<1, RED_T, 0, stream()>>>(S.weight.p, S.log_psi.p, Sp.log_psi.p, S.ns, log_psi_scale, last_mean_deviation, threshold * threshold, dev.p, d_scal.p);
    ANGPU_CHECK_LAUNCH(); count_launch();
    allreduce_sum(d_scal.p, 6);
    if(mode >= 1) {
        g.resize(5 * (size_t)P);                                  // O | dev conj(O) | dev O | |dev|^2 O  (sums over samples)
        rows.evaluated = true;
        rows.prepare_rows(psi_prime, mode == 2);
        col_reduce(rows, dev.p, g.p, g.p + P);                    // mean = sum w O, x = sum w dev conj(O)
        size_t nsum = 2 * (size_t)P;
        if(mode == 2) {
            aux.resize(std::max<size_t>(1, S.ns));
            k_kl_factor<<<grid_for(S.ns), 256, 0, stream()>>>(dev.p, S.ns, 0, aux.p);
            col_reduce(rows, aux.p, nullptr, g.p + 2 * P);        // sum w conj(dev) conj(O) = conj(sum w dev O)
            k_kl_factor<<<grid_for(S.ns), 256, 0, stream()>>>(dev.p, S.ns, 1, aux.p);
            col_reduce(rows, aux.p, nullptr, g.p + 3 * P);        // sum w |dev|^2 conj(O) = conj(sum w |dev|^2 O)
            gabs.resize(4 * (size_t)P);
            k_kl_abs2_cols<<<ceil_div(P, 128), 128, 0, stream()>>>(rows.O.p, S.weight.p, dev.p, S.ns, P, gabs.p);
            ANGPU_CHECK_LAUNCH(); count_launch(3);
            allreduce_sum(gabs.p, 4 * (size_t)P);
            nsum = 4 * (size_t)P;
        }
        allreduce_sum(reinterpret_cast<double*>(g.p), 2 * nsum);
    }
    d_scal.download(h, 6);
    total_weight = h[0];
    mean_deviation = cplx(h[1] / h[0], h[2] / h[0]);              // update_last_mean_deviation (:178-184)
    last_mean_deviation = mean_deviation;
}
static double kl_value(const double h[6]) {
    const double d_re = h[3] / h[0], d_im = h[4] / h[0], d2 = h[5] / h[0];
    return std::sqrt(std::max(1e-8, d2 - (d_re * d_re + d_im * d_im)));
}
double KullbackLeibler::value(Psi& psi, Psi& psi_prime, Ensemble& ens, double threshold) {
    double h[6]; averages(psi, psi_prime, ens, threshold, 0, h);
    return kl_value(h);
}
double KullbackLeibler::gradient(cplx* result, Psi& psi, Psi& psi_prime, Ensemble& ens, double nu, double threshold) {
    double h[6]; averages(psi, psi_prime, ens, threshold, 1, h);
    const double v = kl_value(h), f = std::pow(v, nu), tw = h[0];
    const cplx d(h[3] / tw, h[4] / tw);
    std::vector<cplx> gh(2 * (size_t)P); g.download(gh.data(), gh.size());
    for(unsigned k = 0; k < P; k++) {                              // :232-236
        const cplx O = (1.0 / tw) * gh[k], dOc = (1.0 / tw) * gh[P + k];
        const cplx r = dOc - d * conj(O);
        result[k] = cplx(r.re / f, r.im / f);
    }
    return v;
}
double KullbackLeibler::gradient_with_noise(cplx* result, double* noise, Psi& psi, Psi& psi_prime, Ensemble& ens, double nu, double threshold) {
    double h[6]; averages(psi, psi_prime, ens, threshold, 2, h);
    const double v = kl_value(h), f = std::pow(v, nu), tw = h[0], d2 = h[5] / tw;
    const cplx d(h[3] / tw, h[4] / tw);
    std::vector<cplx> gh(4 * (size_t)P); g.download(gh.data(), gh.size());
    std::vector<double> ga(4 * (size_t)P); gabs.download(ga.data(), ga.size());
    const double steps = (double)ens.num_steps();
    for(unsigned k = 0; k < P; k++) {                              // :286-312
        const cplx O = (1.0 / tw) * gh[k], dOc = (1.0 / tw) * gh[P + k];
        const cplx dO = conj((1.0 / tw) * gh[2 * (size_t)P + k]), d2O = conj((1.0 / tw) * gh[3 * (size_t)P + k]);
        const double O2 = ga[k] / tw, d2O2 = ga[(size_t)P + k] / tw;
        const cplx dO2(ga[2 * (size_t)P + k] / tw, ga[3 * (size_t)P + k] / tw);
        const cplx r = dOc - d * conj(O);
        result[k] = cplx(r.re / f, r.im / f);
        const cplx mix = dO * conj(d) * conj(O) + 2.0 * (conj(dOc) * d * conj(O)) - d2O * conj(O) - dO2 * conj(d);
        const double var = d2O2 - abs2(dOc) + 2.0 * mix.re + d2 * abs2(O) + abs2(d) * O2 - 4.0 * abs2(d) * abs2(O);
        noise[k] = std::sqrt(var / steps) / f;
    }
    return v;
}

// ============================================================================================ FP64 peak probe
// Dependent-free DFMA streams: 8 independent accumulators per thread, 64k FMAs each. Used by bench.py as the measured
// denominator for the FP64-pipe-bound kernels (MEASURED_PEAKS.json has no fp64 figure).
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, double a, double b, int iters) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for(int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
double measure_fp64_tflops() {
    const int blocks = ctx().num_sms * 8, threads = 256, iters = 8192;
    DevBuf<double> out((size_t)blocks * threads);
    cudaEvent_t e0, e1;
    ANGPU_CUDA(cudaEventCreate(&e0)); ANGPU_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for(int rep = 0; rep < 5; rep++) {
        ANGPU_CUDA(cudaEventRecord(e0, stream()));
        k_fp64_peak<<<blocks, threads, 0, stream()>>>(out.p, 0.999999, 1e-7, iters);
        ANGPU_CUDA(cudaEventRecord(e1, stream()));
        ANGPU_CUDA(cudaEventSynchronize(e1));
        float ms; ANGPU_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if(rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
    return flops / (best * 1e-3) / 1e12;
}

// ============================================================================================ free functions

void scalar_sums_eloc(const SampleSet& S, double* out4_dev) { scalar_sums(S, true, false, out4_dev, nullptr); }
void enumerate_probe(uint64_t index, unsigned words, uint64_t* host_out) {
    DevBuf<uint64_t> d(words);
    k_enumerate<<<1, 32, 0, stream()>>>(d.p, (size_t)index, 1, words);
    ANGPU_CHECK_LAUNCH(); count_launch();
    d.download(host_out, words);
}

cplx log_psi_s(Psi& psi, const uint64_t* conf) {
    SampleSet S; S.resize(1, psi.words);
    S.conf.upload(conf, psi.words);
    psi.log_psi(S, false);
    cplx r; S.log_psi.download(&r, 1);
    return r;
}
void psi_O_k(Psi& psi, const uint64_t* conf, cplx* out_host) {
    SampleSet S; S.resize(1, psi.words);
    S.conf.upload(conf, psi.words);
    DevBuf<cplx> row(psi.P);
    psi.ok_rows(S, 0, 1, row.p);
    row.download(out_host, psi.P);
}
void log_psi_vector(Psi& psi, Ensemble& ens, cplx* out_host, bool exponentiate) {
    SampleSet S;
    ens.generate(psi, S);
    if(exponentiate && S.ns) { k_exp_inplace<<<grid_for(S.ns), 256, 0, stream()>>>(S.log_psi.p, S.ns); ANGPU_CHECK_LAUNCH(); count_launch(); }
    S.log_psi.download(out_host, S.ns);
}
cplx log_psi_mean(Psi& psi, Ensemble& ens) {
    SampleSet S;
    ens.generate(psi, S);
    DevBuf<double> d(8);
    scalar_sums(S, false, true, nullptr, d.p);
    allreduce_sum(d.p, 2);
    double h[2]; d.download(h, 2);
    return cplx(h[0], h[1]);
}
double psi_norm(Psi& psi, Ensemble& es) {
    ANGPU_REQUIRE(!es.is_mc, "psi_norm needs an ExactSummation ensemble");
    SampleSet S;
    es.generate(psi, S);
    DevBuf<double> d(8);
    scalar_sums(S, false, false, d.p, nullptr);
    allreduce_sum(d.p, 4);
    double h[4]; d.download(h, 4);
    return std::sqrt(h[3]);
}
void psi_O_k_vector(Psi& psi, Ensemble& es, cplx* out_host) {
    SampleSet S;
    es.generate(psi, S);
    DevBuf<cplx> O(std::max<size_t>(1, S.ns * (size_t)psi.P)), out(psi.P);
    psi.ok_rows(S, 0, S.ns, O.p);
    k_sum_rows<<<ceil_div(psi.P, 128), 128, 0, stream()>>>(O.p, S.ns, psi.P, out.p);
    ANGPU_CHECK_LAUNCH(); count_launch();
    allreduce_sum(reinterpret_cast<double*>(out.p), 2 * (size_t)psi.P);
    out.download(out_host, psi.P);
}
void apply_operator(Psi& psi, const Operator& op, Ensemble& ens, cplx* out_host) {
    SampleSet S;
    ens.generate(psi, S);
    psi.eloc(op, S);
    DevBuf<cplx> out(std::max<size_t>(1, S.ns));
    if(S.ns) { k_mul_exp<<<grid_for(S.ns), 256, 0, stream()>>>(S.log_psi.p, S.eloc.p, out.p, S.ns); ANGPU_CHECK_LAUNCH(); count_launch(); }
    out.download(out_host, S.ns);
}
void local_energies(Psi& psi, const Operator& op, const uint64_t* confs_host, size_t ns, cplx* log_psi_out, cplx* eloc_out) {
    SampleSet S; S.resize(ns, psi.words);
    S.pauli_sites = psi.pauli_sites;
    S.conf.upload(confs_host, ns * psi.words);
    psi.log_psi(S, false);
    psi.eloc(op, S);
    if(log_psi_out) S.log_psi.download(log_psi_out, ns);
    if(eloc_out) S.eloc.download(eloc_out, ns);
}

} // namespace angpu

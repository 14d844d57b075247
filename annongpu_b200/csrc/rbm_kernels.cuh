// PsiRBM fast paths (the BASELINE.json headline path: MC sampling -> E_loc -> O_k for an RBM).
//
//  k_mc_rbm    one warp per Markov chain; the M hidden-unit angles live in REGISTERS (K = ceil(M/32) complex
//              per lane), a spin flip costs one coalesced W-row read (L1/L2 resident) + K polynomial
//              evaluations per lane + ONE warp-shuffle reduction; no shared memory, no barriers, no atomics in
//              the loop (the reference: >= 8 __syncthreads + an M-way shared-atomic reduce + a global atomic per
//              proposal, include/ensembles/MonteCarlo.hpp:135-177, PsiRBM.hpp:92-157).
//              Bound: FP64 pipe (~16*M DFMA per proposal) — see DESIGN.md; compulsory HBM traffic is ~0.
//  k_mc_rbm_block  the same chain on one 256-thread block for 512 < M <= 2048 (one barrier per proposal).
//  k_eloc_rbm  a team of 1..8 warps per sample, FOUR LANES PER FLIP GROUP: each quad walks the M hidden units for its
//              own s', reading theta_j from the team's shared memory and W[site][j] as 64 contiguous bytes, and
//              combines with two shuffles.  Groups whose coefficient is exactly zero on s (XX+YY on aligned
//              spins) are compacted away first.
//  k_rbm_T     the factorised log-derivative: O[s][i*M+j] = sigma_si * T[s][j], T = fw * th0(theta)
//              (PsiRBM.hpp:161-176) — ns*M complex instead of ns*N*M.
//  k_rbm_dense_O  materialises the dense rows only when a caller asks for O_k_samples / a dense S.
#pragma once
#include "kernels.cuh"
#include "dmma.cuh"

namespace angpu {

#ifdef __CUDACC__

__device__ __forceinline__ double conf_spin(const uint64_t (&c)[MAXW], unsigned site) {
    const unsigned w = site >> 6;
    uint64_t v = c[0];
    if(w == 1u) v = c[1];
    if(w == 2u) v = c[2];
    if(w == 3u) v = c[3];
    return ((v >> (site & 63u)) & 1ull) ? 1.0 : -1.0;
}
__device__ __forceinline__ void conf_load(uint64_t (&c)[MAXW], const uint64_t* __restrict__ src, unsigned words) {
    #pragma unroll
    for(unsigned w = 0; w < (unsigned)MAXW; w++) c[w] = (w < words) ? src[w] : 0ull;
}
// register-resident variants of operator.hpp's string_sign / strings_coefficient
__device__ __forceinline__ double string_sign_reg(const OpDev& op, unsigned n, const uint64_t (&c)[MAXW]) {
    unsigned pc = 0;
    #pragma unroll
    for(unsigned w = 0; w < (unsigned)MAXW; w++) if(w < op.words) pc += __popcll(~c[w] & op.b[n * op.words + w]);
    return (pc & 1u) ? -1.0 : 1.0;
}
__device__ __forceinline__ cplx strings_coefficient_reg(const OpDev& op, unsigned lo, unsigned hi, const uint64_t (&c)[MAXW]) {
    cplx r(0.0, 0.0);
    for(unsigned n = lo; n < hi; n++) r += string_sign_reg(op, n, c) * op.coef[n];
    return r;
}
__device__ __forceinline__ void conf_flip(uint64_t (&c)[MAXW], unsigned site) {
    const unsigned w = site >> 6;
    const uint64_t m = 1ull << (site & 63u);
    if(w == 0u) c[0] ^= m;
    if(w == 1u) c[1] ^= m;
    if(w == 2u) c[2] ^= m;
    if(w == 3u) c[3] ^= m;
}

// Re and Im of lc0(z) = z^2/2 - z^4/12 + z^6/45
__device__ __forceinline__ cplx lc0(cplx z) {
    const double x2 = fma(z.re, z.re, -z.im * z.im), y2 = (z.re + z.re) * z.im;
    const double x4 = fma(x2, x2, -y2 * y2), y4 = (x2 + x2) * y2;
    const double x6 = fma(x4, x2, -y4 * y2), y6 = fma(x4, y2, y4 * x2);
    return cplx(fma(1.0 / 45.0, x6, fma(-1.0 / 12.0, x4, 0.5 * x2)),
                fma(1.0 / 45.0, y6, fma(-1.0 / 12.0, y4, 0.5 * y2)));
}
__device__ __forceinline__ double lc0_re(cplx z) {
    const double x2 = fma(z.re, z.re, -z.im * z.im), y2 = (z.re + z.re) * z.im;
    const double x4 = fma(x2, x2, -y2 * y2), y4 = (x2 + x2) * y2;
    const double x6 = fma(x4, x2, -y4 * y2);
    return fma(1.0 / 45.0, x6, fma(-1.0 / 12.0, x4, 0.5 * x2));
}

// 2 warps per block; for K <= 8 (M <= 256, the C2 shape) 9 blocks = 18 warps per SM are requested (<= 112 registers)
constexpr int MC_RBM_THREADS = 64;
#ifndef MC_RBM_MINB
#define MC_RBM_MINB 8
#endif

// Re / Im of lc0(z) in the variables p = Re z^2 = x^2 - y^2, q = x y (Im z^2 = 2 q):
//   Re lc0 = p (1/2 - p/12 + p^2/45) + q^2 (1/3 - 12 p/45)          9 DP instructions from (x, y)
//   Im lc0 = q (1 - p/3 + 2 p^2/15 - 8 q^2/45)
// (same polynomial as lc0(); fewer operations for the sampler's inner loop, rounding differs at the 1e-16 level)
__device__ __forceinline__ double lc0_re_pq(double x, double y) {
    const double p = fma(x, x, -(y * y)), q = x * y, q2 = q * q;
    double A = fma(p, 1.0 / 45.0, -1.0 / 12.0);
    A = fma(A, p, 0.5);
    const double B = fma(p, -12.0 / 45.0, 1.0 / 3.0);
    return fma(q2, B, A * p);
}
// acc += Re lc0(x + i y) with the final product and the accumulation folded into two FMAs (one DMUL and one DADD fewer
// per hidden unit than `acc += lc0_re_pq(x, y)`: 10 instead of 12 FP64 instructions incl. the angle update)
__device__ __forceinline__ void lc0_re_pq_acc(double x, double y, double& acc) {
    const double p = fma(x, x, -(y * y)), q = x * y, q2 = q * q;
    double A = fma(p, 1.0 / 45.0, -1.0 / 12.0);
    A = fma(A, p, 0.5);
    const double B = fma(p, -12.0 / 45.0, 1.0 / 3.0);
    acc = fma(q2, B, acc);
    acc = fma(A, p, acc);
}
__device__ __forceinline__ cplx lc0_pq(double x, double y) {
    const double p = fma(x, x, -(y * y)), q = x * y, q2 = q * q;
    double A = fma(p, 1.0 / 45.0, -1.0 / 12.0);
    A = fma(A, p, 0.5);
    const double B = fma(p, -12.0 / 45.0, 1.0 / 3.0);
    double C = fma(p, 2.0 / 15.0, -1.0 / 3.0);
    C = fma(C, p, 1.0);
    C = fma(q2, -8.0 / 45.0, C);
    return cplx(fma(q2, B, A * p), q * C);
}

// Metropolis acceptance "ratio > 1 || u <= ratio", ratio = exp(d2) (include/ensembles/MonteCarlo.hpp:158-160), decided
// without a double-precision exp in all but ~1e-4 of the proposals: an fp32 exp screens the comparison and only
// the band where fp32 cannot decide falls back to the exact evaluation, so the decision is ALWAYS the fp64 one.
__device__ __forceinline__ bool metropolis_accept(double d2, double u) {
    if(d2 >= 0.0) return true;
    const float rf = __expf((float)d2), uf = (float)u;
    const float band = rf * 1e-4f + 1e-37f;
    if(uf < rf - band) return true;
    if(uf > rf + band) return false;
    return u <= exp(d2);
}

// spin / flip on a register-resident configuration of a compile-time number of words
template<int WORDS>
__device__ __forceinline__ double conf_spin_t(const uint64_t (&c)[MAXW], unsigned site) {
    if(WORDS == 1) return ((c[0] >> site) & 1ull) ? 1.0 : -1.0;
    return conf_spin(c, site);
}
template<int WORDS>
__device__ __forceinline__ void conf_flip_t(uint64_t (&c)[MAXW], unsigned site) {
    if(WORDS == 1) c[0] ^= 1ull << site;
    else conf_flip(c, site);
}

// Initial configurations of the chains (Init_Policy: a random bitmask, policies/Init_Policy.hpp:16-25) into the slots of
// their first recorded sample; the batched angle GEMM (k_rbm_angles_dmma) then leaves theta_0 = sigma_0 W in the matching
// angle slots, and the samplers start from there instead of accumulating N rows per chain themselves.
__global__ void k_mc_init_conf(const McParams mc, unsigned N, unsigned words, uint64_t* __restrict__ conf_out) {
    const unsigned chain = blockIdx.x * blockDim.x + threadIdx.x;
    if(chain >= mc.num_chains_local) return;
    const unsigned tag_init = (mc.call << 1) | 0u;
    uint32_t r[4];
    for(unsigned w = 0; w < words; w++) {
        philox4x32_10(w, 0u, mc.chain0 + chain, tag_init, mc.seed_lo, mc.seed_hi, r);
        uint64_t c = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
        if(w == words - 1u && (N & 63u)) c &= (1ull << (N & 63u)) - 1ull;
        conf_out[(size_t)chain * words + w] = c;
    }
}

// Wp: W with rows padded to Mp = 32*K complex (zeros beyond M) so that the hot loop has no bounds checks and one
// base address per proposal: lane l reads Wp[site][l + 32k], k < K (coalesced 512 B per k).
template<int K, int WORDS, bool FW_REAL, int MINB>
__global__ void __launch_bounds__(MC_RBM_THREADS, MINB)
k_mc_rbm(const RbmDev psi, const cplx* __restrict__ Wp, const McParams mc, uint64_t* __restrict__ conf_out,
         cplx* __restrict__ log_psi_out, cplx* __restrict__ angles_out, unsigned long long* __restrict__ acc_rej) {
    constexpr unsigned Mp = 32u * K;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if(chain >= mc.num_chains_local) return;
    const unsigned gchain = mc.chain0 + chain;
    const unsigned M = psi.M, N = psi.N;
    const unsigned tag_init = (mc.call << 1) | 0u, tag_step = (mc.call << 1) | 1u;
    const cplx* __restrict__ Wl = Wp + lane;

    uint32_t r[4];
    uint64_t conf[MAXW] = {0ull, 0ull, 0ull, 0ull};
    // units beyond M hold theta = 0 (Wp is zero-padded), for which lc0 = 0
    cplx th[K];
    if(angles_out) {
        // the initial configuration and theta_0 = sigma_0 W were left in this chain's first sample slot by k_mc_init_conf
        // and the tensor-core angle GEMM
        #pragma unroll
        for(int w = 0; w < WORDS; w++) conf[w] = conf_out[(size_t)chain * WORDS + w];
        #pragma unroll
        for(int k = 0; k < K; k++) { const unsigned j = lane + 32u * k; th[k] = (j < M) ? angles_out[(size_t)chain * M + j] : cplx(0.0, 0.0); }
    } else {
        #pragma unroll
        for(int w = 0; w < WORDS; w++) {
            philox4x32_10((uint32_t)w, 0u, gchain, tag_init, mc.seed_lo, mc.seed_hi, r);
            conf[w] = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
        }
        if(N & 63u) conf[WORDS - 1] &= (1ull << (N & 63u)) - 1ull;
        // (the host dispatches WORDS == words_for(N), so the last word is the one to mask)
        #pragma unroll
        for(int k = 0; k < K; k++) th[k] = cplx(0.0, 0.0);
        for(unsigned i = 0; i < N; i++) {
            const double s = conf_spin_t<WORDS>(conf, i);
            #pragma unroll
            for(int k = 0; k < K; k++) th[k] += s * ldg(&Wl[(size_t)i * Mp + 32u * k]);
        }
    }
    auto re_log_psi = [&]() -> double {
        if(FW_REAL) {
            double p0 = 0.0, p1 = 0.0;                       // two accumulation chains
            #pragma unroll
            for(int k = 0; k < K; k++) { if(k & 1) lc0_re_pq_acc(th[k].re, th[k].im, p1); else lc0_re_pq_acc(th[k].re, th[k].im, p0); }
            return fma(psi.fw.re, warp_sum(p0 + p1), psi.lp.re);
        } else {
            cplx p(0.0, 0.0);
            #pragma unroll
            for(int k = 0; k < K; k++) p += lc0_pq(th[k].re, th[k].im);
            p = warp_sum(p);
            return psi.lp.re + psi.fw.re * p.re - psi.fw.im * p.im;
        }
    };
    double cur_re = re_log_psi();

    const unsigned therm = mc.num_therm * N, per_sample = mc.num_sweeps * N;
    const unsigned long long total_steps = (unsigned long long)therm + (unsigned long long)per_sample * mc.steps_per_chain;
    unsigned long long acc = 0;
    unsigned long long next_record = (unsigned long long)therm + per_sample;
    unsigned sample = 0;

    for(unsigned long long t0 = 0; t0 < total_steps; t0 += 32u) {
        // one Philox block per lane: lane l draws the random numbers of proposal t0 + l (amortises the generator 32x)
        philox4x32_10((uint32_t)(t0 + lane), (uint32_t)((t0 + lane) >> 32), gchain, tag_step, mc.seed_lo, mc.seed_hi, r);
        const unsigned my_site = r[0] % N;
        const unsigned my_ulo = r[1], my_uhi = r[2];
        const unsigned nb = (unsigned)min((unsigned long long)32u, total_steps - t0);
        for(unsigned b = 0; b < nb; b++) {
            const unsigned site = __shfl_sync(FULL, my_site, b);
            const cplx* __restrict__ row = Wl + (size_t)site * Mp;
            cplx w[K];
            #pragma unroll
            for(int k = 0; k < K; k++) w[k] = ldg(&row[32u * k]);
            const double u = u01_from_bits(__shfl_sync(FULL, my_ulo, b), __shfl_sync(FULL, my_uhi, b));
            const double delta = -2.0 * conf_spin_t<WORDS>(conf, site);           // s'_p - s_p
            #pragma unroll
            for(int k = 0; k < K; k++) { th[k].re = fma(delta, w[k].re, th[k].re); th[k].im = fma(delta, w[k].im, th[k].im); }
            const double new_re = re_log_psi();
            if(metropolis_accept(2.0 * (new_re - cur_re), u)) {
                cur_re = new_re;
                conf_flip_t<WORDS>(conf, site);
                acc++;
            } else {
                // undo, as the reference does on rejection (update_input_units(next -> current), MonteCarlo.hpp:173-175).
                // The row is read again (an L1 hit) instead of being kept in 4 K registers across the evaluation above.
                #pragma unroll
                for(int k = 0; k < K; k++) {
                    const cplx wk = ldg(&row[32u * k]);
                    th[k].re = fma(-delta, wk.re, th[k].re); th[k].im = fma(-delta, wk.im, th[k].im);
                }
            }
            if(t0 + b + 1u == next_record) {
                const size_t idx = (size_t)sample * mc.num_chains_local + chain;
                cplx p(0.0, 0.0);
                #pragma unroll
                for(int k = 0; k < K; k++) {
                    const unsigned j = lane + 32u * k;
                    if(j < M) { p += lc0(th[k]); if(angles_out) angles_out[idx * M + j] = th[k]; }
                }
                p = warp_sum(p);
                if(lane == 0) {
                    log_psi_out[idx] = psi.lp + psi.fw * p;
                    #pragma unroll
                    for(int ww = 0; ww < WORDS; ww++) conf_out[idx * WORDS + ww] = conf[ww];
                }
                sample++; next_record += per_sample;
            }
        }
    }
    if(lane == 0) { atomicAdd(&acc_rej[0], acc); atomicAdd(&acc_rej[1], total_steps - acc); }
}

// Block-per-chain variant for wide RBMs (512 < M <= 2048, e.g. BASELINE C5: M = 1600): the angles still live in
// registers, K = ceil(M/256) per thread of a 256-thread block; one __syncthreads per proposal (the 8 warp partial sums
// are exchanged through a double-buffered shared-memory slot and every thread adds them in the same order, so the
// accept/reject decision is uniform without a second barrier).  Wp rows are padded to 256*K.
constexpr int MC_BLOCK_T = 256;
template<int K, bool FW_REAL>
__global__ void __launch_bounds__(MC_BLOCK_T, 2)
k_mc_rbm_block(const RbmDev psi, const cplx* __restrict__ Wp, const McParams mc, uint64_t* __restrict__ conf_out,
               cplx* __restrict__ log_psi_out, cplx* __restrict__ angles_out, unsigned long long* __restrict__ acc_rej) {
    constexpr unsigned Mp = (unsigned)MC_BLOCK_T * K;
    __shared__ cplx part[2][MC_BLOCK_T / 32];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned chain = blockIdx.x;
    const unsigned gchain = mc.chain0 + chain;
    const unsigned M = psi.M, N = psi.N;
    const unsigned tag_init = (mc.call << 1) | 0u, tag_step = (mc.call << 1) | 1u;
    const cplx* __restrict__ Wl = Wp + tid;

    uint32_t r[4];
    uint64_t conf[MAXW] = {0ull, 0ull, 0ull, 0ull};
    cplx th[K];
    if(angles_out) {
        // initial configuration and theta_0 from this chain's first sample slot (k_mc_init_conf + the angle GEMM)
        #pragma unroll
        for(unsigned w = 0; w < (unsigned)MAXW; w++) if(w < psi.words) conf[w] = conf_out[(size_t)chain * psi.words + w];
        #pragma unroll
        for(int k = 0; k < K; k++) { const unsigned j = tid + (unsigned)MC_BLOCK_T * k; th[k] = (j < M) ? angles_out[(size_t)chain * M + j] : cplx(0.0, 0.0); }
        __syncthreads();                                    // every thread has read the slot before thread 0 may overwrite it
    } else {
        #pragma unroll
        for(unsigned w = 0; w < (unsigned)MAXW; w++) {
            if(w < psi.words) {
                philox4x32_10(w, 0u, gchain, tag_init, mc.seed_lo, mc.seed_hi, r);
                conf[w] = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
                if(w == psi.words - 1u && (N & 63u)) conf[w] &= (1ull << (N & 63u)) - 1ull;
            }
        }
        #pragma unroll
        for(int k = 0; k < K; k++) th[k] = cplx(0.0, 0.0);
        for(unsigned i = 0; i < N; i++) {
            const double s = conf_spin(conf, i);
            #pragma unroll
            for(int k = 0; k < K; k++) th[k] += s * ldg(&Wl[(size_t)i * Mp + (unsigned)MC_BLOCK_T * k]);
        }
    }
    unsigned buf = 0;
    // block-wide sum of one complex per thread, identical on every thread (fixed order), ONE barrier
    auto block_sum = [&](cplx v) -> cplx {
        v = warp_sum(v);
        if(lane == 0) part[buf][warp] = v;
        __syncthreads();
        cplx t(0.0, 0.0);
        #pragma unroll
        for(int q = 0; q < MC_BLOCK_T / 32; q++) t += part[buf][q];
        buf ^= 1u;
        return t;
    };
    auto re_log_psi = [&]() -> double {
        cplx p(0.0, 0.0);
        if(FW_REAL) {
            double p1 = 0.0;
            #pragma unroll
            for(int k = 0; k < K; k++) { if(k & 1) lc0_re_pq_acc(th[k].re, th[k].im, p1); else lc0_re_pq_acc(th[k].re, th[k].im, p.re); }
            p.re += p1;
        } else {
            #pragma unroll
            for(int k = 0; k < K; k++) p += lc0_pq(th[k].re, th[k].im);
        }
        p = block_sum(p);
        return psi.lp.re + psi.fw.re * p.re - psi.fw.im * p.im;
    };
    double cur_re = re_log_psi();

    const unsigned therm = mc.num_therm * N, per_sample = mc.num_sweeps * N;
    const unsigned long long total_steps = (unsigned long long)therm + (unsigned long long)per_sample * mc.steps_per_chain;
    unsigned long long acc = 0;
    unsigned long long next_record = (unsigned long long)therm + per_sample;
    unsigned sample = 0;

    for(unsigned long long t0 = 0; t0 < total_steps; t0 += 32u) {
        philox4x32_10((uint32_t)(t0 + lane), (uint32_t)((t0 + lane) >> 32), gchain, tag_step, mc.seed_lo, mc.seed_hi, r);
        const unsigned my_site = r[0] % N;
        const unsigned my_ulo = r[1], my_uhi = r[2];
        const unsigned nb = (unsigned)min((unsigned long long)32u, total_steps - t0);
        // the W row of the first proposal of this batch
        cplx wn[K];
        {
            const unsigned site0 = __shfl_sync(FULL, my_site, 0);
            #pragma unroll
            for(int k = 0; k < K; k++) wn[k] = ldg(&Wl[(size_t)site0 * Mp + (unsigned)MC_BLOCK_T * k]);
        }
        for(unsigned b = 0; b < nb; b++) {
            const unsigned site = __shfl_sync(FULL, my_site, b);
            const double u = u01_from_bits(__shfl_sync(FULL, my_ulo, b), __shfl_sync(FULL, my_uhi, b));
            const double delta = -2.0 * conf_spin(conf, site);
            cplx w[K];
            #pragma unroll
            for(int k = 0; k < K; k++) { w[k] = wn[k]; th[k].re = fma(delta, w[k].re, th[k].re); th[k].im = fma(delta, w[k].im, th[k].im); }
            if(b + 1u < nb) {                                   // prefetch the next proposal's row (L2 latency)
                const unsigned site_n = __shfl_sync(FULL, my_site, b + 1u);
                #pragma unroll
                for(int k = 0; k < K; k++) wn[k] = ldg(&Wl[(size_t)site_n * Mp + (unsigned)MC_BLOCK_T * k]);
            }
            const double new_re = re_log_psi();
            if(metropolis_accept(2.0 * (new_re - cur_re), u)) {
                cur_re = new_re;
                conf_flip(conf, site);
                acc++;
            } else {
                #pragma unroll
                for(int k = 0; k < K; k++) { th[k].re = fma(-delta, w[k].re, th[k].re); th[k].im = fma(-delta, w[k].im, th[k].im); }
            }
            if(t0 + b + 1u == next_record) {
                const size_t idx = (size_t)sample * mc.num_chains_local + chain;
                cplx p(0.0, 0.0);
                #pragma unroll
                for(int k = 0; k < K; k++) {
                    const unsigned j = tid + (unsigned)MC_BLOCK_T * k;
                    if(j < M) { p += lc0(th[k]); if(angles_out) angles_out[idx * M + j] = th[k]; }
                }
                p = block_sum(p);
                if(tid == 0) {
                    log_psi_out[idx] = psi.lp + psi.fw * p;
                    #pragma unroll
                    for(unsigned ww = 0; ww < (unsigned)MAXW; ww++) if(ww < psi.words) conf_out[idx * psi.words + ww] = conf[ww];
                }
                sample++; next_record += per_sample;
            }
        }
    }
    if(tid == 0) { atomicAdd(&acc_rej[0], acc); atomicAdd(&acc_rej[1], total_steps - acc); }
}

// theta = W^T s and log psi for given configurations (ExactSummation / probes); one warp per configuration.
__global__ void k_rbm_angles(const RbmDev psi, const uint64_t* __restrict__ confs, size_t ns,
                             cplx* __restrict__ angles_out, cplx* __restrict__ log_psi_out, double* __restrict__ weight_out) {
    const unsigned lane = threadIdx.x & 31u, wpb = blockDim.x >> 5;
    const unsigned M = psi.M, N = psi.N;
    for(size_t s = (size_t)blockIdx.x * wpb + (threadIdx.x >> 5); s < ns; s += (size_t)gridDim.x * wpb) {
        uint64_t conf[MAXW];
        conf_load(conf, confs + s * psi.words, psi.words);
        cplx p(0.0, 0.0);
        for(unsigned j = lane; j < M; j += 32u) {
            cplx a(0.0, 0.0);
            for(unsigned i = 0; i < N; i++) a += conf_spin(conf, i) * ldg(&psi.W[i * M + j]);
            if(angles_out) angles_out[s * M + j] = a;
            p += lc0(a);
        }
        p = warp_sum(p);
        if(lane == 0) {
            const cplx lp = psi.lp + psi.fw * p;
            if(log_psi_out) log_psi_out[s] = lp;
            if(weight_out) weight_out[s] = exp(2.0 * lp.re);
        }
    }
}

// The batched initial-angle GEMM on the FP64 tensor cores:  theta [ns][M] = sigma [ns][N] . W [N][M]  (PsiRBM.hpp:71-80 for
// every configuration at once), W viewed as a real [N][2M] matrix ((re, im) adjacent) so that the +-1 operand stays real and
// each lane of an m8n8k4 tile ends up with one complex angle.  8 samples per warp, AG_CB real columns per pass, W streamed
// through a cp.async double buffer of 32 sites; the +-1 operand comes from the configuration bits in registers.
// Epilogue per sample: angles (optional), log psi = lp + fw sum_j lc0(theta_j), ExactSummation weight exp(2 Re log psi).
// Products with +-1 are exact and the accumulation is fp64, so the result differs from the sequential sum only by the
// order of the N additions (<= N ulp).
constexpr int AG_KC = 32, AG_PAD = 8;
constexpr size_t ag_smem(int cb) { return 2 * (size_t)AG_KC * (cb + AG_PAD) * sizeof(double); }
template<int RW, int AG_CB>
__global__ void __launch_bounds__(RW * 32) k_rbm_angles_dmma(const RbmDev psi, const uint64_t* __restrict__ confs, size_t ns,
        cplx* __restrict__ angles_out, cplx* __restrict__ log_psi_out, double* __restrict__ weight_out) {
    extern __shared__ __align__(16) double ag_buf[];
    constexpr int STRIDE = AG_CB + AG_PAD, STAGE = AG_KC * STRIDE;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, row = lane >> 2, kq = lane & 3u;
    const unsigned N = psi.N, M = psi.M, words = psi.words;
    const size_t s = (size_t)blockIdx.x * (RW * 8) + warp * 8u + row;          // this lane's sample (A row / C row)
    uint64_t cw[MAXW] = {0ull, 0ull, 0ull, 0ull};
    #pragma unroll
    for(unsigned w = 0; w < (unsigned)MAXW; w++) if(s < ns && w < words) cw[w] = confs[s * words + w];
    const double* __restrict__ wr = reinterpret_cast<const double*>(psi.W);
    const unsigned ncol = 2u * M;
    const unsigned nkc = (N + AG_KC - 1) / AG_KC, ncb = (ncol + AG_CB - 1) / AG_CB, nstage = nkc * ncb;
    auto issue = [&](unsigned q) {
        double* dst = ag_buf + (q & 1u) * STAGE;
        const unsigned cb = (q / nkc) * AG_CB, i0 = (q % nkc) * AG_KC;
        for(unsigned e = threadIdx.x; e < AG_KC * (AG_CB / 2); e += RW * 32) {
            const unsigned kk = e / (AG_CB / 2), c = (e % (AG_CB / 2)) * 2u;
            const bool ok = (i0 + kk < N) && (cb + c < ncol);
            cp_async16_zfill(dst + kk * STRIDE + c, ok ? (const void*)(wr + (size_t)(i0 + kk) * ncol + cb + c) : (const void*)wr, ok);
        }
        cp_async_commit();
    };
    cplx p(0.0, 0.0);
    double acc[AG_CB / 8][2];
    #pragma unroll
    for(int t = 0; t < AG_CB / 8; t++) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    issue(0);
    for(unsigned q = 0; q < nstage; q++) {
        if(q + 1u < nstage) { issue(q + 1u); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();
        const double* Ws = ag_buf + (q & 1u) * STAGE;
        const unsigned cb = (q / nkc) * AG_CB, i0 = (q % nkc) * AG_KC;
        #pragma unroll 2
        for(unsigned k4 = 0; k4 < AG_KC / 4; k4++) {
            const unsigned i = i0 + k4 * 4u + kq;
            const unsigned wd = i >> 6;
            uint64_t word = cw[0];
            if(wd == 1u) word = cw[1];
            if(wd == 2u) word = cw[2];
            if(wd == 3u) word = cw[3];
            const double asg = (s < ns && i < N) ? (((word >> (i & 63u)) & 1ull) ? 1.0 : -1.0) : 0.0;
            const double* wrow = Ws + (k4 * 4u + kq) * STRIDE + row;
            #pragma unroll
            for(int t = 0; t < AG_CB / 8; t++) dmma(acc[t][0], acc[t][1], asg, wrow[t * 8]);
        }
        if(q % nkc == nkc - 1u) {                          // column block complete: these AG_CB / 2 angles of the 8 samples are final
            #pragma unroll
            for(int t = 0; t < AG_CB / 8; t++) {
                const unsigned j = (cb >> 1) + (unsigned)t * 4u + kq;
                if(s < ns && j < M) {
                    const cplx th(acc[t][0], acc[t][1]);
                    if(angles_out) angles_out[s * M + j] = th;
                    p += lc0(th);
                }
                acc[t][0] = 0.0; acc[t][1] = 0.0;
            }
        }
        __syncthreads();                                   // the buffer is refilled by the next iteration's issue
    }
    p.re += __shfl_xor_sync(FULL, p.re, 1); p.im += __shfl_xor_sync(FULL, p.im, 1);
    p.re += __shfl_xor_sync(FULL, p.re, 2); p.im += __shfl_xor_sync(FULL, p.im, 2);
    if(kq == 0 && s < ns) {
        const cplx lp = psi.lp + psi.fw * p;
        if(log_psi_out) log_psi_out[s] = lp;
        if(weight_out) weight_out[s] = exp(2.0 * lp.re);
    }
}

constexpr int RBM_ELOC_MAXF = 4;   // flips per group handled by the fast path (Heisenberg/TFIM: <= 2)
constexpr int RBM_ELOC_SPLIT = 4;  // lanes per flip group: each takes the hidden units j = p (mod 4)
constexpr int RBM_ELOC_WARPS = 8;  // warps per block

// smem per TEAM (= the WPS warps sharing one sample): theta[M] cplx | list_C[num_groups] cplx | red[2*WPS] cplx |
// list_g[num_groups] unsigned | count (padded to 16 B)
__host__ __device__ inline size_t rbm_eloc_slice_bytes(unsigned M, unsigned num_groups, unsigned wps = 1u) {
    return (size_t)(M + num_groups + 2u * wps) * sizeof(cplx) + ((((size_t)num_groups + 1u) * sizeof(unsigned) + 15u) & ~(size_t)15u);
}

// barrier over the WPS warps of a team (a whole block when WPS == RBM_ELOC_WARPS): named barrier 1 + team id
template<int WPS>
__device__ __forceinline__ void team_sync(unsigned team) {
    if(WPS == 1) __syncwarp();
    else if(WPS == RBM_ELOC_WARPS) __syncthreads();
    else asm volatile("bar.sync %0, %1;" :: "r"(team + 1u), "r"(WPS * 32) : "memory");
}

// E_loc for PsiRBM: a TEAM of WPS warps per sample.  Work item = (active flip group, quarter of the hidden units): 8 items
// per warp and pass, so a sample with `a` active groups needs ceil(a / (8 WPS)) passes of M/4 units each (a = number of
// antiparallel bonds ~ N/2 for the Heisenberg ring).  theta_j is shared by the team in shared memory (4 consecutive
// complex per quad: conflict-free, broadcast across the 8 groups of a warp); W[site][j] from L1/L2 (4 consecutive
// hidden units = 64 contiguous bytes per quad and site).  NF = max flips per group (compile time; unused slots have
// delta = 0).  WPS > 1 keeps the per-sample scratch (16 M bytes) at one copy per team -- for M = 1600 a warp-private copy
// limits an SM to 5 resident warps -- and makes the work per scheduling unit finer (less tail at ns / SMs ~ 55).
template<int NF, int WPS>
__global__ void __launch_bounds__(RBM_ELOC_WARPS * 32)
k_eloc_rbm(const RbmDev psi, const OpDev op, const uint64_t* __restrict__ confs, const cplx* __restrict__ angles,
           size_t ns, cplx* __restrict__ eloc_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned TEAMS = RBM_ELOC_WARPS / WPS, TT = WPS * 32u;      // teams per block, threads per team
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned team = warp / WPS, wid = warp % WPS, ttid = wid * 32u + lane;
    const unsigned M = psi.M, G = op.num_groups;
    unsigned char* base = smem_raw + (size_t)team * rbm_eloc_slice_bytes(M, G, WPS);
    cplx* theta = reinterpret_cast<cplx*>(base);
    cplx* list_C = theta + M;
    cplx* red = list_C + G;                                               // [2][WPS]
    unsigned* list_g = reinterpret_cast<unsigned*>(red + 2 * WPS);
    unsigned* count_sh = list_g + G;
    const cplx* __restrict__ W = psi.W;
    const unsigned part = lane & (RBM_ELOC_SPLIT - 1), slot = wid * (32u / RBM_ELOC_SPLIT) + lane / RBM_ELOC_SPLIT;

    // team-wide sum of one complex per thread, identical on all threads of the team
    auto team_sum = [&](cplx v, unsigned which) -> cplx {
        v = warp_sum(v);
        if(WPS == 1) return v;
        if(lane == 0) red[which * WPS + wid] = v;
        team_sync<WPS>(team);
        cplx t(0.0, 0.0);
        #pragma unroll
        for(int q = 0; q < WPS; q++) t += red[which * WPS + q];
        return t;
    };

    for(size_t s = (size_t)blockIdx.x * TEAMS + team; s < ns; s += (size_t)gridDim.x * TEAMS) {
        uint64_t conf[MAXW];
        conf_load(conf, confs + s * psi.words, psi.words);
        cplx basep(0.0, 0.0);
        for(unsigned j = ttid; j < M; j += TT) { const cplx a = angles[s * M + j]; theta[j] = a; basep += lc0_pq(a.re, a.im); }

        // diagonal strings (spread over the team) + per-group coefficients, compacted by the team's first warp
        cplx E(0.0, 0.0);
        for(unsigned n = ttid; n < op.num_diag; n += TT) E += string_sign_reg(op, n, conf) * op.coef[n];
        if(wid == 0) {
            unsigned count = 0;
            for(unsigned g0 = 0; g0 < G; g0 += 32u) {
                const unsigned g = g0 + lane;
                cplx C(0.0, 0.0);
                if(g < G) C = strings_coefficient_reg(op, op.group_begin[g], op.group_begin[g + 1u], conf);
                const bool active = (C.re != 0.0 || C.im != 0.0);
                const unsigned ballot = __ballot_sync(FULL, active);
                if(active) {
                    const unsigned pos = count + __popc(ballot & ((1u << lane) - 1u));
                    list_C[pos] = C; list_g[pos] = g;
                }
                count += __popc(ballot);
            }
            if(lane == 0) *count_sh = count;
        }
        const cplx base_sum = team_sum(basep, 0u);       // (WPS > 1: its barrier also publishes theta, the lists and count)
        if(WPS == 1) __syncwarp();
        const unsigned count = *count_sh;

        for(unsigned idx0 = 0; idx0 < count; idx0 += TT / RBM_ELOC_SPLIT) {
            const unsigned idx = idx0 + slot;
            const bool valid = idx < count;
            unsigned site[NF]; double dl[NF];
            #pragma unroll
            for(int f = 0; f < NF; f++) { site[f] = 0u; dl[f] = 0.0; }
            if(valid) {
                const unsigned g = list_g[idx];
                int nf = 0;
                for(unsigned w = 0; w < op.words; w++) {
                    uint64_t m = op.flip[g * op.words + w];
                    while(m) {
                        const unsigned p = w * 64u + (unsigned)__ffsll((long long)m) - 1u;
                        #pragma unroll
                        for(int f = 0; f < NF; f++) if(f == nf) { site[f] = p; dl[f] = -2.0 * conf_spin(conf, p); }
                        nf++;
                        m &= m - 1ull;
                    }
                }
            }
            const cplx* __restrict__ wr[NF];
            #pragma unroll
            for(int f = 0; f < NF; f++) wr[f] = W + (size_t)site[f] * M;
            cplx acc(0.0, 0.0);
            #pragma unroll 4
            for(unsigned j = part; j < M; j += RBM_ELOC_SPLIT) {
                cplx a = theta[j];
                #pragma unroll
                for(int f = 0; f < NF; f++) { const cplx w = ldg(&wr[f][j]); a.re = fma(dl[f], w.re, a.re); a.im = fma(dl[f], w.im, a.im); }
                acc += lc0_pq(a.re, a.im);
            }
            // combine the 4 quarters of each group
            #pragma unroll
            for(int o = 1; o < RBM_ELOC_SPLIT; o <<= 1) {
                acc.re += __shfl_xor_sync(FULL, acc.re, o);
                acc.im += __shfl_xor_sync(FULL, acc.im, o);
            }
            if(valid && part == 0) E += list_C[idx] * cexp(psi.fw * (acc - base_sum));
        }
        E = team_sum(E, 1u);
        if(ttid == 0) eloc_out[s] = E;
        team_sync<WPS>(team);                             // the scratch is reused by the next sample
    }
}

// E_loc for PsiRBM, M <= 256 and <= 2 flips per group: the W ROWS OF A FLIP GROUP LIVE IN REGISTERS and are applied to a TILE of
// samples whose angles sit in shared memory.  k_eloc_rbm streams the rows of every active group of every sample through L1 (C2:
// 32 groups x 8 KB per sample, 2.1 GB per call, L1 throughput 85 %: the bound); here a warp owns a flip group, reads its rows once
// (lane l holds the hidden units l + 32 k) and walks the samples of the tile -- 8 KB of W per (group, tile) instead of per
// (group, sample), the angles (4 KB per sample and group) being the only per-pair traffic, from shared memory.
//   block = 8 warps, tile = st <= 32 samples (lane t keeps sample t's configuration, coefficient and partial E_loc);
//   warp w takes the groups g = w, w + 8, ...; per group: coefficients of all samples (lane-parallel), then for every ACTIVE sample
//   sum_j lc0(theta_tj + d0 W_aj + d1 W_bj) over the warp; the exponentials of a group are evaluated together, one lane per sample.
// The tile size is chosen by the host so that the tiles fill the resident blocks in whole rounds (launch_eloc_rbm_tile).
constexpr int ET_WARPS = 8, ET_MAXS = 32, ET_BLOCKS_PER_SM = 2;   // two rows of K complex per lane: <= 128 registers
__host__ __device__ inline size_t eloc_tile_smem(unsigned Mp, unsigned st, unsigned words) {
    return (size_t)st * Mp * sizeof(cplx) + (size_t)(2 * ET_MAXS + ET_WARPS * ET_MAXS) * sizeof(cplx) + (size_t)ET_MAXS * words * sizeof(uint64_t);
}
template<int K>
__global__ void __launch_bounds__(ET_WARPS * 32, ET_BLOCKS_PER_SM)
k_eloc_rbm_tile(const RbmDev psi, const OpDev op, const uint64_t* __restrict__ confs, const cplx* __restrict__ angles,
                size_t ns, unsigned st, cplx* __restrict__ eloc_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned Mp = 32u * K;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned M = psi.M, G = op.num_groups, words = op.words;
    cplx* theta = reinterpret_cast<cplx*>(smem_raw);                        // [st][Mp], zero beyond M
    cplx* base = theta + (size_t)st * Mp;                                    // [ET_MAXS]  sum_j lc0(theta_tj)
    cplx* diag = base + ET_MAXS;                                             // [ET_MAXS]  diagonal strings
    cplx* epart = diag + ET_MAXS;                                            // [ET_WARPS][ET_MAXS]
    uint64_t* cs = reinterpret_cast<uint64_t*>(epart + ET_WARPS * ET_MAXS);  // [ET_MAXS][words]
    const cplx* __restrict__ W = psi.W;
    const size_t ntiles = (ns + st - 1) / st;
    for(size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t s0 = tile * st;
        const unsigned cnt = (unsigned)min((size_t)st, ns - s0);
        __syncthreads();                                                     // the previous tile is fully consumed
        for(unsigned e = threadIdx.x; e < cnt * Mp; e += ET_WARPS * 32) {
            const unsigned t = e / Mp, j = e % Mp;
            theta[e] = (j < M) ? angles[(s0 + t) * M + j] : cplx(0.0, 0.0);
        }
        for(unsigned e = threadIdx.x; e < cnt * words; e += ET_WARPS * 32) cs[e] = confs[s0 * words + e];
        __syncthreads();
        for(unsigned t = warp; t < cnt; t += ET_WARPS) {                     // per-sample constants
            cplx b(0.0, 0.0), dg(0.0, 0.0);
            #pragma unroll
            for(int k = 0; k < K; k++) { const cplx a = theta[t * Mp + lane + 32u * k]; b += lc0_pq(a.re, a.im); }
            uint64_t c[MAXW]; conf_load(c, cs + t * words, words);
            for(unsigned n = lane; n < op.num_diag; n += 32u) dg += string_sign_reg(op, n, c) * op.coef[n];
            b = warp_sum(b); dg = warp_sum(dg);
            if(lane == 0) { base[t] = b; diag[t] = dg; }
        }
        __syncthreads();
        uint64_t cmine[MAXW] = {0ull, 0ull, 0ull, 0ull};                      // lane t: the configuration of sample t
        if(lane < cnt) conf_load(cmine, cs + lane * words, words);
        const cplx base_mine = (lane < cnt) ? base[lane] : cplx(0.0, 0.0);
        cplx E_mine(0.0, 0.0);
        for(unsigned g = warp; g < G; g += ET_WARPS) {
            unsigned site[2] = {0u, 0u}; int nf = 0;
            for(unsigned w = 0; w < words; w++) {
                uint64_t m = op.flip[g * words + w];
                while(m) { if(nf < 2) site[nf] = w * 64u + (unsigned)__ffsll((long long)m) - 1u; nf++; m &= m - 1ull; }
            }
            cplx w0[K], w1[K];
            #pragma unroll
            for(int k = 0; k < K; k++) {
                const unsigned j = lane + 32u * k;
                w0[k] = (j < M) ? ldg(&W[(size_t)site[0] * M + j]) : cplx(0.0, 0.0);
                w1[k] = (j < M && nf > 1) ? ldg(&W[(size_t)site[1] * M + j]) : cplx(0.0, 0.0);
            }
            cplx C(0.0, 0.0);
            if(lane < cnt) C = strings_coefficient_reg(op, op.group_begin[g], op.group_begin[g + 1u], cmine);
            // the two spins of every sample at the group's sites, as +-2 factors: bit 0 / 1 of `sb`
            const unsigned sb = (conf_spin(cmine, site[0]) > 0.0 ? 1u : 0u) | ((nf > 1 && conf_spin(cmine, site[1]) > 0.0) ? 2u : 0u);
            unsigned active = __ballot_sync(FULL, C.re != 0.0 || C.im != 0.0);
            cplx mine(0.0, 0.0);
            // two active samples per pass: their warp reductions (5 dependent shuffle rounds each) overlap
            auto trial = [&](unsigned t) -> cplx {
                const unsigned sbt = __shfl_sync(FULL, sb, t);
                const double d0 = (sbt & 1u) ? -2.0 : 2.0, d1 = (nf > 1) ? ((sbt & 2u) ? -2.0 : 2.0) : 0.0;       // s' - s
                const cplx* __restrict__ th = theta + t * Mp + lane;
                cplx acc(0.0, 0.0);
                #pragma unroll
                for(int k = 0; k < K; k++) {
                    const cplx a = th[32u * k];
                    const double x = fma(d1, w1[k].re, fma(d0, w0[k].re, a.re)), y = fma(d1, w1[k].im, fma(d0, w0[k].im, a.im));
                    acc += lc0_pq(x, y);
                }
                return acc;
            };
            while(active) {
                const unsigned t0 = (unsigned)__ffs((int)active) - 1u;
                active &= active - 1u;
                const bool two = active != 0u;                                   // warp-uniform
                const unsigned t1 = two ? (unsigned)__ffs((int)active) - 1u : t0;
                if(two) active &= active - 1u;
                cplx a0 = trial(t0), a1(0.0, 0.0);
                if(two) a1 = trial(t1);
                #pragma unroll
                for(int o = 16; o > 0; o >>= 1) {
                    a0.re += __shfl_xor_sync(FULL, a0.re, o); a0.im += __shfl_xor_sync(FULL, a0.im, o);
                    a1.re += __shfl_xor_sync(FULL, a1.re, o); a1.im += __shfl_xor_sync(FULL, a1.im, o);
                }
                if(lane == t0) mine = a0;
                if(two && lane == t1) mine = a1;
            }
            if(C.re != 0.0 || C.im != 0.0) E_mine += C * cexp(psi.fw * (mine - base_mine));
        }
        epart[warp * ET_MAXS + lane] = E_mine;
        __syncthreads();
        if(threadIdx.x < cnt) {
            cplx E = diag[threadIdx.x];
            #pragma unroll
            for(int w = 0; w < ET_WARPS; w++) E += epart[w * ET_MAXS + threadIdx.x];
            eloc_out[s0 + threadIdx.x] = E;
        }
    }
}

// T[s][j] = final_weight * th0(theta_sj)
__global__ void k_rbm_T(const RbmDev psi, const cplx* __restrict__ angles, size_t total, cplx* __restrict__ T) {
    for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        T[i] = psi.fw * act_th(angles[i], 0u);
}

// dense rows from the factorised form: O[s][i*M + j] = sigma_si * T[s][j]   (HBM-write bound: ns*N*M*16 B)
__global__ void k_rbm_dense_O(const RbmDev psi, const uint64_t* __restrict__ confs, const cplx* __restrict__ T,
                              size_t ns, cplx* __restrict__ O) {
    const unsigned M = psi.M, N = psi.N;
    const size_t s = blockIdx.x;
    if(s >= ns) return;
    uint64_t conf[MAXW];
    conf_load(conf, confs + s * psi.words, psi.words);
    cplx* row = O + s * (size_t)psi.P;
    for(unsigned k = threadIdx.x; k < N * M; k += blockDim.x) {
        const unsigned i = k / M, j = k - i * M;
        row[k] = conf_spin(conf, i) * T[s * M + j];
    }
}

#endif // __CUDACC__

} // namespace angpu

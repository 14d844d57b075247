// PsiRBM fast paths (the BASELINE.json headline path: MC sampling -> E_loc -> O_k for an RBM).
//
//  k_mc_rbm    one warp per Markov chain; the M hidden-unit angles live in REGISTERS (K = ceil(M/32) complex
//              per lane), a spin flip costs one coalesced W-row read (L1/L2 resident) + K polynomial
//              evaluations per lane + ONE warp-shuffle reduction; no shared memory, no barriers, no atomics in
//              the loop (the reference: >= 8 __syncthreads + an M-way shared-atomic reduce + a global atomic per
//              proposal, include/ensembles/MonteCarlo.hpp:135-177, PsiRBM.hpp:92-157).
//              Bound: FP64 pipe (~16*M DFMA per proposal) — see DESIGN.md; compulsory HBM traffic is ~0.
//  k_eloc_rbm  one warp per sample, ONE LANE PER FLIP GROUP: each lane walks the M hidden units for its own s',
//              reading theta_j as a shared-memory broadcast and W^T[j][site] coalesced across lanes, so a
//              psi(s')/psi(s) evaluation needs no cross-lane reduction at all.  Groups whose coefficient is
//              exactly zero on s (XX+YY on aligned spins) are compacted away first.
//  k_rbm_T     the factorised log-derivative: O[s][i*M+j] = sigma_si * T[s][j], T = fw * th0(theta)
//              (PsiRBM.hpp:161-176) — ns*M complex instead of ns*N*M.
//  k_rbm_dense_O  materialises the dense rows only when a caller asks for O_k_samples / a dense S.
#pragma once
#include "kernels.cuh"

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

template<int K, bool FW_REAL>
__global__ void __launch_bounds__(128)
k_mc_rbm(const RbmDev psi, const McParams mc, uint64_t* __restrict__ conf_out, cplx* __restrict__ log_psi_out,
         cplx* __restrict__ angles_out, unsigned long long* __restrict__ acc_rej) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if(chain >= mc.num_chains_local) return;
    const unsigned gchain = mc.chain0 + chain;
    const unsigned M = psi.M, N = psi.N;
    const cplx* __restrict__ W = psi.W;

    uint32_t r[4];
    uint64_t conf[MAXW] = {0ull, 0ull, 0ull, 0ull};
    #pragma unroll
    for(unsigned w = 0; w < (unsigned)MAXW; w++) {
        if(w < psi.words) {
            philox4x32_10(w, 0u, gchain, (mc.call << 1) | 0u, mc.seed_lo, mc.seed_hi, r);
            conf[w] = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
            if(w == psi.words - 1u && (N & 63u)) conf[w] &= (1ull << (N & 63u)) - 1ull;
        }
    }

    cplx th[K];
    #pragma unroll
    for(int k = 0; k < K; k++) th[k] = cplx(0.0, 0.0);
    for(unsigned i = 0; i < N; i++) {
        const double s = conf_spin(conf, i);
        #pragma unroll
        for(int k = 0; k < K; k++) {
            const unsigned j = lane + 32u * k;
            if(j < M) th[k] += s * ldg(&W[i * M + j]);
        }
    }
    // current value of Re log psi (only its changes matter for the Metropolis ratio)
    auto re_log_psi = [&](const cplx (&a)[K]) -> double {
        if(FW_REAL) {
            double p = 0.0;
            #pragma unroll
            for(int k = 0; k < K; k++) if(lane + 32u * k < M) p += lc0_re(a[k]);
            return fma(psi.fw.re, warp_sum(p), psi.lp.re);
        } else {
            cplx p(0.0, 0.0);
            #pragma unroll
            for(int k = 0; k < K; k++) if(lane + 32u * k < M) p += lc0(a[k]);
            p = warp_sum(p);
            return psi.lp.re + psi.fw.re * p.re - psi.fw.im * p.im;
        }
    };
    double cur_re = re_log_psi(th);

    unsigned long long t = 0, acc = 0, rej = 0;
    const unsigned therm = mc.num_therm * N, per_sample = mc.num_sweeps * N;
    for(unsigned s = 0; s <= mc.steps_per_chain; s++) {
        const unsigned nsteps = (s == 0) ? therm : per_sample;
        for(unsigned i = 0; i < nsteps; i++, t++) {
            philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), gchain, (mc.call << 1) | 1u, mc.seed_lo, mc.seed_hi, r);
            const unsigned site = r[0] % N;
            const double delta = -2.0 * conf_spin(conf, site);       // s'_p - s_p
            const cplx* __restrict__ row = W + (size_t)site * M;
            cplx nth[K];
            #pragma unroll
            for(int k = 0; k < K; k++) {
                const unsigned j = lane + 32u * k;
                nth[k] = th[k];
                if(j < M) { const cplx w = ldg(&row[j]); nth[k].re = fma(delta, w.re, th[k].re); nth[k].im = fma(delta, w.im, th[k].im); }
            }
            const double new_re = re_log_psi(nth);
            const double ratio = exp(2.0 * (new_re - cur_re));
            const double u = u01_from_bits(r[1], r[2]);
            if(ratio > 1.0 || u <= ratio) {
                #pragma unroll
                for(int k = 0; k < K; k++) th[k] = nth[k];
                cur_re = new_re;
                conf_flip(conf, site);
                acc++;
            } else rej++;
        }
        if(s == 0) continue;
        const size_t idx = (size_t)(s - 1u) * mc.num_chains_local + chain;
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
            for(unsigned w = 0; w < (unsigned)MAXW; w++) if(w < psi.words) conf_out[idx * psi.words + w] = conf[w];
        }
    }
    if(lane == 0) { atomicAdd(&acc_rej[0], acc); atomicAdd(&acc_rej[1], rej); }
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

constexpr int RBM_ELOC_MAXF = 4;   // flips per group handled by the fast path (Heisenberg/TFIM: <= 2)

// smem per warp: theta[M] cplx | list_C[num_groups] cplx | list_g[num_groups] unsigned (padded to 16 B)
__host__ __device__ inline size_t rbm_eloc_slice_bytes(unsigned M, unsigned num_groups) {
    return (size_t)M * sizeof(cplx) + (size_t)num_groups * sizeof(cplx) + (((size_t)num_groups * sizeof(unsigned) + 15u) & ~(size_t)15u);
}

__global__ void __launch_bounds__(256)
k_eloc_rbm(const RbmDev psi, const OpDev op, const uint64_t* __restrict__ confs, const cplx* __restrict__ angles,
           size_t ns, cplx* __restrict__ eloc_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned lane = threadIdx.x & 31u, wpb = blockDim.x >> 5;
    const unsigned M = psi.M, N = psi.N, G = op.num_groups;
    unsigned char* base = smem_raw + (size_t)(threadIdx.x >> 5) * rbm_eloc_slice_bytes(M, G);
    cplx* theta = reinterpret_cast<cplx*>(base);
    cplx* list_C = theta + M;
    unsigned* list_g = reinterpret_cast<unsigned*>(list_C + G);
    const cplx* __restrict__ Wt = psi.Wt;

    for(size_t s = (size_t)blockIdx.x * wpb + (threadIdx.x >> 5); s < ns; s += (size_t)gridDim.x * wpb) {
        uint64_t conf[MAXW];
        conf_load(conf, confs + s * psi.words, psi.words);
        cplx basep(0.0, 0.0);
        for(unsigned j = lane; j < M; j += 32u) { const cplx a = angles[s * M + j]; theta[j] = a; basep += lc0(a); }
        const cplx base_sum = warp_sum(basep);

        // diagonal strings + per-group coefficients, compacted
        cplx E(0.0, 0.0);
        for(unsigned n = lane; n < op.num_diag; n += 32u) E += string_sign_reg(op, n, conf) * op.coef[n];
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
        __syncwarp();

        for(unsigned idx0 = 0; idx0 < count; idx0 += 32u) {
            const unsigned idx = idx0 + lane;
            if(idx < count) {
                const unsigned g = list_g[idx];
                unsigned site[RBM_ELOC_MAXF]; double dl[RBM_ELOC_MAXF];
                int nf = 0;
                #pragma unroll
                for(int f = 0; f < RBM_ELOC_MAXF; f++) { site[f] = 0u; dl[f] = 0.0; }
                for(unsigned w = 0; w < op.words; w++) {
                    uint64_t m = op.flip[g * op.words + w];
                    while(m) {
                        const unsigned p = w * 64u + (unsigned)__ffsll((long long)m) - 1u;
                        #pragma unroll
                        for(int f = 0; f < RBM_ELOC_MAXF; f++) if(f == nf) { site[f] = p; dl[f] = -2.0 * conf_spin(conf, p); }
                        nf++;
                        m &= m - 1ull;
                    }
                }
                cplx acc(0.0, 0.0);
                for(unsigned j = 0; j < M; j++) {
                    cplx a = theta[j];
                    const cplx* __restrict__ wr = Wt + (size_t)j * N;
                    #pragma unroll
                    for(int f = 0; f < RBM_ELOC_MAXF; f++) {
                        if(f < nf) { const cplx w = ldg(&wr[site[f]]); a.re = fma(dl[f], w.re, a.re); a.im = fma(dl[f], w.im, a.im); }
                    }
                    acc += lc0(a);
                }
                E += list_C[idx] * cexp(psi.fw * (acc - base_sum));
            }
        }
        E = warp_sum(E);
        if(lane == 0) eloc_out[s] = E;
        __syncwarp();
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

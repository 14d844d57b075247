// fp32-screened Metropolis samplers for PsiRBM with a real final weight (the BASELINE.json headline kernel).
//
// The reference decides a proposal with  ratio = exp(2 (Re log psi(s') - Re log psi(s))),  accept iff ratio > 1 || u <= ratio
// (include/ensembles/MonteCarlo.hpp:135-177), every quantity in fp64.  What the decision needs is only the SIGN of
//     2 fw sum_j [ Re lc0(theta_j + delta W_pj) - Re lc0(theta_j) ]  -  ln u ,
// and that sign is known from a cheap evaluation whenever the evaluation's error is smaller than the distance to zero.
// So every proposal is first evaluated in PACKED FP32 (FFMA2/FMUL2: two hidden units per instruction) against an fp32
// shadow of the angles and an fp32 copy of W, together with a PROVEN bound on its error; only when the bound cannot
// separate the two outcomes (~1e-3 of the proposals at the BASELINE shapes) the fp64 evaluation of the old sampler is
// run.  The accepted decision is therefore always the fp64 one -- chains stay identical, configuration by
// configuration, to the fp64 sampler and to the CPU oracle -- while the FP64 pipe is only used for the 2 DFMA per unit
// of the exact angle update on ACCEPTED proposals (the angles handed to E_loc / O_k are the fp64 ones).
// Nothing is updated, and nothing has to be undone, on rejection.
//
// Error bound (u = 2^-24, m >= |theta_j|^2, |theta'_j|^2 for every unit, K units per lane; derivation in DESIGN.md §4a):
//   shadow/trial inputs   |d theta'| <= (4 + 3 n) u sqrt(m)        n = accepted proposals since the shadow was last
//                                                                   converted from the fp64 angles (0 with REFRESH = 1)
//   -> through f = Re lc0: <= (4 + 3 n) u (m + m^2/3 + 2 m^3/15)    since |lc0'(z)| <= |z| + |z|^3/3 + 2|z|^5/15
//   polynomial rounding   <= u (2 m + 1.1 m^2 + 0.7 m^3)
//   accumulation (K FMAs per packed accumulator half, the final add, the lane difference)  <= u Fm(m) (K/2 + 3) per unit,
//                          Fm(m) = m/2 + m^2/6 + 0.09 m^3 >= |f|
//   fixed-point warp sum  0.5 * 2^-20 per lane (REDUX.SUM.S32 on round(2^20 * lane difference): exact integer sum)
// The bound is evaluated with the MEASURED max of |theta'|^2 over the warp (one FFMA2 + one FMNMX3 per unit pair and a
// REDUX.MAX), not with an a-priori estimate, so it adapts to the network.
#pragma once
#include "rbm_kernels.cuh"

namespace angpu {

#ifdef __CUDACC__

#ifndef SCR_MINB
#define SCR_MINB 10
#endif
constexpr float SCR_U = 5.9604645e-8f;             // 2^-24
constexpr float SCR_FIX = 1048576.f;               // 2^20
constexpr float SCR_FIX_INV = 1.f / 1048576.f;

// largest max|theta|^2 for which a lane difference cannot overflow the fixed-point sum: 32 * 2 K Fm(m) < 2^31 / 2^20
// <=> K Fm(m) < 32 (a margin of 2 is kept)
__host__ __device__ constexpr float scr_mcap(int K) { return K <= 2 ? 4.5f : K <= 4 ? 3.3f : K <= 8 ? 2.4f : K <= 16 ? 1.7f : 1.1f; }

// bound on |fp32 sum - exact sum| of Re lc0 over `units` hidden units of ONE state (see the header); n = accepted
// proposals since the last shadow refresh
template<int K>
__device__ __forceinline__ float scr_state_bound(float m, float n, float units) {
    const float m2 = m * m, m3 = m2 * m;
    const float in = (4.f + 3.f * n) * (m + m2 * (1.f / 3.f) + m3 * (2.f / 15.f));
    const float ev = 2.f * m + 1.1f * m2 + 0.7f * m3;
    const float fm = 0.5f * m + m2 * (1.f / 6.f) + 0.09f * m3;
    return 1.001f * units * SCR_U * (in + ev + fm * (0.5f * K + 3.f));
}

// 1 = accept, 0 = reject, -1 = the fp32 evaluation cannot decide.  d2 ~ 2 (Re log psi' - Re log psi) with |error| <= e;
// the uniform number lies in (uf_lo, uf_lo + 2^-24].
__device__ __forceinline__ int scr_decide(float d2, float e, float uf_lo) {
    if(!(e < 0.125f)) return -1;
    if(d2 >= e * 1.001f) return 1;                              // ratio >= 1
    const float rf = __expf(d2);                                 // |rel. error| <= 1.3e-5 for |d2| < 88
    const float er = e + e * e + 4e-5f;                          // exp(+-e) within 1 +- (e + e^2)
    if(uf_lo + SCR_U <= rf * (1.f - er)) return 1;               // u <= lower bound of the ratio
    if(uf_lo >= rf * (1.f + er)) return 0;                       // u >  upper bound of the ratio
    return -1;
}

// 16-byte read-only load that does not allocate in L1 (the fp64 rows are touched once per synchronisation; L1 is kept
// for the fp32 table, which every proposal reads)
__device__ __forceinline__ cplx ldg_stream(const cplx* p) {
    double a, b;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
    return cplx(a, b);
}

// packed evaluation of sum_k Re lc0 over the K = 2 KK units of a lane for the trial angles shadow + d * w
// (d = 0, +-2); also the lane's max of |theta'|^2
template<int KK>
__device__ __forceinline__ void scr_eval(const float2 (&xs)[KK], const float2 (&ys)[KK], float d, const float4* __restrict__ row,
                                         float& L, float& mx) {
    float2 acc = make_float2(0.f, 0.f);
    float mm = 0.f;
    const float2 dd = make_float2(d, d);
    #pragma unroll
    for(int kk = 0; kk < KK; kk++) {
        const float4 w = __ldg(&row[32 * kk]);
        const float2 x = __ffma2_rn(dd, make_float2(w.x, w.y), xs[kk]);
        const float2 y = __ffma2_rn(dd, make_float2(w.z, w.w), ys[kk]);
        const float2 yy = __fmul2_rn(y, y);
        const float2 p = __ffma2_rn(x, x, make_float2(-yy.x, -yy.y));
        const float2 m = __ffma2_rn(x, x, yy);
        const float2 q = __fmul2_rn(x, y);
        const float2 q2 = __fmul2_rn(q, q);
        float2 A = __ffma2_rn(p, make_float2(1.f / 45.f, 1.f / 45.f), make_float2(-1.f / 12.f, -1.f / 12.f));
        A = __ffma2_rn(A, p, make_float2(0.5f, 0.5f));
        const float2 B = __ffma2_rn(p, make_float2(-12.f / 45.f, -12.f / 45.f), make_float2(1.f / 3.f, 1.f / 3.f));
        acc = __ffma2_rn(q2, B, acc);
        acc = __ffma2_rn(A, p, acc);
        mm = fmaxf(fmaxf(mm, m.x), m.y);
    }
    L = acc.x + acc.y;
    mx = mm;
}

// bound on |fp32 sum - exact sum| over `units` hidden units of one state as a cubic in m with coefficients linear in n
// (the same polynomial as scr_state_bound, Horner form: 7 instructions in the proposal loop)
template<int K>
__device__ __forceinline__ float scr_state_bound_fast(float m, float n, float scale) {
    constexpr float H = 0.5f * K + 3.f;
    const float c1 = fmaf(3.f, n, 4.f + 2.f + 0.5f * H);
    const float c2 = fmaf(1.f, n, 4.f / 3.f + 1.1f + H / 6.f);
    const float c3 = fmaf(0.4f, n, 8.f / 15.f + 0.7f + 0.09f * H);
    return scale * m * fmaf(m, fmaf(m, c3, c2), c1);
}

// Wp: W with rows padded to Mp = 64 KK complex (zeros beyond M); lane l owns the units j = 64 kk + 2 l + h (h = 0, 1):
// two adjacent complex numbers per kk (32 contiguous bytes per lane, 1 KB per warp).
// Wf: the fp32 copy for the screen, float4 (Re W_pj0, Re W_pj1, Im W_pj0, Im W_pj1) at [(p KK + kk) 32 + l].
// The exact fp64 angles are brought up to date LAZILY: an accepted flip changes the configuration and the fp32 shadow
// (one more FFMA2 per unit pair on the re-read fp32 row) and is appended to a per-warp list in shared memory; every
// REFRESH accepted proposals -- and before a sample is recorded or a decision needs the fp64 evaluation -- the listed
// rows are added to the fp64 angles in the order they were accepted (2 DFMA per unit and flip, RP rows per pass so that
// their L2 latencies overlap) and the shadow is re-derived from them (F2F runs at a quarter of the DFMA rate on sm_100,
// hence not per proposal).  An undecided proposal first synchronises and is re-screened with the tight (n = 0) bound;
// only if it is still undecided the fp64 evaluation runs.
// stats[0..1] = accepted / rejected proposals, stats[2] = proposals decided by the fp64 evaluation.
template<int KK, int WORDS, int MINB, int REFRESH>
__global__ void __launch_bounds__(MC_RBM_THREADS, MINB)
k_mc_rbm_scr(const RbmDev psi, const cplx* __restrict__ Wp, const float4* __restrict__ Wf, const McParams mc,
             uint64_t* __restrict__ conf_out, cplx* __restrict__ log_psi_out, cplx* __restrict__ angles_out,
             unsigned long long* __restrict__ stats) {
    constexpr int K = 2 * KK;
    constexpr unsigned Mp = 64u * KK;
    constexpr int RP = (K >= 8) ? 2 : (K == 4) ? 4 : 8;        // rows per synchronisation pass (<= 32 complex in flight per lane)
    static_assert(REFRESH <= 32 && REFRESH % RP == 0, "REFRESH: a multiple of the pass size, at most 32");
    __shared__ unsigned short flips_sh[MC_RBM_THREADS / 32][32];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if(chain >= mc.num_chains_local) return;
    unsigned short* flips = flips_sh[threadIdx.x >> 5];         // accepted flips since the last sync: site | (new spin up ? 0x8000 : 0)
    const unsigned gchain = mc.chain0 + chain;
    const unsigned M = psi.M, N = psi.N;
    const unsigned tag_init = (mc.call << 1) | 0u, tag_step = (mc.call << 1) | 1u;
    const cplx* __restrict__ Wl = Wp + 2u * lane;
    const float4* __restrict__ Wfl = Wf + lane;

    uint32_t r[4];
    uint64_t conf[MAXW] = {0ull, 0ull, 0ull, 0ull};
    #pragma unroll
    for(int w = 0; w < WORDS; w++) {
        philox4x32_10((uint32_t)w, 0u, gchain, tag_init, mc.seed_lo, mc.seed_hi, r);
        conf[w] = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
    }
    if(N & 63u) conf[WORDS - 1] &= (1ull << (N & 63u)) - 1ull;

    cplx th[K];                                                // exact angles (up to the listed flips)
    #pragma unroll
    for(int k = 0; k < K; k++) th[k] = cplx(0.0, 0.0);
    for(unsigned i = 0; i < N; i++) {
        const double s = conf_spin_t<WORDS>(conf, i);
        #pragma unroll
        for(int k = 0; k < K; k++) th[k] += s * ldg_stream(&Wl[(size_t)i * Mp + 64u * (k >> 1) + (k & 1)]);
    }

    float2 xs[KK], ys[KK];                                     // fp32 shadow of the angles of `conf`
    const float bscale = 1.001f * 32.f * K * SCR_U;
    const float c2fw = psi.c2fw, ac2fw = fabsf(c2fw) * 1.001f;
    const float mcap = scr_mcap(K);
    const float INF = __int_as_float(0x7f800000);
    float curL, ecur, mrun;                                    // fp32 sum of the current state; its error bound (+ the fixed-point
                                                               // quantisation); max |theta|^2 since the last sync
    unsigned nacc = 0;                                         // accepted proposals since the last sync

    // th <- exact angles of `conf`; shadow, current sum and bounds re-derived from them
    auto sync_exact = [&]() {
        __syncwarp();
        for(unsigned q0 = 0; q0 < nacc; q0 += RP) {
            const cplx* __restrict__ row[RP];
            double dl[RP];
            #pragma unroll
            for(int q = 0; q < RP; q++) {
                const unsigned e = (q0 + q < nacc) ? (unsigned)flips[q0 + q] : 0u;
                dl[q] = (q0 + q < nacc) ? ((e & 0x8000u) ? 2.0 : -2.0) : 0.0;
                row[q] = Wl + (size_t)(e & 0x7fffu) * Mp;
            }
            cplx wv[RP][K];
            #pragma unroll
            for(int q = 0; q < RP; q++)
                #pragma unroll
                for(int k = 0; k < K; k++) wv[q][k] = ldg_stream(&row[q][64u * (k >> 1) + (k & 1)]);
            #pragma unroll
            for(int q = 0; q < RP; q++)
                #pragma unroll
                for(int k = 0; k < K; k++) { th[k].re = fma(dl[q], wv[q][k].re, th[k].re); th[k].im = fma(dl[q], wv[q][k].im, th[k].im); }
        }
        __syncwarp();
        #pragma unroll
        for(int kk = 0; kk < KK; kk++) {
            xs[kk] = make_float2(__double2float_rn(th[2 * kk].re), __double2float_rn(th[2 * kk + 1].re));
            ys[kk] = make_float2(__double2float_rn(th[2 * kk].im), __double2float_rn(th[2 * kk + 1].im));
        }
        float mx;
        scr_eval<KK>(xs, ys, 0.f, Wfl, curL, mx);
        mrun = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(mx))) * 1.0001f;
        ecur = (mrun < mcap) ? scr_state_bound_fast<K>(mrun, 0.f, bscale) + 16.f * SCR_FIX_INV : INF;
        nacc = 0;
    };

    const unsigned per_sample = mc.num_sweeps * N;
    const unsigned total_steps = mc.num_therm * N + per_sample * mc.steps_per_chain;   // < 2^32, checked by the host
    unsigned acc = 0, exact = 0, sample = 0;
    unsigned until_record = mc.num_therm * N + per_sample;      // proposals until the next recorded sample
    unsigned todo = 1u;                                         // bit 0: synchronise, bit 1: record a sample, bit 2: re-screened proposal
    unsigned my_site = 0, my_ulo = 0, my_uhi = 0;
    float my_uf = 0.f;

    for(unsigned t = 0; ; t++) {
        if(todo & 3u) {
            if((todo & 1u) || nacc) sync_exact();
            if(todo & 2u) {
                const size_t idx = (size_t)sample * mc.num_chains_local + chain;
                cplx p(0.0, 0.0);
                #pragma unroll
                for(int k = 0; k < K; k++) {
                    const unsigned j = 64u * (k >> 1) + 2u * lane + (k & 1);
                    if(j < M) { p += lc0(th[k]); if(angles_out) angles_out[idx * M + j] = th[k]; }
                }
                p = warp_sum(p);
                if(lane == 0) {
                    log_psi_out[idx] = psi.lp + psi.fw * p;
                    #pragma unroll
                    for(int ww = 0; ww < WORDS; ww++) conf_out[idx * WORDS + ww] = conf[ww];
                }
                sample++;
            }
            todo &= 4u;
        }
        if(t >= total_steps) break;
        const unsigned b = t & 31u;
        if(b == 0u && !todo) {
            // one Philox block per lane: lane l draws the random numbers of proposal t + l (amortises the generator 32x)
            philox4x32_10(t + lane, 0u, gchain, tag_step, mc.seed_lo, mc.seed_hi, r);
            my_site = r[0] % N; my_ulo = r[1]; my_uhi = r[2];
            my_uf = (float)(my_uhi >> 8) * SCR_U;               // u in (my_uf, my_uf + 2^-24]  (u01_from_bits)
        }
        todo = 0u;
        const unsigned site = __shfl_sync(FULL, my_site, b);
        const float uf_lo = __shfl_sync(FULL, my_uf, b);
        const bool up = conf_spin_t<WORDS>(conf, site) > 0.0;                              // current spin of the proposed site
        const float deltaf = up ? -2.f : 2.f;                                           // s'_p - s_p
        const float4* __restrict__ rowf = Wfl + site * (32u * KK);

        float newL, mx;
        scr_eval<KK>(xs, ys, deltaf, rowf, newL, mx);
        const int sum = __reduce_add_sync(FULL, __float2int_rn((newL - curL) * SCR_FIX));
        const float mb = fmaxf(__uint_as_float(__reduce_max_sync(FULL, __float_as_uint(mx))) * 1.0001f, mrun);
        const float enew = (mb < mcap) ? scr_state_bound_fast<K>(mb, (float)nacc, bscale) : INF;
        const float d2 = c2fw * ((float)sum * SCR_FIX_INV);
        const float e = fmaf(ac2fw, enew + ecur, 4e-7f * fabsf(d2));
        int dec = scr_decide(d2, e, uf_lo);
        if(dec < 0) {
            if(nacc) { todo = 5u; t--; continue; }              // re-screen against freshly converted angles
            // the fp64 evaluation of the reference, both states from the exact angles (th is in sync: nacc == 0)
            const double delta = (double)deltaf;
            const cplx* __restrict__ row = Wl + (size_t)site * Mp;
            double po0 = 0.0, po1 = 0.0, pn0 = 0.0, pn1 = 0.0;
            #pragma unroll
            for(int k = 0; k < K; k++) {
                const cplx w = ldg_stream(&row[64u * (k >> 1) + (k & 1)]);
                const double xn = fma(delta, w.re, th[k].re), yn = fma(delta, w.im, th[k].im);
                if(k & 1) { lc0_re_pq_acc(th[k].re, th[k].im, po1); lc0_re_pq_acc(xn, yn, pn1); }
                else      { lc0_re_pq_acc(th[k].re, th[k].im, po0); lc0_re_pq_acc(xn, yn, pn0); }
            }
            const double cur_re = fma(psi.fw.re, warp_sum(po0 + po1), psi.lp.re);
            const double new_re = fma(psi.fw.re, warp_sum(pn0 + pn1), psi.lp.re);
            const double u = u01_from_bits(__shfl_sync(FULL, my_ulo, b), __shfl_sync(FULL, my_uhi, b));
            dec = metropolis_accept(2.0 * (new_re - cur_re), u) ? 1 : 0;
            exact++;
        }
        if(dec) {
            const float2 dd = make_float2(deltaf, deltaf);
            #pragma unroll
            for(int kk = 0; kk < KK; kk++) {
                const float4 w = __ldg(&rowf[32 * kk]);
                xs[kk] = __ffma2_rn(dd, make_float2(w.x, w.y), xs[kk]);
                ys[kk] = __ffma2_rn(dd, make_float2(w.z, w.w), ys[kk]);
            }
            if(lane == 0) flips[nacc] = (unsigned short)(site | (up ? 0u : 0x8000u));
            conf_flip_t<WORDS>(conf, site);
            curL = newL; ecur = enew + 16.f * SCR_FIX_INV; mrun = mb; nacc++; acc++;
            if(nacc >= (unsigned)REFRESH) todo = 1u;
        }
        if(--until_record == 0u) { todo |= 2u; until_record = per_sample; }
    }
    if(lane == 0) {
        atomicAdd(&stats[0], (unsigned long long)acc); atomicAdd(&stats[1], (unsigned long long)(total_steps - acc));
        atomicAdd(&stats[2], (unsigned long long)exact);
    }
}

#endif // __CUDACC__

} // namespace angpu

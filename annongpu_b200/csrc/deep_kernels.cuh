// PsiDeep fast Metropolis sampler for hidden layers of width <= 64 (the reference's own maximum, PsiDeep.hpp:78-80):
// ONE BLOCK OF 256 THREADS RUNS 4 CHAINS with the deep-layer weight matrices held in registers as dense 64 x 64 tables
// wd[input k][unit j] (built on the host from lhs_connections / lhs_weights, zero where unconnected).
//
// The generic warp-per-chain kernel (kernels.cuh: k_mc<DeepDev>) re-reads the 64 x 64 complex weight matrix from L1
// for every proposal (one 16-byte load + one index load per complex FMA: L1-bandwidth and latency bound, ~15 % of the
// FP64 pipe at C4).  Here thread (ib, j) = (tid >> 6, tid & 63) keeps wd[ib*16 .. ib*16+15][j] of every deep layer in
// registers and applies them to the 4 chains of the block at once (4 independent accumulators per weight register);
// the 64 threads of group ib also own chain ib: its configuration, first-layer angle of unit j, Philox stream and
// Metropolis decision.  One round = one proposal for each of the 4 chains, three block barriers:
//   A  (group c)      first-layer angle of unit j updated in its register (update_input_units, PsiDeep.hpp:311-343)
//                     from the dense first-layer table w1d[site][j]; a_j = lc(angle_j)            -> smem actv[c]
//   B  (all threads)  partial_j^(c, ib) = sum_r wd[ib*16+r][j] * actv[c][ib*16+r], c = 0..3       -> smem part
//   C  (group c)      deep angle = sum_ib partial + bias; activation; * final_weight; warp sums   -> smem fin[c]
//   group c           log psi' = log_prefactor + fin[c][0] + fin[c][1]; Metropolis decision
// (forward_pass + log_psi_s, PsiDeep.hpp:173-266).  Rejections undo the angle update in place, like the reference.
// Same Philox stream as every other sampler: chains are trajectory-identical to the generic kernel and the CPU oracle up
// to the summation order inside a layer (4 partial sums over the inputs in index order instead of one running sum in
// connection order).
#pragma once
#include "kernels.cuh"
#include "rbm_kernels.cuh"

namespace angpu {

constexpr int DEEP_BLK_T = 256, DEEP_BLK_W = 64, DEEP_BLK_R = 16, DEEP_BLK_NC = DEEP_BLK_T / DEEP_BLK_W;   // 4 chains per block

#ifdef __CUDACC__

template<int NDEEP>
__global__ void __launch_bounds__(DEEP_BLK_T, NDEEP == 1 ? 2 : 1)
k_mc_deep_block(const DeepDev psi, const cplx* __restrict__ w1d, const McParams mc, uint64_t* __restrict__ conf_out,
                cplx* __restrict__ log_psi_out, unsigned long long* __restrict__ acc_rej) {
    __shared__ cplx actv[DEEP_BLK_NC][DEEP_BLK_W];
    __shared__ cplx part[DEEP_BLK_NC][DEEP_BLK_NC][DEEP_BLK_W];      // [chain][input block][unit]
    __shared__ cplx fin[DEEP_BLK_NC][2];
    const unsigned tid = threadIdx.x, lane = tid & 31u, j = tid & (DEEP_BLK_W - 1), ib = tid / DEEP_BLK_W;
    const unsigned c = ib;                                    // the chain slot this thread's group owns
    const unsigned chain_raw = blockIdx.x * DEEP_BLK_NC + c;
    const bool valid = chain_raw < mc.num_chains_local;       // a ragged last block re-runs the last chain without output
    const unsigned chain = valid ? chain_raw : mc.num_chains_local - 1u, gchain = mc.chain0 + chain;
    const unsigned N = psi.N, S1 = psi.L[1].size;
    const unsigned tag_init = (mc.call << 1) | 0u, tag_step = (mc.call << 1) | 1u;

    // register-resident parameters
    cplx wreg[NDEEP][DEEP_BLK_R], bias[NDEEP], fw(0.0, 0.0);
    #pragma unroll
    for(int d = 0; d < NDEEP; d++) {
        const DeepLayerDev& ly = psi.L[2 + d];
        const cplx* wd = w1d + ((size_t)N + (size_t)d * DEEP_BLK_W) * DEEP_BLK_W;      // dense [input k][unit j], zero padded
        #pragma unroll
        for(int r = 0; r < DEEP_BLK_R; r++) wreg[d][r] = ldg(&wd[(ib * DEEP_BLK_R + r) * DEEP_BLK_W + j]);
        bias[d] = (j < ly.size) ? ldg(&ly.bias[j]) : cplx(0.0, 0.0);
    }
    const unsigned S_last = psi.L[1 + NDEEP].size;
    if(j < S_last) fw = ldg(&psi.final_w[j]);

    // initial configuration (Init_Policy.hpp:16-24) and first-layer angles (compute_angles, PsiDeep.hpp:140-157)
    uint32_t r4[4];
    uint64_t conf[MAXW] = {0ull, 0ull, 0ull, 0ull};
    #pragma unroll
    for(unsigned w = 0; w < (unsigned)MAXW; w++) {
        if(w < psi.words) {
            philox4x32_10(w, 0u, gchain, tag_init, mc.seed_lo, mc.seed_hi, r4);
            conf[w] = (uint64_t)r4[0] | ((uint64_t)r4[1] << 32);
            if(w == psi.words - 1u && (N & 63u)) conf[w] &= (1ull << (N & 63u)) - 1ull;
        }
    }
    cplx ang(0.0, 0.0);
    if(j < S1) {
        const DeepLayerDev& l1 = psi.L[1];
        for(unsigned i = 0; i < l1.conn; i++) ang += conf_spin(conf, l1.lhs_c[i * S1 + j]) * ldg(&l1.lhs_w[i * S1 + j]);
        ang += ldg(&l1.bias[j]);
    }

    // deep layers + final sum for the current first-layer angles of the 4 chains; returns log psi of this group's chain
    auto forward = [&]() -> cplx {
        actv[c][j] = (j < S1) ? act_lc(ang, 0u) : cplx(0.0, 0.0);
        __syncthreads();
        cplx y(0.0, 0.0);
        #pragma unroll
        for(int d = 0; d < NDEEP; d++) {
            cplx p[DEEP_BLK_NC];
            #pragma unroll
            for(int cc = 0; cc < DEEP_BLK_NC; cc++) p[cc] = cplx(0.0, 0.0);
            #pragma unroll
            for(int r = 0; r < DEEP_BLK_R; r++) {
                #pragma unroll
                for(int cc = 0; cc < DEEP_BLK_NC; cc++) cfma(p[cc], wreg[d][r], actv[cc][ib * DEEP_BLK_R + r]);
            }
            #pragma unroll
            for(int cc = 0; cc < DEEP_BLK_NC; cc++) part[cc][ib][j] = p[cc];
            __syncthreads();
            const cplx s = ((part[c][0][j] + part[c][1][j]) + (part[c][2][j] + part[c][3][j])) + bias[d];
            y = (j < psi.L[2 + d].size) ? act_lc(s, (unsigned)d + 1u) : cplx(0.0, 0.0);
            if(d + 1 < NDEEP) { actv[c][j] = y; __syncthreads(); }
        }
        const cplx v = warp_sum(y * fw);
        if(lane == 0) fin[c][(tid >> 5) & 1u] = v;
        __syncthreads();
        return psi.lp + (fin[c][0] + fin[c][1]);
    };
    cplx cur = forward();

    const unsigned therm = mc.num_therm * N, per_sample = mc.num_sweeps * N;
    const unsigned long long total_steps = (unsigned long long)therm + (unsigned long long)per_sample * mc.steps_per_chain;
    unsigned long long acc = 0, next_record = (unsigned long long)therm + per_sample;
    unsigned sample = 0;
    const cplx* __restrict__ w1 = w1d + j;

    for(unsigned long long t0 = 0; t0 < total_steps; t0 += 32u) {
        // one Philox draw per lane = the next 32 proposals of this group's chain (both warps of the group compute it)
        philox4x32_10((uint32_t)(t0 + lane), (uint32_t)((t0 + lane) >> 32), gchain, tag_step, mc.seed_lo, mc.seed_hi, r4);
        const unsigned my_site = r4[0] % N, my_ulo = r4[1], my_uhi = r4[2];
        const unsigned nb = (unsigned)min((unsigned long long)32u, total_steps - t0);
        cplx wn = ldg(&w1[(size_t)__shfl_sync(FULL, my_site, 0) * DEEP_BLK_W]);
        for(unsigned b = 0; b < nb; b++) {
            const unsigned site = __shfl_sync(FULL, my_site, b);
            const double u = u01_from_bits(__shfl_sync(FULL, my_ulo, b), __shfl_sync(FULL, my_uhi, b));
            const double delta = -2.0 * conf_spin(conf, site);
            const cplx w = wn;
            ang.re = fma(delta, w.re, ang.re); ang.im = fma(delta, w.im, ang.im);
            if(b + 1u < nb) wn = ldg(&w1[(size_t)__shfl_sync(FULL, my_site, b + 1u) * DEEP_BLK_W]);
            const cplx nlp = forward();
            if(metropolis_accept(2.0 * (nlp.re - cur.re), u)) {
                cur = nlp;
                conf_flip(conf, site);
                acc++;
            } else {
                ang.re = fma(-delta, w.re, ang.re); ang.im = fma(-delta, w.im, ang.im);
            }
            if(t0 + b + 1u == next_record) {
                if(j == 0 && valid) {
                    const size_t idx = (size_t)sample * mc.num_chains_local + chain;
                    log_psi_out[idx] = cur;
                    #pragma unroll
                    for(unsigned ww = 0; ww < (unsigned)MAXW; ww++) if(ww < psi.words) conf_out[idx * psi.words + ww] = conf[ww];
                }
                sample++; next_record += per_sample;
            }
        }
    }
    if(j == 0 && valid) { atomicAdd(&acc_rej[0], acc); atomicAdd(&acc_rej[1], total_steps - acc); }
}

// E_loc for PsiDeep with the same register-resident deep layers: ONE BLOCK PER SAMPLE, the 4 thread groups evaluate 4
// active flip groups per round (psi(s')/psi(s) for 4 different s' at once), three barriers per round.
// (Operator.hpp:38-121 with PsiDeep.hpp:219-343.)  Dynamic shared memory: list_C[G] cplx | list_g[G] unsigned.
template<int NDEEP>
__global__ void __launch_bounds__(DEEP_BLK_T, NDEEP == 1 ? 2 : 1)
k_eloc_deep_block(const DeepDev psi, const cplx* __restrict__ w1d, const OpDev op, const uint64_t* __restrict__ confs, size_t ns,
                  cplx* __restrict__ eloc_out) {
    extern __shared__ __align__(16) unsigned char deep_dyn[];
    __shared__ cplx actv[DEEP_BLK_NC][DEEP_BLK_W];
    __shared__ cplx part[DEEP_BLK_NC][DEEP_BLK_NC][DEEP_BLK_W];
    __shared__ cplx fin[DEEP_BLK_NC][2];
    __shared__ cplx esum[DEEP_BLK_T / 32];
    __shared__ unsigned count_sh;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, j = tid & (DEEP_BLK_W - 1), ib = tid / DEEP_BLK_W;
    const unsigned c = ib, N = psi.N, S1 = psi.L[1].size, G = op.num_groups;
    cplx* list_C = reinterpret_cast<cplx*>(deep_dyn);
    unsigned* list_g = reinterpret_cast<unsigned*>(list_C + G);

    cplx wreg[NDEEP][DEEP_BLK_R], bias[NDEEP], fw(0.0, 0.0);
    #pragma unroll
    for(int d = 0; d < NDEEP; d++) {
        const DeepLayerDev& ly = psi.L[2 + d];
        const cplx* wd = w1d + ((size_t)N + (size_t)d * DEEP_BLK_W) * DEEP_BLK_W;
        #pragma unroll
        for(int r = 0; r < DEEP_BLK_R; r++) wreg[d][r] = ldg(&wd[(ib * DEEP_BLK_R + r) * DEEP_BLK_W + j]);
        bias[d] = (j < ly.size) ? ldg(&ly.bias[j]) : cplx(0.0, 0.0);
    }
    const unsigned S_last = psi.L[1 + NDEEP].size;
    if(j < S_last) fw = ldg(&psi.final_w[j]);
    const cplx* __restrict__ w1 = w1d + j;

    // deep layers + final sum for the first-layer angle `ang` of this thread's unit; log psi of this group's configuration
    auto forward = [&](cplx ang) -> cplx {
        actv[c][j] = (j < S1) ? act_lc(ang, 0u) : cplx(0.0, 0.0);
        __syncthreads();
        cplx y(0.0, 0.0);
        #pragma unroll
        for(int d = 0; d < NDEEP; d++) {
            cplx p[DEEP_BLK_NC];
            #pragma unroll
            for(int cc = 0; cc < DEEP_BLK_NC; cc++) p[cc] = cplx(0.0, 0.0);
            #pragma unroll
            for(int r = 0; r < DEEP_BLK_R; r++) {
                #pragma unroll
                for(int cc = 0; cc < DEEP_BLK_NC; cc++) cfma(p[cc], wreg[d][r], actv[cc][ib * DEEP_BLK_R + r]);
            }
            #pragma unroll
            for(int cc = 0; cc < DEEP_BLK_NC; cc++) part[cc][ib][j] = p[cc];
            __syncthreads();
            const cplx s = ((part[c][0][j] + part[c][1][j]) + (part[c][2][j] + part[c][3][j])) + bias[d];
            y = (j < psi.L[2 + d].size) ? act_lc(s, (unsigned)d + 1u) : cplx(0.0, 0.0);
            if(d + 1 < NDEEP) { actv[c][j] = y; __syncthreads(); }
        }
        const cplx v = warp_sum(y * fw);
        if(lane == 0) fin[c][(tid >> 5) & 1u] = v;
        __syncthreads();
        return psi.lp + (fin[c][0] + fin[c][1]);
    };

    for(size_t s = blockIdx.x; s < ns; s += gridDim.x) {
        uint64_t conf[MAXW];
        conf_load(conf, confs + s * psi.words, psi.words);
        cplx ang0(0.0, 0.0);
        if(j < S1) {
            const DeepLayerDev& l1 = psi.L[1];
            for(unsigned i = 0; i < l1.conn; i++) ang0 += conf_spin(conf, l1.lhs_c[i * S1 + j]) * ldg(&l1.lhs_w[i * S1 + j]);
            ang0 += ldg(&l1.bias[j]);
        }
        // diagonal strings over the block; active flip groups compacted by warp 0
        cplx E(0.0, 0.0);
        for(unsigned n = tid; n < op.num_diag; n += DEEP_BLK_T) E += string_sign_reg(op, n, conf) * op.coef[n];
        if(warp == 0) {
            unsigned count = 0;
            for(unsigned g0 = 0; g0 < G; g0 += 32u) {
                const unsigned g = g0 + lane;
                cplx C(0.0, 0.0);
                if(g < G) C = strings_coefficient_reg(op, op.group_begin[g], op.group_begin[g + 1u], conf);
                const bool active = (C.re != 0.0 || C.im != 0.0);
                const unsigned ballot = __ballot_sync(FULL, active);
                if(active) { const unsigned pos = count + __popc(ballot & ((1u << lane) - 1u)); list_C[pos] = C; list_g[pos] = g; }
                count += __popc(ballot);
            }
            if(lane == 0) count_sh = count;
        }
        const cplx lp = forward(ang0);                       // its barriers also publish the lists and count_sh
        const unsigned count = count_sh;
        for(unsigned idx0 = 0; idx0 < count; idx0 += DEEP_BLK_NC) {
            const unsigned idx = idx0 + c;
            const bool valid = idx < count;
            cplx ang = ang0;
            if(valid) {
                const unsigned g = list_g[idx];
                for(unsigned w = 0; w < op.words; w++) {
                    uint64_t m = op.flip[g * op.words + w];
                    while(m) {
                        const unsigned p = w * 64u + (unsigned)__ffsll((long long)m) - 1u;
                        const double delta = -2.0 * conf_spin(conf, p);
                        const cplx wv = ldg(&w1[(size_t)p * DEEP_BLK_W]);
                        ang.re = fma(delta, wv.re, ang.re); ang.im = fma(delta, wv.im, ang.im);
                        m &= m - 1ull;
                    }
                }
            }
            const cplx lp2 = forward(ang);
            if(valid && j == 0u) E += list_C[idx] * cexp(lp2 - lp);
        }
        // block sum of E
        E = warp_sum(E);
        if(lane == 0) esum[warp] = E;
        __syncthreads();
        if(tid == 0) {
            cplx t(0.0, 0.0);
            #pragma unroll
            for(int q = 0; q < DEEP_BLK_T / 32; q++) t += esum[q];
            eloc_out[s] = t;
        }
        __syncthreads();                                     // lists / esum are reused by the next sample
    }
}

#endif // __CUDACC__

} // namespace angpu

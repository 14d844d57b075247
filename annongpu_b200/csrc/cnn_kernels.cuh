// PsiCNN Metropolis sampler with INCREMENTAL forward passes.
//
// The reference (and the generic kernel k_mc<CnnDev>) recompute the whole network for every proposal
// (include/quantum_state/PsiCNN.hpp:177-181: "no fast update").  A single spin flip at site p only changes the outputs
// inside p's receptive cone: for 3x3 kernels on the 10x10 lattice 9 / 25 / 49 of the 100 sites of layers 1 / 2 / 3.
// This kernel keeps the activations of ALL layers of its chain resident in shared memory, recomputes in place only the
// affected outputs (lists precomputed on the host per flipped site), and restores them from a backup when the proposal
// is rejected.  Every recomputed output is evaluated from its full receptive field with the same operation order as the
// full forward pass, and log psi is re-summed over the whole last layer in the same order, so log psi -- and with it
// the Markov chain -- is BIT-IDENTICAL to the full recomputation; only the amount of arithmetic changes
// (C3: 6.2 k instead of 18.9 k complex MACs per proposal).
//
// One warp per chain, one lane per affected site with all output channels of that site in registers; the weights are
// staged once per block in shared memory in consumption order (CnnDev::stage).
#pragma once
#include "kernels.cuh"
#include "rbm_kernels.cuh"

namespace angpu {

struct CnnIncDev {
    const unsigned* aff[CNN_MAX_LAYERS];      // [N][aff_max[l]]: output sites of layer l affected by a flip of input site p
    const unsigned* aff_cnt[CNN_MAX_LAYERS];  // [N]
    unsigned        aff_max[CNN_MAX_LAYERS];
    unsigned        backup_elems;             // sum_l nch_l * aff_max[l]
};

// per-warp scratch: act[num_angles] | backup[backup_elems] (complex) | spin[N] (double, padded to 16 B)
__host__ __device__ inline size_t cnn_inc_slice_bytes(const CnnDev& psi, const CnnIncDev& inc) {
    return (size_t)(psi.num_angles + inc.backup_elems) * sizeof(cplx) + (((size_t)psi.N * sizeof(double) + 15u) & ~(size_t)15u);
}

#ifdef __CUDACC__

// Recompute in place, layer by layer, the outputs listed for entry `idx` (a flipped site for the sampler, a flip group for
// E_loc), keeping the old values in `backup`; `spin` already holds the flipped configuration.  All lanes of the warp.
__device__ __forceinline__ void cnn_cone_recompute(const CnnDev& psi, const CnnIncDev& inc, unsigned idx, const cplx* __restrict__ wgt,
                                                   const double* spin, cplx* act, cplx* backup) {
    const unsigned lane = threadIdx.x & 31u, N = psi.N;
    unsigned boff = 0;
    for(unsigned l = 0; l < psi.num_layers; l++) {
        const CnnLayerDev& ly = psi.L[l];
        const unsigned cnt = inc.aff_cnt[l][idx];
        const unsigned* list = inc.aff[l] + (size_t)idx * inc.aff_max[l];
        cplx* out = act + ly.angle_off;
        const cplx* in = l ? act + psi.L[l - 1u].angle_off : nullptr;
        for(unsigned k = lane; k < cnt; k += 32u) {
            const unsigned x = list[k];
            for(unsigned cj = 0; cj < ly.nch; cj++) backup[boff + cj * inc.aff_max[l] + k] = out[cj * N + x];
            psi.site_outputs_any(ly, l, wgt, in, spin, x, out, nullptr);
        }
        boff += ly.nch * inc.aff_max[l];
        __syncwarp();
    }
}
__device__ __forceinline__ void cnn_cone_restore(const CnnDev& psi, const CnnIncDev& inc, unsigned idx, cplx* act, const cplx* backup) {
    const unsigned lane = threadIdx.x & 31u, N = psi.N;
    unsigned boff = 0;
    for(unsigned l = 0; l < psi.num_layers; l++) {
        const CnnLayerDev& ly = psi.L[l];
        const unsigned cnt = inc.aff_cnt[l][idx];
        const unsigned* list = inc.aff[l] + (size_t)idx * inc.aff_max[l];
        cplx* out = act + ly.angle_off;
        for(unsigned k = lane; k < cnt; k += 32u) {
            const unsigned x = list[k];
            for(unsigned cj = 0; cj < ly.nch; cj++) out[cj * N + x] = backup[boff + cj * inc.aff_max[l] + k];
        }
        boff += ly.nch * inc.aff_max[l];
    }
    __syncwarp();
}
// full forward pass into the resident activations (forward_pass, PsiCNN.hpp:99-160)
__device__ __forceinline__ void cnn_full_forward(const CnnDev& psi, const cplx* __restrict__ wgt, const double* spin, cplx* act) {
    const unsigned lane = threadIdx.x & 31u;
    for(unsigned l = 0; l < psi.num_layers; l++) {
        const CnnLayerDev& ly = psi.L[l];
        const cplx* in = l ? act + psi.L[l - 1u].angle_off : nullptr;
        for(unsigned x = lane; x < psi.N; x += 32u) psi.site_outputs_any(ly, l, wgt, in, spin, x, act + ly.angle_off, nullptr);
        __syncwarp();
    }
}
// log psi from the resident last-layer activations, summed in the order of CnnDev::forward
__device__ __forceinline__ cplx cnn_log_psi_resident(const CnnDev& psi, const cplx* act) {
    const unsigned lane = threadIdx.x & 31u;
    const CnnLayerDev& last = psi.L[psi.num_layers - 1u];
    const cplx* o = act + last.angle_off;
    cplx r(0.0, 0.0);
    for(unsigned idx = lane; idx < last.nch * psi.N; idx += 32u) r += o[idx];
    const cplx total = warp_sum(r);
    __syncwarp();                       // the callers overwrite the activations next (restore / next proposal)
    return psi.lp + psi.final_factor * total;
}

__global__ void __launch_bounds__(128)
k_mc_cnn_inc(const CnnDev psi, const CnnIncDev inc, const McParams mc, uint64_t* __restrict__ conf_out,
             cplx* __restrict__ log_psi_out, unsigned long long* __restrict__ acc_rej) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned lane = threadIdx.x & 31u, wpb = blockDim.x >> 5, warp = threadIdx.x >> 5;
    const size_t slice = cnn_inc_slice_bytes(psi, inc);
    // block-shared staged weights live after the per-warp slices
    const cplx* __restrict__ wgt = reinterpret_cast<const cplx*>(psi.stage(smem_raw + (size_t)wpb * slice));
    cplx* act = reinterpret_cast<cplx*>(smem_raw + (size_t)warp * slice);
    cplx* backup = act + psi.num_angles;
    double* spin = reinterpret_cast<double*>(backup + inc.backup_elems);       // the configuration as +-1.0 (layer-0 input)
    const unsigned chain = blockIdx.x * wpb + warp;
    if(chain >= mc.num_chains_local) return;
    const unsigned gchain = mc.chain0 + chain, N = psi.N;
    const unsigned tag_init = (mc.call << 1) | 0u, tag_step = (mc.call << 1) | 1u;

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
    for(unsigned x = lane; x < N; x += 32u) spin[x] = conf_spin(conf, x);
    __syncwarp();
    cnn_full_forward(psi, wgt, spin, act);
    auto log_psi_now = [&]() -> cplx { return cnn_log_psi_resident(psi, act); };
    cplx cur = log_psi_now();

    const unsigned therm = mc.num_therm * N, per_sample = mc.num_sweeps * N;
    const unsigned long long total_steps = (unsigned long long)therm + (unsigned long long)per_sample * mc.steps_per_chain;
    unsigned long long acc = 0, next_record = (unsigned long long)therm + per_sample;
    unsigned sample = 0;

    for(unsigned long long t0 = 0; t0 < total_steps; t0 += 32u) {
        philox4x32_10((uint32_t)(t0 + lane), (uint32_t)((t0 + lane) >> 32), gchain, tag_step, mc.seed_lo, mc.seed_hi, r4);
        const unsigned my_site = r4[0] % N, my_ulo = r4[1], my_uhi = r4[2];
        const unsigned nb = (unsigned)min((unsigned long long)32u, total_steps - t0);
        for(unsigned b = 0; b < nb; b++) {
            const unsigned site = __shfl_sync(FULL, my_site, b);
            const double u = u01_from_bits(__shfl_sync(FULL, my_ulo, b), __shfl_sync(FULL, my_uhi, b));
            conf_flip(conf, site);
            if(lane == 0) spin[site] = -spin[site];
            __syncwarp();
            cnn_cone_recompute(psi, inc, site, wgt, spin, act, backup);      // the receptive cone of `site`, old values kept
            const cplx nlp = log_psi_now();
            if(metropolis_accept(2.0 * (nlp.re - cur.re), u)) {
                cur = nlp;
                acc++;
            } else {
                conf_flip(conf, site);
                if(lane == 0) spin[site] = -spin[site];
                cnn_cone_restore(psi, inc, site, act, backup);
            }
            if(t0 + b + 1u == next_record) {
                if(lane == 0) {
                    const size_t idx = (size_t)sample * mc.num_chains_local + chain;
                    log_psi_out[idx] = cur;
                    #pragma unroll
                    for(unsigned ww = 0; ww < (unsigned)MAXW; ww++) if(ww < psi.words) conf_out[idx * psi.words + ww] = conf[ww];
                }
                sample++; next_record += per_sample;
            }
        }
    }
    if(lane == 0) { atomicAdd(&acc_rej[0], acc); atomicAdd(&acc_rej[1], total_steps - acc); }
}

// E_loc for PsiCNN with the same machinery: one warp per sample keeps the activations of s resident, and every active
// flip group re-evaluates only the union of the receptive cones of its flipped sites (lists per group precomputed on the
// host, `inc` indexed by group), restoring afterwards.  psi(s')/psi(s) is bit-identical to two full forward passes
// (Operator.hpp:38-121 with PsiCNN.hpp:164-175).
__global__ void __launch_bounds__(128)
k_eloc_cnn_inc(const CnnDev psi, const CnnIncDev inc, const OpDev op, const uint64_t* __restrict__ confs, size_t ns,
               cplx* __restrict__ eloc_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned lane = threadIdx.x & 31u, wpb = blockDim.x >> 5, warp = threadIdx.x >> 5;
    const size_t slice = cnn_inc_slice_bytes(psi, inc);
    const cplx* __restrict__ wgt = reinterpret_cast<const cplx*>(psi.stage(smem_raw + (size_t)wpb * slice));
    cplx* act = reinterpret_cast<cplx*>(smem_raw + (size_t)warp * slice);
    cplx* backup = act + psi.num_angles;
    double* spin = reinterpret_cast<double*>(backup + inc.backup_elems);
    const unsigned N = psi.N, G = op.num_groups;
    for(size_t s = (size_t)blockIdx.x * wpb + warp; s < ns; s += (size_t)gridDim.x * wpb) {
        uint64_t conf[MAXW];
        conf_load(conf, confs + s * psi.words, psi.words);
        for(unsigned x = lane; x < N; x += 32u) spin[x] = conf_spin(conf, x);
        __syncwarp();
        cnn_full_forward(psi, wgt, spin, act);
        const cplx lp = cnn_log_psi_resident(psi, act);
        cplx E(0.0, 0.0);
        for(unsigned n = lane; n < op.num_diag; n += 32u) E += string_sign_reg(op, n, conf) * op.coef[n];
        E = warp_sum(E);
        for(unsigned g0 = 0; g0 < G; g0 += 32u) {
            const unsigned gl = g0 + lane;
            cplx C(0.0, 0.0);
            if(gl < G) C = strings_coefficient_reg(op, op.group_begin[gl], op.group_begin[gl + 1u], conf);
            unsigned active = __ballot_sync(FULL, C.re != 0.0 || C.im != 0.0);
            while(active) {                                      // warp-uniform loop over the active groups of this batch
                const unsigned src = (unsigned)__ffs((int)active) - 1u, g = g0 + src;
                active &= active - 1u;
                const cplx Cg(__shfl_sync(FULL, C.re, src), __shfl_sync(FULL, C.im, src));
                // flip the group's sites in the +-1 copy (lane w handles word w of the mask)
                if(lane < op.words) {
                    uint64_t m = op.flip[g * op.words + lane];
                    while(m) { const unsigned p = lane * 64u + (unsigned)__ffsll((long long)m) - 1u; spin[p] = -spin[p]; m &= m - 1ull; }
                }
                __syncwarp();
                cnn_cone_recompute(psi, inc, g, wgt, spin, act, backup);
                const cplx lp2 = cnn_log_psi_resident(psi, act);
                E += Cg * cexp(lp2 - lp);
                if(lane < op.words) {
                    uint64_t m = op.flip[g * op.words + lane];
                    while(m) { const unsigned p = lane * 64u + (unsigned)__ffsll((long long)m) - 1u; spin[p] = -spin[p]; m &= m - 1ull; }
                }
                cnn_cone_restore(psi, inc, g, act, backup);
            }
        }
        if(lane == 0) eloc_out[s] = E;
        __syncwarp();
    }
}

#endif // __CUDACC__

} // namespace angpu

// Model-generic VMC kernels: one warp per configuration / Markov chain, templated on the device model view
// (psi_dev.cuh).  These cover every model x ensemble combination; PsiRBM additionally has the
// register-resident fast paths in rbm_kernels.cuh.
//
// Pipeline (replaces the reference's "one kernel per ensemble.foreach with a fused consumer lambda and
// global atomics"):
//   sampler (k_mc / enumerate)  ->  conf[ns][words], log_psi[ns]
//   k_eloc                      ->  E_loc[ns]
//   k_ok                        ->  O[ns][P]           (dense sample x parameter matrix, HBM resident)
//   reductions.cu               ->  <E>, <|E|^2>, <O_k>, F, S, S.v   (deterministic two-stage, no float atomics)
#pragma once
#include "psi_dev.cuh"

namespace angpu {

struct McParams {
    unsigned long long num_samples;        // as given to the ensemble (weight = 1/num_samples)
    unsigned           num_sweeps, num_therm;
    unsigned           steps_per_chain;    // num_samples / num_chains (global, integer division)
    unsigned           num_chains_local;   // chains run by this process
    unsigned           chain0;             // global id of the first local chain
    unsigned           seed_lo, seed_hi, call;
    unsigned           pauli_sites;        // 0: spin basis; else the chains run over Pauli strings of that many sites (pauli_basis.cuh)
};

// per-warp shared-memory slice: [payload cplx x pl_elems][conf u64 x MAXW][conf' u64 x MAXW]
__host__ __device__ inline unsigned warp_slice_bytes(unsigned pl_elems) { return (pl_elems + 4u) * (unsigned)sizeof(cplx); }

#ifdef __CUDACC__

struct WarpScratch {
    cplx* pl; uint64_t* conf; uint64_t* conf2;
    __device__ WarpScratch(unsigned pl_elems) {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        unsigned char* base = smem_raw + (size_t)(threadIdx.x >> 5) * warp_slice_bytes(pl_elems);
        pl = reinterpret_cast<cplx*>(base);
        conf = reinterpret_cast<uint64_t*>(base + (size_t)pl_elems * sizeof(cplx));
        conf2 = conf + MAXW;
    }
};

// log psi of given configurations (psi_vector / log_psi_vector / psi_norm / ExactSummation weights;
// source/network_functions/PsiVector.cu.template:15-133, include/ensembles/ExactSummation.hpp:54-65).
// weight_out (optional) = exp(2 Re log psi), the un-normalised ExactSummation weight.
// block-level scratch (after the per-warp slices): used by models that stage parameters in shared memory
__device__ __forceinline__ unsigned char* block_scratch(unsigned pl_elems) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    return smem_raw + (size_t)(blockDim.x >> 5) * warp_slice_bytes(pl_elems);
}

template<class Psi>
__global__ void k_log_psi(const Psi psi, const uint64_t* __restrict__ confs, size_t ns,
                          cplx* __restrict__ log_psi_out, double* __restrict__ weight_out) {
    const unsigned char* blk = psi.stage(block_scratch(psi.payload_elems()));
    WarpScratch ws(psi.payload_elems());
    const unsigned lane = threadIdx.x & 31u, wpb = blockDim.x >> 5;
    for(size_t s = (size_t)blockIdx.x * wpb + (threadIdx.x >> 5); s < ns; s += (size_t)gridDim.x * wpb) {
        if(lane < psi.words) ws.conf[lane] = confs[s * psi.words + lane];
        __syncwarp();
        psi.init(ws.conf, ws.pl, blk);
        const cplx lp = psi.log_psi(ws.conf, ws.pl, blk);
        if(lane == 0) {
            log_psi_out[s] = lp;
            if(weight_out) weight_out[s] = exp(2.0 * lp.re);
        }
        __syncwarp();
    }
}

// Local energy E_loc(s) = sum_n c_n <s|P_n|s'> psi(s')/psi(s)  (include/operator/Operator.hpp:38-121),
// with strings grouped by flip mask (operator.hpp).
template<class Psi>
__global__ void k_eloc(const Psi psi, const OpDev op, const uint64_t* __restrict__ confs,
                       const cplx* __restrict__ log_psi, size_t ns, cplx* __restrict__ eloc_out) {
    const unsigned char* blk = psi.stage(block_scratch(psi.payload_elems()));
    WarpScratch ws(psi.payload_elems());
    const unsigned lane = threadIdx.x & 31u, wpb = blockDim.x >> 5;
    for(size_t s = (size_t)blockIdx.x * wpb + (threadIdx.x >> 5); s < ns; s += (size_t)gridDim.x * wpb) {
        if(lane < psi.words) ws.conf[lane] = confs[s * psi.words + lane];
        __syncwarp();
        psi.init(ws.conf, ws.pl, blk);
        const cplx lp = log_psi[s];
        cplx diag(0.0, 0.0);
        for(unsigned n = lane; n < op.num_diag; n += 32u) diag += string_sign(op, n, ws.conf) * op.coef[n];
        cplx E = warp_sum(diag);
        for(unsigned g = 0; g < op.num_groups; g++) {
            const cplx C = strings_coefficient(op, op.group_begin[g], op.group_begin[g + 1u], ws.conf);
            if(C.re == 0.0 && C.im == 0.0) continue;       // warp-uniform
            if(lane < psi.words) ws.conf2[lane] = ws.conf[lane] ^ op.flip[g * op.words + lane];
            __syncwarp();
            psi.update(ws.conf, ws.conf2, ws.pl, blk);
            const cplx lp2 = psi.log_psi(ws.conf2, ws.pl, blk);
            E += C * cexp(lp2 - lp);
            psi.update(ws.conf2, ws.conf, ws.pl, blk);
            __syncwarp();
        }
        if(lane == 0) eloc_out[s] = E;
        __syncwarp();
    }
}

// Dense log-derivative rows O[s][k] = d log psi(s) / d theta_k  (foreach_O_k of each model).
template<class Psi>
__global__ void k_ok(const Psi psi, const uint64_t* __restrict__ confs, size_t ns, cplx* __restrict__ O) {
    const unsigned char* blk = psi.stage(block_scratch(psi.payload_elems()));
    WarpScratch ws(psi.payload_elems());
    const unsigned lane = threadIdx.x & 31u, wpb = blockDim.x >> 5;
    for(size_t s = (size_t)blockIdx.x * wpb + (threadIdx.x >> 5); s < ns; s += (size_t)gridDim.x * wpb) {
        if(lane < psi.words) ws.conf[lane] = confs[s * psi.words + lane];
        __syncwarp();
        psi.init(ws.conf, ws.pl, blk);
        psi.O_k(ws.conf, ws.pl, O + s * (size_t)psi.P, blk);
        __syncwarp();
    }
}

// Single-spin-flip Metropolis, one warp per chain (MonteCarlo_t::kernel_foreach / mc_update,
// include/ensembles/MonteCarlo.hpp:57-177; Init_Policy.hpp:16-24; Update_Policy.hpp:20-28).
// Sample index = step * num_chains_local + chain (MonteCarlo.hpp:107-108).
template<class Psi>
__global__ void k_mc(const Psi psi, const McParams mc, uint64_t* __restrict__ conf_out,
                     cplx* __restrict__ log_psi_out, unsigned long long* __restrict__ acc_rej) {
    const unsigned char* blk = psi.stage(block_scratch(psi.payload_elems()));
    WarpScratch ws(psi.payload_elems());
    const unsigned lane = threadIdx.x & 31u, wpb = blockDim.x >> 5;
    const unsigned chain = blockIdx.x * wpb + (threadIdx.x >> 5);
    if(chain >= mc.num_chains_local) return;
    const unsigned gchain = mc.chain0 + chain;
    uint32_t r[4];
    if(lane < psi.words) {
        philox4x32_10(lane, 0u, gchain, (mc.call << 1) | 0u, mc.seed_lo, mc.seed_hi, r);
        uint64_t w = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
        if(lane == psi.words - 1u && (psi.N & 63u)) w &= (1ull << (psi.N & 63u)) - 1ull;
        ws.conf[lane] = w;
    }
    __syncwarp();
    psi.init(ws.conf, ws.pl, blk);
    cplx lp = psi.log_psi(ws.conf, ws.pl, blk);
    unsigned long long t = 0, acc = 0, rej = 0;
    const unsigned therm = mc.num_therm * psi.N, per_sample = mc.num_sweeps * psi.N;
    for(unsigned s = 0; s <= mc.steps_per_chain; s++) {
        const unsigned nsteps = (s == 0) ? therm : per_sample;
        for(unsigned i = 0; i < nsteps; i++, t++) {
            philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), gchain, (mc.call << 1) | 1u, mc.seed_lo, mc.seed_hi, r);
            const unsigned site = r[0] % psi.N;
            if(lane < psi.words) ws.conf2[lane] = ws.conf[lane] ^ ((lane == (site >> 6)) ? (1ull << (site & 63u)) : 0ull);
            __syncwarp();
            psi.update(ws.conf, ws.conf2, ws.pl, blk);
            const cplx nlp = psi.log_psi(ws.conf2, ws.pl, blk);
            const double ratio = exp(2.0 * (nlp.re - lp.re));
            const double u = u01_from_bits(r[1], r[2]);
            if(ratio > 1.0 || u <= ratio) {                 // warp-uniform (MonteCarlo.hpp:158-160)
                lp = nlp;
                if(lane < psi.words) ws.conf[lane] = ws.conf2[lane];
                acc++;
            } else {
                psi.update(ws.conf2, ws.conf, ws.pl, blk);
                rej++;
            }
            __syncwarp();
        }
        if(s == 0) continue;
        const size_t idx = (size_t)(s - 1u) * mc.num_chains_local + chain;
        if(lane < psi.words) conf_out[idx * psi.words + lane] = ws.conf[lane];
        if(lane == 0) log_psi_out[idx] = lp;
    }
    if(lane == 0) { atomicAdd(&acc_rej[0], acc); atomicAdd(&acc_rej[1], rej); }
}

#endif // __CUDACC__

} // namespace angpu

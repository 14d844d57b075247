// Pauli-string (density-matrix) basis: MonteCarloPaulis / ExactSummationPaulis and the local energy of a Pauli operator acting by
// Pauli multiplication (include/basis/PauliString.hpp:33-38, 41-65, 84-90, 257-277;
// include/ensembles/policies/{Init,Update}_Policy.hpp:30-44, 38-55; include/quantum_state/PsiDeep.hpp:282-308).
//
// A configuration is a Pauli string x = (a, b) over num_sites sites, I=(0,0) X=(1,0) Y=(0,1) Z=(1,1).  The network reads it through
// 3 num_sites input units, unit 3 s + t = +1 iff x[s] - 1 == t, else -1 (PauliString::network_unit_at) -- so a Pauli string IS a spin
// configuration of 3 num_sites units with at most one unit up per site, and a change of x[s] moves the angles by -2 w(old unit)
// + 2 w(new unit) (PsiDeep::update_angles for PauliString): exactly the spin-basis update of the two units that differ.  The
// configuration is therefore STORED as that units mask (SampleSet::conf, psi.words = words_for(3 num_sites)); log psi, the angle
// update, O_k, the reductions, S, S.v and the solvers are the spin-basis kernels unchanged.  What differs, and lives here:
//   k_enumerate_paulis   PauliString::enumerate: site s takes type (index >> 2 s) & 3
//   k_mc_paulis          Init_Policy (two random masks) and Update_Policy (site x % num_sites takes type x >> 30) of the chains
//   k_eloc_paulis        E_loc(x) = sum_n c_n f_n(x) psi(P_n x) / psi(x) with P x = f (P xor x); every string is its own "flip"
//                        (only the identity is diagonal), evaluated serially like the reference (Operator.hpp:38-121)
#pragma once
#include "kernels.cuh"

namespace angpu {

constexpr int PAULI_SITE_WORDS = 2;                 // 3 num_sites <= 64 MAXW  =>  num_sites <= 85

struct PauliOpDev {
    unsigned        num_strings, words;             // words of a SITE mask (1 or 2)
    const cplx*     coef;                           // [num_strings]  the caller's coefficients (no prefactor folded in)
    const uint64_t* a;                              // [num_strings][words]
    const uint64_t* b;
};

#ifdef __CUDACC__

__host__ __device__ __forceinline__ unsigned units_type(const uint64_t* units, unsigned s) {
    unsigned t = 0;
    #pragma unroll
    for(unsigned k = 0; k < 3u; k++) { const unsigned u = 3u * s + k; if((units[u >> 6] >> (u & 63u)) & 1ull) t = k + 1u; }
    return t;
}
__host__ __device__ __forceinline__ void units_set(uint64_t* units, unsigned s, unsigned type) {
    #pragma unroll
    for(unsigned k = 0; k < 3u; k++) { const unsigned u = 3u * s + k; units[u >> 6] &= ~(1ull << (u & 63u)); }
    if(type) { const unsigned u = 3u * s + type - 1u; units[u >> 6] |= 1ull << (u & 63u); }
}
// (a, b) site masks of the units mask in `units` (shared memory), identical on every lane
__device__ __forceinline__ void units_to_ab(const uint64_t* units, unsigned num_sites, uint64_t (&xa)[PAULI_SITE_WORDS], uint64_t (&xb)[PAULI_SITE_WORDS]) {
    const unsigned lane = threadIdx.x & 31u;
    #pragma unroll
    for(int w = 0; w < PAULI_SITE_WORDS; w++) { xa[w] = 0ull; xb[w] = 0ull; }
    #pragma unroll
    for(unsigned k = 0; k < 2u * PAULI_SITE_WORDS; k++) {
        const unsigned s = 32u * k + lane;
        const unsigned t = (s < num_sites) ? units_type(units, s) : 0u;
        const uint64_t ba = __ballot_sync(FULL, t & 1u), bb = __ballot_sync(FULL, t & 2u);
        xa[k >> 1] |= ba << (32u * (k & 1u)); xb[k >> 1] |= bb << (32u * (k & 1u));
    }
}

static __global__ void k_enumerate_paulis(uint64_t* __restrict__ conf, size_t begin, size_t n, unsigned num_sites, unsigned words) {
    for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t u[MAXW] = {0ull, 0ull, 0ull, 0ull};
        const uint64_t index = (uint64_t)(begin + i);
        for(unsigned s = 0; s < num_sites && s < 32u; s++) units_set(u, s, (unsigned)((index >> (2u * s)) & 3ull));
        for(unsigned w = 0; w < words; w++) conf[i * words + w] = u[w];
    }
}

// One warp per Markov chain over Pauli strings (MonteCarlo_t<PauliString>::kernel_foreach / mc_update, MonteCarlo.hpp:57-177):
// a sweep is psi.N = 3 num_sites proposals (get_num_input_units), a proposal redraws the type of one site -- possibly the type it
// has, which is then accepted with ratio 1 as in the reference.
template<class Psi>
__global__ void k_mc_paulis(const Psi psi, const McParams mc, uint64_t* __restrict__ conf_out,
                            cplx* __restrict__ log_psi_out, unsigned long long* __restrict__ acc_rej) {
    const unsigned char* blk = psi.stage(block_scratch(psi.payload_elems()));
    WarpScratch ws(psi.payload_elems());
    const unsigned lane = threadIdx.x & 31u, wpb = blockDim.x >> 5;
    const unsigned chain = blockIdx.x * wpb + (threadIdx.x >> 5);
    if(chain >= mc.num_chains_local) return;
    const unsigned gchain = mc.chain0 + chain, num_sites = mc.pauli_sites;
    uint32_t r[4];
    if(lane == 0) {
        // PauliString::set_randomly (PauliString.hpp:59-65): both masks cut with (1 << (num_sites % 64)) - 1, which is 0 for 64 sites
        uint64_t u[MAXW] = {0ull, 0ull, 0ull, 0ull};
        const unsigned sw = (num_sites + 63u) / 64u;
        const uint64_t cut = (1ull << (num_sites & 63u)) - 1ull;
        for(unsigned w = 0; w < sw; w++) {
            philox4x32_10(w, 0u, gchain, (mc.call << 1) | 0u, mc.seed_lo, mc.seed_hi, r);
            uint64_t a = (uint64_t)r[0] | ((uint64_t)r[1] << 32), b = (uint64_t)r[2] | ((uint64_t)r[3] << 32);
            if(w == sw - 1u) { a &= cut; b &= cut; }
            for(unsigned q = 0; q < 64u && 64u * w + q < num_sites; q++)
                units_set(u, 64u * w + q, (unsigned)((a >> q) & 1ull) | ((unsigned)((b >> q) & 1ull) << 1));
        }
        for(unsigned w = 0; w < psi.words; w++) ws.conf[w] = u[w];
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
            if(lane < psi.words) ws.conf2[lane] = ws.conf[lane];
            __syncwarp();
            if(lane == 0) units_set(ws.conf2, r[0] % num_sites, r[0] >> 30);
            __syncwarp();
            psi.update(ws.conf, ws.conf2, ws.pl, blk);
            const cplx nlp = psi.log_psi(ws.conf2, ws.pl, blk);
            const double ratio = exp(2.0 * (nlp.re - lp.re));
            const double u = u01_from_bits(r[1], r[2]);
            if(ratio > 1.0 || u <= ratio) {
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

template<class Psi>
__global__ void k_eloc_paulis(const Psi psi, const PauliOpDev op, unsigned num_sites, const uint64_t* __restrict__ confs,
                              const cplx* __restrict__ log_psi, size_t ns, cplx* __restrict__ eloc_out) {
    const unsigned char* blk = psi.stage(block_scratch(psi.payload_elems()));
    WarpScratch ws(psi.payload_elems());
    const unsigned lane = threadIdx.x & 31u, wpb = blockDim.x >> 5;
    for(size_t s = (size_t)blockIdx.x * wpb + (threadIdx.x >> 5); s < ns; s += (size_t)gridDim.x * wpb) {
        if(lane < psi.words) ws.conf[lane] = confs[s * psi.words + lane];
        __syncwarp();
        psi.init(ws.conf, ws.pl, blk);
        const cplx lp = log_psi[s];
        uint64_t xa[PAULI_SITE_WORDS], xb[PAULI_SITE_WORDS];
        units_to_ab(ws.conf, num_sites, xa, xb);
        cplx E(0.0, 0.0);
        for(unsigned n = 0; n < op.num_strings; n++) {
            // P x = factor (P xor x), PauliString::apply(PauliString) (PauliString.hpp:257-277); warp-uniform
            unsigned nneg = 0, neps = 0; uint64_t any = 0ull;
            #pragma unroll
            for(unsigned w = 0; w < (unsigned)PAULI_SITE_WORDS; w++) {
                if(w < op.words) {
                    const uint64_t pa = op.a[n * op.words + w], pb = op.b[n * op.words + w];
                    const uint64_t px = pa & ~pb, py = ~pa & pb, pz = pa & pb;
                    const uint64_t xx = xa[w] & ~xb[w], xy = ~xa[w] & xb[w], xz = xa[w] & xb[w];
                    nneg += __popcll((px & xz) | (py & xx) | (pz & xy));
                    neps += __popcll((pa | pb) & (xa[w] | xb[w]) & ((pa ^ xa[w]) | (pb ^ xb[w])));
                    any |= pa | pb;
                }
            }
            cplx C = op.coef[n];
            if(nneg & 1u) C = -C;
            if((neps & 3u) > 1u) C = -C;
            if(neps & 1u) C = C * cplx(0.0, -1.0);
            if(any == 0ull) { E += C; continue; }                       // the identity: the only diagonal string
            if(lane < psi.words) ws.conf2[lane] = ws.conf[lane];
            __syncwarp();
            if(lane == 0) {
                for(unsigned w = 0; w < op.words; w++) {
                    const uint64_t pa = op.a[n * op.words + w], pb = op.b[n * op.words + w];
                    uint64_t m = pa | pb;
                    while(m) {
                        const unsigned q = (unsigned)__ffsll((long long)m) - 1u;
                        const uint64_t na = pa ^ xa[w], nb = pb ^ xb[w];
                        units_set(ws.conf2, 64u * w + q, (unsigned)((na >> q) & 1ull) | ((unsigned)((nb >> q) & 1ull) << 1));
                        m &= m - 1ull;
                    }
                }
            }
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

#endif // __CUDACC__

} // namespace angpu

// Pauli-string operator, device layout.
//
// The reference keeps an SoA of (coefficient, PauliString{a,b}) and walks it strictly serially, one
// psi'/psi evaluation per off-diagonal string (include/operator/Operator.hpp:38-121).  Here the strings
// are pre-processed once on the host:
//   * the configuration-independent prefactor (-i)^{n_Y} (PauliString.hpp:193-208) is folded into the
//     coefficient; what remains per configuration is the sign (-1)^{popc(~s & b)} (PauliString.hpp:247-249);
//   * strings are grouped by flip mask f = a ^ b (PauliString.hpp:251): all strings of a group share
//     s' = s ^ f and therefore ONE psi(s')/psi(s) evaluation (XX and YY on a bond; for aligned spins their
//     coefficients cancel exactly and the evaluation is skipped);
//   * masks are multi-word (N <= 256) — the reference is limited to 64 sites.
#pragma once
#include "runtime.hpp"
#include <map>
#include <array>

namespace angpu {

struct OpDev {
    unsigned        num_strings;      // all strings, sorted: diagonal ones first, then by group
    unsigned        num_diag;         // strings [0, num_diag) have f == 0
    unsigned        num_groups;       // off-diagonal groups
    unsigned        words;
    unsigned        max_flips;        // max popcount of a flip mask
    const cplx*     coef;             // [num_strings]  c_n * (-i)^{n_Y}
    const uint64_t* b;                // [num_strings][words]  sign mask (Y|Z sites)
    const uint64_t* flip;             // [num_groups][words]
    const unsigned* group_begin;      // [num_groups + 1] into the sorted string list
};

struct Operator {
    unsigned num_strings = 0, words = 1;
    // caller's original order (kept for copies / introspection)
    std::vector<cplx> h_coef; std::vector<uint64_t> h_a, h_b;
    std::vector<uint64_t> h_flip;          // [num_groups][words]: the flip mask of every off-diagonal group (device order)
    unsigned num_sites_touched = 0;        // 1 + the highest site any string acts on (0 for a pure identity)
    DevBuf<cplx> d_coef; DevBuf<uint64_t> d_b, d_flip; DevBuf<unsigned> d_group_begin;
    OpDev dev{};
    // the strings as given (caller's order, nothing folded in): the layout of the Pauli-string basis, where a string acts by Pauli
    // multiplication (pauli_basis.cuh)
    DevBuf<cplx> d_pcoef; DevBuf<uint64_t> d_pa, d_pb;

    Operator(unsigned n, const double* coeffs, const uint64_t* a, const uint64_t* b, unsigned words_) : num_strings(n), words(words_) {
        ANGPU_REQUIRE(words >= 1 && words <= (unsigned)MAXW, "operator: words must be in 1..4");
        h_coef.resize(n); h_a.assign(a, a + (size_t)n * words); h_b.assign(b, b + (size_t)n * words);
        for(unsigned i = 0; i < n; i++) h_coef[i] = cplx(coeffs[2 * i], coeffs[2 * i + 1]);
        build();
    }
    Operator(const Operator& o) : num_strings(o.num_strings), words(o.words), h_coef(o.h_coef), h_a(o.h_a), h_b(o.h_b) { build(); }

    void build() {
        using Mask = std::array<uint64_t, MAXW>;
        std::map<Mask, std::vector<unsigned>> groups;   // ordered => deterministic layout
        for(unsigned i = 0; i < num_strings; i++) {
            Mask f{}; for(unsigned w = 0; w < words; w++) f[w] = h_a[i * words + w] ^ h_b[i * words + w];
            groups[f].push_back(i);
        }
        std::vector<cplx> coef; std::vector<uint64_t> bmask, flip; std::vector<unsigned> begin;
        auto push_string = [&](unsigned i) {
            unsigned ny = 0;
            for(unsigned w = 0; w < words; w++) ny += (unsigned)__builtin_popcountll(~h_a[i * words + w] & h_b[i * words + w]);
            cplx f(1.0, 0.0);
            if((ny & 3u) > 1u) f = -f;
            if(ny & 1u) f = f * cplx(0.0, -1.0);
            coef.push_back(h_coef[i] * f);
            for(unsigned w = 0; w < words; w++) bmask.push_back(h_b[i * words + w]);
        };
        const Mask zero{};
        unsigned num_diag = 0, max_flips = 0;
        auto it0 = groups.find(zero);
        if(it0 != groups.end()) { for(unsigned i : it0->second) push_string(i); num_diag = (unsigned)it0->second.size(); }
        unsigned ng = 0;
        for(auto& kv : groups) {
            if(kv.first == zero) continue;
            begin.push_back((unsigned)coef.size());
            unsigned pc = 0;
            for(unsigned w = 0; w < words; w++) { flip.push_back(kv.first[w]); pc += (unsigned)__builtin_popcountll(kv.first[w]); }
            if(pc > max_flips) max_flips = pc;
            for(unsigned i : kv.second) push_string(i);
            ng++;
        }
        begin.push_back((unsigned)coef.size());
        h_flip = flip;
        num_sites_touched = 0;
        for(unsigned i = 0; i < num_strings; i++)
            for(unsigned w = 0; w < words; w++) {
                const uint64_t m = h_a[i * words + w] | h_b[i * words + w];
                if(m) num_sites_touched = std::max(num_sites_touched, w * 64u + 64u - (unsigned)__builtin_clzll(m));
            }
        d_coef.upload(coef); d_b.upload(bmask); d_flip.upload(flip); d_group_begin.upload(begin);
        d_pcoef.upload(h_coef); d_pa.upload(h_a); d_pb.upload(h_b);
        dev = OpDev{num_strings, num_diag, ng, words, max_flips, d_coef.p, d_b.p, d_flip.p, d_group_begin.p};
    }
};

// every kernel indexes the wavefunction by the sites an operator touches: reject operators wider than the state
inline void require_operator_fits(const Operator& op, unsigned num_sites, unsigned words) {
    ANGPU_REQUIRE(op.words == words, "operator / wavefunction word count mismatch");
    ANGPU_REQUIRE(op.num_sites_touched <= num_sites, "operator acts on site " + std::to_string(op.num_sites_touched - 1) + " but the wavefunction has " + std::to_string(num_sites) + " sites");
}

#ifdef __CUDACC__
// sign (-1)^{popc(~s & b)} of string n on configuration s (uniform over the warp)
__device__ __forceinline__ double string_sign(const OpDev& op, unsigned n, const uint64_t* conf) {
    unsigned pc = 0;
    for(unsigned w = 0; w < op.words; w++) pc += __popcll(~conf[w] & op.b[n * op.words + w]);
    return (pc & 1u) ? -1.0 : 1.0;
}
// sum over strings [lo, hi) of coef * sign — serial (ranges are short)
__device__ __forceinline__ cplx strings_coefficient(const OpDev& op, unsigned lo, unsigned hi, const uint64_t* conf) {
    cplx c(0.0, 0.0);
    for(unsigned n = lo; n < hi; n++) c += string_sign(op, n, conf) * op.coef[n];
    return c;
}
// StandartOperator::fast_local_energy (include/operator/Operator.hpp:125-136): sum over ALL strings of
// coefficient * apply(conf).coefficient, evaluated cooperatively by the warp.
__device__ __forceinline__ cplx fast_local_energy_warp(const OpDev& op, const uint64_t* conf) {
    cplx c(0.0, 0.0);
    for(unsigned n = threadIdx.x & 31u; n < op.num_strings; n += 32u) c += string_sign(op, n, conf) * op.coef[n];
    return warp_sum(c);
}
__device__ __forceinline__ cplx fast_local_energy_serial(const OpDev& op, const uint64_t* conf) {
    return strings_coefficient(op, 0u, op.num_strings, conf);
}
#endif

} // namespace angpu

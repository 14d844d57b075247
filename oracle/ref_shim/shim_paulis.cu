// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C interface over the UNMODIFIED reference compiled with -DENABLE_PAULIS (the Pauli-string / density-matrix basis,
// SURVEY.md §8f rank 3) and PsiDeep only: PsiRBM does not compile with that basis (PsiRBM.hpp:172), so this is a second
// library (oracle/_ref/liboracle_ref_paulis.so) next to the Spins one.  Host path (gpu=false) only.
// This file contains no reference source: it only *calls* the reference's public C++ API
//   PauliString::{enumerate, apply(PauliString), network_unit_at}     (include/basis/PauliString.hpp:33-38, 84-90, 257-277)
//   PsiDeep with N = 3 num_sites input units                          (include/quantum_state/PsiDeep.hpp:282-308)
//   ExactSummationPaulis / MonteCarloPaulis                           (include/ensembles/*.hpp)
//   ExpectationValue, TDVP, log_psi_s, psi_O_k, psi_vector            (include/network_functions/*.hpp)

#define __PYTHONCC__

#include "network_functions/ExpectationValue.hpp"
#include "network_functions/TDVP.hpp"
#include "network_functions/PsiVector.hpp"
#include "network_functions/PsiOkVector.hpp"
#include "quantum_states.hpp"
#include "ensembles.hpp"
#include "operators.hpp"
#include "bases.hpp"
#include "types.h"

#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>

using namespace ann_on_gpu;
using cplx = std::complex<double>;

struct RawExpr {
    unsigned int     n;
    const double*    coeffs;   // interleaved
    const uint64_t*  a;
    const uint64_t*  b;
};

namespace ann_on_gpu {

// Raw-array replacement for the QuantumExpression-consuming constructor (source/operator/Operator.cpp:18-39)
template<>
template<>
StandartOperator<PauliString>::StandartOperator(const RawExpr& expr, const bool gpu)
    : gpu(gpu), coefficients(expr.n, gpu), quantum_strings(expr.n, gpu)
{
    for(auto i = 0u; i < expr.n; i++) {
        this->coefficients[i] = complex_t(expr.coeffs[2 * i], expr.coeffs[2 * i + 1]);
        this->quantum_strings[i] = PauliString(expr.a[i], expr.b[i]);
    }
    this->coefficients.update_device();
    this->quantum_strings.update_device();
    this->kernel().num_strings = expr.n;
    this->kernel().coefficients = this->coefficients.data();
    this->kernel().quantum_strings = this->quantum_strings.data();
}

} // namespace ann_on_gpu

namespace {
enum EnsKind { ES = 0, MC = 1 };
inline xt::pytensor<cplx, 1> ctensor1(const double* src, long n) { return xt::pytensor<cplx, 1>(reinterpret_cast<const cplx*>(src), {n}); }
inline xt::pytensor<cplx, 2> ctensor2(const double* src, long n, long m) { return xt::pytensor<cplx, 2>(reinterpret_cast<const cplx*>(src), {n, m}); }
template<typename F> void with_ens(int kind, void* h, F f) {
    if(kind == ES) f(*static_cast<ExactSummationPaulis*>(h));
    else           f(*static_cast<MonteCarloPaulis*>(h));
}
inline void store(double* out, const complex_t& z) { out[0] = z.real(); out[1] = z.imag(); }
void copy_out(double* out, const Array<complex_t>& ar) { std::memcpy(out, ar.host_data(), sizeof(complex_t) * ar.size()); }
} // namespace

extern "C" {

// PauliString::enumerate  (PauliString.hpp:33-38)
void refp_enumerate(unsigned int index, uint64_t* a_out, uint64_t* b_out) {
    const auto p = PauliString::enumerate(index); *a_out = p.a; *b_out = p.b;
}
// PauliString::apply(PauliString)  (PauliString.hpp:257-277)
void refp_pauli_mul(uint64_t a, uint64_t b, uint64_t xa, uint64_t xb, double* coeff_out, uint64_t* a_out, uint64_t* b_out) {
    const auto me = PauliString(a, b).apply(PauliString(xa, xb));
    store(coeff_out, me.coefficient); *a_out = me.vector.a; *b_out = me.vector.b;
}
// PauliString::network_unit_at  (PauliString.hpp:84-90)
int refp_network_unit_at(uint64_t a, uint64_t b, unsigned int idx) { return PauliString(a, b).network_unit_at(idx); }

void* refp_op_create(unsigned int n, const double* coeffs, const uint64_t* a, const uint64_t* b) {
    RawExpr e{n, coeffs, a, b};
    return new Operator(e, false);
}
void refp_op_destroy(void* op) { delete static_cast<Operator*>(op); }

void* refp_deep_create(unsigned int num_sites, unsigned int N, const double* input_weights,
                       unsigned int num_hidden, const unsigned int* sizes, const unsigned int* conn,
                       const double* biases, const unsigned int* lhs_connections, const double* lhs_weights,
                       const double* final_weights, double lp_re, double lp_im) {
    std::vector<xt::pytensor<cplx, 1>> b_list;
    std::vector<xt::pytensor<unsigned int, 2>> c_list;
    std::vector<xt::pytensor<cplx, 2>> w_list;
    size_t off_b = 0, off_w = 0;
    for(auto l = 0u; l < num_hidden; l++) {
        b_list.push_back(ctensor1(biases + 2 * off_b, sizes[l]));
        c_list.push_back(xt::pytensor<unsigned int, 2>(lhs_connections + off_w, {(long)conn[l], (long)sizes[l]}));
        w_list.push_back(ctensor2(lhs_weights + 2 * off_w, conn[l], sizes[l]));
        off_b += sizes[l];
        off_w += size_t(conn[l]) * sizes[l];
    }
    return new PsiDeep(num_sites, ctensor1(input_weights, N), b_list, c_list, w_list,
                       ctensor1(final_weights, sizes[num_hidden - 1]), cplx(lp_re, lp_im), false);
}
void refp_psi_destroy(void* h) { delete static_cast<PsiDeep*>(h); }
unsigned int refp_psi_num_params(void* h) { return static_cast<PsiDeep*>(h)->num_params; }

void* refp_es_create(unsigned int num_sites) { return new ExactSummationPaulis(num_sites, false); }
void* refp_mc_create(unsigned int num_samples, unsigned int num_sweeps, unsigned int num_therm, unsigned int num_chains) {
    return new MonteCarloPaulis(num_samples, num_sweeps, num_therm, num_chains, Update_Policy<PauliString>(), false);
}
void refp_ens_destroy(int kind, void* h) { with_ens(kind, h, [](auto& e) { delete &e; }); }
unsigned int refp_ens_num_steps(int kind, void* h) { unsigned int n = 0; with_ens(kind, h, [&](auto& e) { n = e.get_num_steps(); }); return n; }

void refp_log_psi_s(void* h, uint64_t a, uint64_t b, double* out) { store(out, log_psi_s(*static_cast<PsiDeep*>(h), PauliString(a, b))); }
void refp_psi_O_k(void* h, uint64_t a, uint64_t b, double* out) { copy_out(out, psi_O_k(*static_cast<PsiDeep*>(h), PauliString(a, b))); }
void refp_log_psi_vector(void* h, int ek, void* e, double* out) {
    with_ens(ek, e, [&](auto& ens) { copy_out(out, log_psi_vector(*static_cast<PsiDeep*>(h), ens)); });
}

void refp_expectation(void* h, void* op, int ek, void* e, double* out) {
    ExpectationValue ev(false);
    with_ens(ek, e, [&](auto& ens) { store(out, ev(*static_cast<Operator*>(op), *static_cast<PsiDeep*>(h), ens)); });
}
// out = {fluctuation, Re <A>, Im <A>}
void refp_fluctuation(void* h, void* op, int ek, void* e, double* out) {
    ExpectationValue ev(false);
    with_ens(ek, e, [&](auto& ens) {
        const auto r = ev.fluctuation(*static_cast<Operator*>(op), *static_cast<PsiDeep*>(h), ens);
        out[0] = r.first; store(out + 1, r.second);
    });
}
void refp_gradient(void* h, void* op, int ek, void* e, double* grad_out, double* E_out) {
    ExpectationValue ev(false);
    with_ens(ek, e, [&](auto& ens) {
        const auto r = ev.gradient(*static_cast<Operator*>(op), *static_cast<PsiDeep*>(h), ens);
        copy_out(grad_out, r.first); store(E_out, r.second);
    });
}

void* refp_tdvp_create(unsigned int num_params) { return new TDVP(num_params, false); }
void  refp_tdvp_destroy(void* t) { delete static_cast<TDVP*>(t); }
void refp_tdvp_eval(void* t, void* h, void* op, int ek, void* e) {
    with_ens(ek, e, [&](auto& ens) { static_cast<TDVP*>(t)->eval(*static_cast<Operator*>(op), *static_cast<PsiDeep*>(h), ens, false_t()); });
}
// scal_out = {Re E, Im E, E2, var_H}; any output pointer may be null.
void refp_tdvp_get(void* t, double* S_out, double* F_out, double* Ok_out, double* scal_out) {
    auto& tdvp = *static_cast<TDVP*>(t);
    if(S_out)  copy_out(S_out, tdvp.S_matrix);
    if(F_out)  copy_out(F_out, tdvp.F_vector);
    if(Ok_out) copy_out(Ok_out, tdvp.O_k_ar);
    if(scal_out) { store(scal_out, tdvp.E_local.front()); scal_out[2] = tdvp.E2_local.front(); scal_out[3] = tdvp.var_H(); }
}

} // extern "C"

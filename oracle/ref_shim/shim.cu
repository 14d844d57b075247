// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C interface over the UNMODIFIED reference (heikoburau/ANNonGPU) compiled from the sources
// where they lie under /root/reference (see oracle/Makefile).  By default only the reference's
// serial host path (gpu=false) is exercised: that path makes no CUDA calls and is the bit-for-bit
// CPU restatement of the reference's own arithmetic (include/cuda_kernel_defines.h:16-29).
// ref_set_gpu(1) switches every object created afterwards to the reference's own CUDA path
// (gpu=true): used ONLY by oracle/ref_gpu_bench.py to time the reference's kernels on the same B200
// as a second baseline.
//
// This file contains no reference source: it only *calls* the reference's public C++ API
//   PsiRBM / PsiDeep / PsiCNN / PsiClassicalFP / PsiClassicalANN   (include/quantum_state/*.hpp)
//   Operator                                                    (include/operator/Operator.hpp:163-203)
//   ExactSummationSpins / MonteCarloSpins                         (include/ensembles/*.hpp)
//   ExpectationValue, TDVP, log_psi_s, psi_O_k, psi_vector, ...   (include/network_functions/*.hpp)
// All arrays are caller-owned host pointers; complex numbers are interleaved (re, im) doubles.
//
// __PYTHONCC__ is defined for this TU only: it unlocks the general constructors that the
// reference keeps under that macro (PsiRBM.hpp:223-245, PsiDeep.hpp:502-618, PsiCNN.hpp:306-342,
// PsiClassical.hpp:192-214); class layouts do not depend on the macro.  xt::pytensor is the
// stub in oracle/ref_shim/stubs/.

#define __PYTHONCC__

#include "network_functions/ExpectationValue.hpp"
#include "network_functions/TDVP.hpp"
#include "network_functions/PsiVector.hpp"
#include "network_functions/PsiNorm.hpp"
#include "network_functions/PsiOkVector.hpp"
#include "network_functions/ApplyOperator.hpp"
#include "network_functions/HilbertSpaceDistance.hpp"
#include "network_functions/KullbackLeibler.hpp"
#include "quantum_states.hpp"
#include "quantum_state/psi_functions.hpp"
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

static bool g_gpu = false;    // see ref_set_gpu

struct RawExpr {
    unsigned int     n;
    const double*    coeffs;   // interleaved
    const uint64_t*  a;
    const uint64_t*  b;
};

namespace ann_on_gpu {

// Raw-array replacement for the QuantumExpression-consuming constructor
// (source/operator/Operator.cpp:18-39, which only reads term.first.a/.b and term.second).
template<>
template<>
StandartOperator<PauliString>::StandartOperator(const RawExpr& expr, const bool gpu)
    : gpu(gpu), coefficients(expr.n, gpu), quantum_strings(expr.n, gpu)
{
    for(auto i = 0u; i < expr.n; i++) {
        this->coefficients[i] = complex_t(expr.coeffs[2 * i], expr.coeffs[2 * i + 1]);
        this->quantum_strings[i] = PauliString(expr.a[i], expr.b[i]);
    }
    this->coefficients.update_device();       // no-ops for gpu == false
    this->quantum_strings.update_device();
    this->kernel().num_strings = expr.n;
    this->kernel().coefficients = this->coefficients.data();
    this->kernel().quantum_strings = this->quantum_strings.data();
}

} // namespace ann_on_gpu


namespace {

enum PsiKind { RBM = 0, DEEP = 1, CNN = 2, CLFP1 = 3, CLFP2 = 4, CLANN1 = 5, CLANN2 = 6 };
enum EnsKind { ES = 0, MC = 1 };

template<typename T, size_t dim>
xt::pytensor<T, dim> tensor_from(const T* src, std::array<long int, dim> shape) {
    return xt::pytensor<T, dim>(src, shape);
}

inline xt::pytensor<cplx, 1> ctensor1(const double* src, long n) {
    return xt::pytensor<cplx, 1>(reinterpret_cast<const cplx*>(src), {n});
}
inline xt::pytensor<cplx, 2> ctensor2(const double* src, long n, long m) {
    return xt::pytensor<cplx, 2>(reinterpret_cast<const cplx*>(src), {n, m});
}

template<typename F>
void with_psi(int kind, void* h, F f) {
    switch(kind) {
        case RBM:    f(*static_cast<PsiRBM*>(h)); break;
        case DEEP:   f(*static_cast<PsiDeep*>(h)); break;
        case CNN:    f(*static_cast<PsiCNN*>(h)); break;
        case CLFP1:  f(*static_cast<PsiClassicalFP<1u>*>(h)); break;
        case CLFP2:  f(*static_cast<PsiClassicalFP<2u>*>(h)); break;
        case CLANN1: f(*static_cast<PsiClassicalANN<1u>*>(h)); break;
        case CLANN2: f(*static_cast<PsiClassicalANN<2u>*>(h)); break;
    }
}

template<typename F>
void with_classical(int kind, void* h, F f) {
    switch(kind) {
        case CLFP1:  f(*static_cast<PsiClassicalFP<1u>*>(h)); break;
        case CLFP2:  f(*static_cast<PsiClassicalFP<2u>*>(h)); break;
        case CLANN1: f(*static_cast<PsiClassicalANN<1u>*>(h)); break;
        case CLANN2: f(*static_cast<PsiClassicalANN<2u>*>(h)); break;
    }
}

template<typename F>
void with_ens(int kind, void* h, F f) {
    if(kind == ES) f(*static_cast<ExactSummationSpins*>(h));
    else           f(*static_cast<MonteCarloSpins*>(h));
}

inline void store(double* out, const complex_t& z) { out[0] = z.real(); out[1] = z.imag(); }
inline void store(double* out, const cplx& z) { out[0] = z.real(); out[1] = z.imag(); }

void copy_out(double* out, const Array<complex_t>& ar) {
    std::memcpy(out, ar.host_data(), sizeof(complex_t) * ar.size());
}

} // namespace


extern "C" {

// ---------------------------------------------------------------- basis / operator primitives

// PauliString::apply(Spins)  (include/basis/PauliString.hpp:242-255)
void ref_pauli_apply(uint64_t a, uint64_t b, uint64_t conf, double* coeff_out, uint64_t* conf_out) {
    const auto me = PauliString(a, b).apply(Spins(conf, 64u));
    store(coeff_out, me.coefficient);
    *conf_out = me.vector.configuration();
}

// Spins::enumerate / operator[]  (include/basis/Spins.h:291-296, 104-108)
uint64_t ref_spins_enumerate(unsigned int index) { return Spins::enumerate(index).configuration(); }
double   ref_spins_at(uint64_t conf, unsigned int i) { return Spins(conf, 64u)[i]; }

// my_logcosh / my_tanh  (include/quantum_state/psi_functions.hpp:11-49, 80-116)
void ref_activation(double re, double im, unsigned int layer, double* lc_out, double* th_out) {
    store(lc_out, my_logcosh(complex_t(re, im), layer));
    store(th_out, my_tanh(complex_t(re, im), layer));
}

void* ref_op_create(unsigned int n, const double* coeffs, const uint64_t* a, const uint64_t* b) {
    RawExpr e{n, coeffs, a, b};
    return new Operator(e, g_gpu);
}
void ref_op_destroy(void* op) { delete static_cast<Operator*>(op); }

// ---------------------------------------------------------------- wavefunctions

void* ref_rbm_create(unsigned int N, unsigned int M, const double* W,
                     double fw_re, double fw_im, double lp_re, double lp_im) {
    return new PsiRBM(ctensor2(W, N, M), cplx(fw_re, fw_im), cplx(lp_re, lp_im), g_gpu);
}

// hidden layer l (0-based): sizes[l] units, conn[l] lhs-connections per unit;
// biases / lhs_connections / lhs_weights are the concatenation over layers.
void* ref_deep_create(unsigned int num_sites, unsigned int N, const double* input_weights,
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
    return new PsiDeep(
        num_sites, ctensor1(input_weights, N), b_list, c_list, w_list,
        ctensor1(final_weights, sizes[num_hidden - 1]), cplx(lp_re, lp_im), g_gpu
    );
}

void* ref_cnn_create(const unsigned int* extent, unsigned int num_layers, const unsigned int* num_channels,
                     const unsigned int* connectivity, const unsigned int* symmetry_classes,
                     const double* params, unsigned int num_params,
                     double final_factor, double lp_re, double lp_im) {
    const std::array<unsigned int, 3> ext{extent[0], extent[1], extent[2]};
    const long N = long(extent[0]) * extent[1] * extent[2];
    return new PsiCNN(
        ext,
        xt::pytensor<unsigned int, 1>(num_channels, {(long)num_layers}),
        xt::pytensor<unsigned int, 2>(connectivity, {(long)num_layers, 3l}),
        xt::pytensor<unsigned int, 1>(symmetry_classes, {N}),
        ctensor1(params, num_params),
        final_factor, cplx(lp_re, lp_im), g_gpu
    );
}
void ref_cnn_init_gradient(void* psi, unsigned int num_steps) {
    static_cast<PsiCNN*>(psi)->init_gradient(num_steps);
}

// PsiClassical: order in {1,2}; psi_ref == nullptr -> PsiFullyPolarized, else a PsiCNN handle.
void* ref_classical_create(unsigned int num_sites, unsigned int order, unsigned int num_ops, void** ops,
                           const double* params, unsigned int num_own_params, void* cnn_ref,
                           double lp_re, double lp_im) {
    std::vector<Operator> H_local;
    for(auto i = 0u; i < num_ops; i++) H_local.push_back(*static_cast<Operator*>(ops[i]));
    const auto p = ctensor1(params, num_own_params);
    const cplx lp(lp_re, lp_im);
    if(cnn_ref == nullptr) {
        PsiFullyPolarized fp(num_sites, lp);
        if(order == 1u) return new PsiClassicalFP<1u>(num_sites, H_local, p, fp, lp, g_gpu);
        return new PsiClassicalFP<2u>(num_sites, H_local, p, fp, lp, g_gpu);
    }
    auto& ref = *static_cast<PsiCNN*>(cnn_ref);
    if(order == 1u) return new PsiClassicalANN<1u>(num_sites, H_local, p, ref, lp, g_gpu);
    return new PsiClassicalANN<2u>(num_sites, H_local, p, ref, lp, g_gpu);
}

void ref_psi_destroy(int kind, void* h) { with_psi(kind, h, [](auto& psi) { delete &psi; }); }

unsigned int ref_psi_num_params(int kind, void* h) {
    unsigned int r = 0;
    with_psi(kind, h, [&](auto& psi) { r = psi.num_params; });
    return r;
}
void ref_psi_get_params(int kind, void* h, double* out) {
    with_psi(kind, h, [&](auto& psi) { copy_out(out, psi.get_params()); });
}
void ref_psi_set_params(int kind, void* h, const double* in) {
    with_psi(kind, h, [&](auto& psi) {
        Array<complex_t> p(psi.num_params, false);
        std::memcpy(p.host_data(), in, sizeof(complex_t) * psi.num_params);
        psi.set_params(p);
    });
}
void ref_psi_get_log_prefactor(int kind, void* h, double* out) {
    with_psi(kind, h, [&](auto& psi) { store(out, psi.log_prefactor); });
}
void ref_psi_set_log_prefactor(int kind, void* h, double re, double im) {
    with_psi(kind, h, [&](auto& psi) { psi.log_prefactor = complex_t(re, im); });
}

// ---------------------------------------------------------------- ensembles

void ref_set_gpu(int on) { g_gpu = on != 0; }
int  ref_device_synchronize() { return (int)cudaDeviceSynchronize(); }
void* ref_es_create(unsigned int num_sites) { return new ExactSummationSpins(num_sites, g_gpu); }
void* ref_mc_create(unsigned int num_samples, unsigned int num_sweeps, unsigned int num_therm, unsigned int num_chains) {
    return new MonteCarloSpins(num_samples, num_sweeps, num_therm, num_chains, Update_Policy<Spins>(), g_gpu);
}
void ref_ens_destroy(int kind, void* h) { with_ens(kind, h, [](auto& e) { delete &e; }); }
unsigned int ref_ens_num_steps(int kind, void* h) {
    unsigned int r = 0;
    with_ens(kind, h, [&](auto& e) { r = e.get_num_steps(); });
    return r;
}
void ref_mc_acceptance(void* h, unsigned int* out) {
    auto& mc = *static_cast<MonteCarloSpins*>(h);
    out[0] = mc.acceptances_ar.front();
    out[1] = mc.rejections_ar.front();
}

// ---------------------------------------------------------------- per-configuration probes

void ref_log_psi_s(int kind, void* h, uint64_t conf, double* out) {
    with_psi(kind, h, [&](auto& psi) { store(out, log_psi_s(psi, Spins(conf, 64u))); });
}
void ref_psi_O_k(int kind, void* h, uint64_t conf, double* out) {
    with_psi(kind, h, [&](auto& psi) { copy_out(out, psi_O_k(psi, Spins(conf, 64u))); });
}

// ---------------------------------------------------------------- whole-ensemble vectors

void ref_psi_vector(int kind, void* h, int ek, void* e, double* out) {
    with_psi(kind, h, [&](auto& psi) { with_ens(ek, e, [&](auto& ens) { copy_out(out, psi_vector(psi, ens)); }); });
}
void ref_log_psi_vector(int kind, void* h, int ek, void* e, double* out) {
    with_psi(kind, h, [&](auto& psi) { with_ens(ek, e, [&](auto& ens) { copy_out(out, log_psi_vector(psi, ens)); }); });
}
void ref_log_psi_mean(int kind, void* h, int ek, void* e, double* out) {
    with_psi(kind, h, [&](auto& psi) { with_ens(ek, e, [&](auto& ens) { store(out, log_psi(psi, ens)); }); });
}
double ref_psi_norm(int kind, void* h, void* es) {
    double r = 0.0;
    with_psi(kind, h, [&](auto& psi) { r = psi_norm(psi, *static_cast<ExactSummationSpins*>(es)); });
    return r;
}
void ref_psi_O_k_vector(int kind, void* h, void* es, double* out) {
    with_psi(kind, h, [&](auto& psi) { copy_out(out, psi_O_k_vector(psi, *static_cast<ExactSummationSpins*>(es))); });
}
void ref_apply_operator(int kind, void* h, void* op, int ek, void* e, double* out) {
    with_psi(kind, h, [&](auto& psi) {
        with_ens(ek, e, [&](auto& ens) { copy_out(out, apply_operator(psi, *static_cast<Operator*>(op), ens)); });
    });
}

// ---------------------------------------------------------------- ExpectationValue

void ref_expectation(int kind, void* h, void* op, int ek, void* e, double* out) {
    ExpectationValue ev(g_gpu);
    with_psi(kind, h, [&](auto& psi) {
        with_ens(ek, e, [&](auto& ens) { store(out, ev(*static_cast<Operator*>(op), psi, ens)); });
    });
}
// out = {fluctuation, Re <A>, Im <A>}
void ref_fluctuation(int kind, void* h, void* op, int ek, void* e, double* out) {
    ExpectationValue ev(g_gpu);
    with_psi(kind, h, [&](auto& psi) {
        with_ens(ek, e, [&](auto& ens) {
            const auto r = ev.fluctuation(*static_cast<Operator*>(op), psi, ens);
            out[0] = r.first; store(out + 1, r.second);
        });
    });
}
void ref_exp_sigma_z(int kind, void* h, void* op, int ek, void* e, double* out) {
    ExpectationValue ev(g_gpu);
    with_psi(kind, h, [&](auto& psi) {
        with_ens(ek, e, [&](auto& ens) { store(out, ev.exp_sigma_z(*static_cast<Operator*>(op), psi, ens)); });
    });
}
void ref_gradient(int kind, void* h, void* op, int ek, void* e, double* grad_out, double* E_out) {
    ExpectationValue ev(g_gpu);
    with_psi(kind, h, [&](auto& psi) {
        with_ens(ek, e, [&](auto& ens) {
            const auto r = ev.gradient(*static_cast<Operator*>(op), psi, ens);
            copy_out(grad_out, r.first); store(E_out, r.second);
        });
    });
}

// ---------------------------------------------------------------- HilbertSpaceDistance
// (include/network_functions/HilbertSpaceDistance.hpp:55-116; instantiated for (PsiDeep, PsiDeep) and (PsiCNN, PsiCNN))
// grad_out == null: distance only.  Returns the distance.
double ref_hilbert_space_distance(int kind, void* h, void* h_prime, void* op, int is_unitary, int ek, void* e,
                                  double* grad_out, float nu) {
    double d = 0.0;
    auto run = [&](auto& psi, auto& psi_prime) {
        HilbertSpaceDistance hsd(psi_prime.num_params, g_gpu);
        with_ens(ek, e, [&](auto& ens) {
            if(grad_out) d = hsd.gradient(reinterpret_cast<std::complex<double>*>(grad_out), psi, psi_prime, *static_cast<Operator*>(op), is_unitary != 0, ens, nu);
            else         d = hsd.distance(psi, psi_prime, *static_cast<Operator*>(op), is_unitary != 0, ens);
        });
    };
    if(kind == DEEP) run(*static_cast<PsiDeep*>(h), *static_cast<PsiDeep*>(h_prime));
    else if(kind == CNN) run(*static_cast<PsiCNN*>(h), *static_cast<PsiCNN*>(h_prime));
    return d;
}

// ---------------------------------------------------------------- KullbackLeibler
// (include/network_functions/KullbackLeibler.hpp:52-124; psi: a PsiClassical kind, psi_prime: PsiDeep or PsiCNN.)
// mode 0: value, 1: gradient, 2: gradient_with_noise.  last_md: last_mean_deviation carried in from a previous call.
// extra_out = {total_weight, Re mean_deviation, Im mean_deviation} after the call.  Returns the value.
double ref_kullback_leibler(int kind, void* h, int pkind, void* hp, int ek, void* e, int mode, double nu, double threshold,
                            double log_psi_scale, double last_md_re, double last_md_im,
                            double* grad_out, double* noise_out, double* extra_out) {
    double v = 0.0;
    auto run = [&](auto& psi, auto& psi_prime) {
        KullbackLeibler kl(psi_prime.num_params, g_gpu);
        kl.log_psi_scale = log_psi_scale;
        kl.last_mean_deviation.front() = complex_t(last_md_re, last_md_im);
        kl.last_mean_deviation.update_device();
        with_ens(ek, e, [&](auto& ens) {
            if(mode == 0) v = kl.value(psi, psi_prime, ens, threshold);
            else if(mode == 1) v = kl.gradient(reinterpret_cast<std::complex<double>*>(grad_out), psi, psi_prime, ens, nu, threshold);
            else {
                auto r = kl.gradient_with_noise(psi, psi_prime, ens, nu, threshold);
                copy_out(grad_out, std::get<0>(r));
                std::memcpy(noise_out, std::get<1>(r).host_data(), sizeof(double) * psi_prime.num_params);
                v = std::get<2>(r);
            }
        });
        extra_out[0] = kl.total_weight.front();
        store(extra_out + 1, kl.mean_deviation.front());
    };
    with_classical(kind, h, [&](auto& psi) {
        if(pkind == DEEP) run(psi, *static_cast<PsiDeep*>(hp));
        else if(pkind == CNN) run(psi, *static_cast<PsiCNN*>(hp));
    });
    return v;
}

// ---------------------------------------------------------------- TDVP

void* ref_tdvp_create(unsigned int num_params) { return new TDVP(num_params, g_gpu); }
void  ref_tdvp_destroy(void* t) { delete static_cast<TDVP*>(t); }

void ref_tdvp_eval(void* t, int kind, void* h, void* op, int ek, void* e) {
    with_psi(kind, h, [&](auto& psi) {
        with_ens(ek, e, [&](auto& ens) { static_cast<TDVP*>(t)->eval(*static_cast<Operator*>(op), psi, ens, false_t()); });
    });
}
// TDVP::eval(..., true_t) = eval_with_psi_ref (TDVP.hpp:90-93): PsiClassical kinds only.  total_weight is never cleared by
// the reference (TDVP.cu.template:206), so it is zeroed here before the call and returned.
double ref_tdvp_eval_with_psi_ref(void* t, int kind, void* h, void* op, int ek, void* e) {
    auto& tdvp = *static_cast<TDVP*>(t);
    tdvp.total_weight.clear();
    with_classical(kind, h, [&](auto& psi) {
        with_ens(ek, e, [&](auto& ens) { tdvp.eval(*static_cast<Operator*>(op), psi, ens, true_t()); });
    });
    tdvp.total_weight.update_host();
    return tdvp.total_weight.front();
}
void ref_tdvp_eval_F(void* t, int kind, void* h, void* op, int ek, void* e) {
    with_psi(kind, h, [&](auto& psi) {
        with_ens(ek, e, [&](auto& ens) { static_cast<TDVP*>(t)->eval_F_vector(*static_cast<Operator*>(op), psi, ens); });
    });
}
// scal_out = {Re E, Im E, E2, var_H}; any output pointer may be null.
void ref_tdvp_get(void* t, double* S_out, double* F_out, double* Ok_out, double* scal_out) {
    auto& tdvp = *static_cast<TDVP*>(t);
    if(S_out)  copy_out(S_out, tdvp.S_matrix);
    if(F_out)  copy_out(F_out, tdvp.F_vector);
    if(Ok_out) copy_out(Ok_out, tdvp.O_k_ar);
    if(scal_out) {
        store(scal_out, tdvp.E_local.front());
        scal_out[2] = tdvp.E2_local.front();
        scal_out[3] = tdvp.var_H();
    }
}
void ref_tdvp_get_samples(void* t, double* Ok_samples_out, double* weights_out) {
    auto& tdvp = *static_cast<TDVP*>(t);
    if(Ok_samples_out) copy_out(Ok_samples_out, *tdvp.O_k_samples);
    if(weights_out) std::memcpy(weights_out, tdvp.weight_samples->host_data(), sizeof(double) * tdvp.weight_samples->size());
}
void ref_tdvp_S_dot_vector(void* t, const double* vec, int ek, void* e, double* out) {
    auto& tdvp = *static_cast<TDVP*>(t);
    std::memcpy(tdvp.input_vector.host_data(), vec, sizeof(complex_t) * tdvp.num_params);
    with_ens(ek, e, [&](auto& ens) { tdvp.S_dot_vector(ens); });
    copy_out(out, tdvp.output_vector);
}

} // extern "C"

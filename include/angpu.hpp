// C++ face of libangpu: header-only RAII classes with the reference's names and argument order over the C ABI of
// angpu.h (which stays the drop-in boundary; nothing here adds symbols to the library).
//
//   reference (namespace ann_on_gpu)                                   here (namespace angpu_cxx, alias ann_on_gpu_b200)
//   PsiRBM(W, final_weight, log_prefactor, gpu)   PsiRBM.hpp:223-245    PsiRBM(N, M, W, final_weight, log_prefactor)
//   PsiDeep(...)                                  PsiDeep.hpp:502-560   PsiDeep(num_sites, input_weights, layers, final_weights, log_prefactor)
//   PsiCNN(...)                                   PsiCNN.hpp:306-342    PsiCNN(extent, channels, connectivity, symmetry_classes, params, final_factor, lp)
//   PsiClassical<order, ref>                      PsiClassical.hpp:192  PsiClassical(num_sites, order, H_local, params, psi_ref, log_prefactor)
//   Operator(expr, gpu)                           Operator.cpp:18-39    Operator(coefficients, a, b, words)
//   MonteCarlo_t / make MonteCarloSpins           MonteCarlo.hpp:195    MonteCarloSpins(num_samples, num_sweeps, num_therm, num_chains, seed)
//   ExactSummation_t                              ExactSummation.hpp:82 ExactSummationSpins(num_sites)
//   ExpectationValue                              ExpectationValue.hpp  ExpectationValue: operator(), fluctuation, gradient
//   TDVP                                          TDVP.hpp:16-101       TDVP: eval, eval_F_vector, S_dot_vector, var_H + solve_cg / solve (new)
//
// Ownership follows the reference: every object owns its device state (value semantics; copying deep-copies through
// angpu_psi_copy / angpu_ensemble_copy, as Array<T>'s copy constructor does, source/Array.cu:36-45); getters return
// copies.  Errors: every failing call throws std::runtime_error with angpu_last_error() (the reference: CUDA_CHECK ->
// std::runtime_error, include/types.h:139-145).  There is no gpu flag: the library is GPU-only.
#pragma once
#include <complex>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "angpu.h"

namespace angpu_cxx {

using complex_t = std::complex<double>;

inline void check(int status, const char* what) {
    if(status != 0) throw std::runtime_error(std::string(what) + ": " + angpu_last_error());
}
#define ANGPU_CXX(call) ::angpu_cxx::check((call), #call)

inline void setDevice(int device) { ANGPU_CXX(angpu_init(device)); }
inline const double* dptr(const complex_t* p) { return reinterpret_cast<const double*>(p); }
inline double* dptr(complex_t* p) { return reinterpret_cast<double*>(p); }

// multi-GPU: rank 0 creates the id, every rank initialises (angpu.h, "Multi-GPU")
struct Communicator {
    static std::vector<unsigned char> unique_id() { std::vector<unsigned char> id(ANGPU_COMM_ID_BYTES); ANGPU_CXX(angpu_comm_unique_id(id.data())); return id; }
    static void init(const std::vector<unsigned char>& id, int rank, int world) { ANGPU_CXX(angpu_comm_init(id.data(), rank, world)); }
    static void destroy() { ANGPU_CXX(angpu_comm_destroy()); }
};

class Operator {
    angpu_operator_t h_ = nullptr;
    std::vector<complex_t> coef_; std::vector<uint64_t> a_, b_; unsigned words_ = 1;
    void create() { ANGPU_CXX(angpu_operator_create((unsigned)coef_.size(), dptr(coef_.data()), a_.data(), b_.data(), words_, &h_)); }
public:
    // Pauli strings as (coefficient, a, b) with per-site masks I = (0,0), X = (1,0), Y = (0,1), Z = (1,1) (PauliString.hpp:18-21)
    Operator(std::vector<complex_t> coefficients, std::vector<uint64_t> a, std::vector<uint64_t> b, unsigned words = 1)
        : coef_(std::move(coefficients)), a_(std::move(a)), b_(std::move(b)), words_(words) { create(); }
    Operator(const Operator& o) : coef_(o.coef_), a_(o.a_), b_(o.b_), words_(o.words_) { create(); }
    Operator(Operator&& o) noexcept : h_(o.h_), coef_(std::move(o.coef_)), a_(std::move(o.a_)), b_(std::move(o.b_)), words_(o.words_) { o.h_ = nullptr; }
    Operator& operator=(Operator o) { std::swap(h_, o.h_); coef_.swap(o.coef_); a_.swap(o.a_); b_.swap(o.b_); std::swap(words_, o.words_); return *this; }
    ~Operator() { if(h_) angpu_operator_destroy(h_); }
    unsigned num_strings() const { return (unsigned)coef_.size(); }
    angpu_operator_t handle() const { return h_; }
};

class Ensemble {
protected:
    angpu_ensemble_t h_ = nullptr;
    Ensemble() = default;
public:
    Ensemble(const Ensemble& o) { ANGPU_CXX(angpu_ensemble_copy(o.h_, &h_)); }
    Ensemble(Ensemble&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Ensemble& operator=(Ensemble o) { std::swap(h_, o.h_); return *this; }
    ~Ensemble() { if(h_) angpu_ensemble_destroy(h_); }
    unsigned long long get_num_steps() const { unsigned long long n; ANGPU_CXX(angpu_ensemble_num_steps(h_, &n)); return n; }
    unsigned long long local_steps() const { unsigned long long n; ANGPU_CXX(angpu_ensemble_local_steps(h_, &n)); return n; }
    void set_shard(unsigned rank, unsigned world) { ANGPU_CXX(angpu_ensemble_set_shard(h_, rank, world)); }
    angpu_ensemble_t handle() const { return h_; }
};
struct ExactSummationSpins : Ensemble {
    explicit ExactSummationSpins(unsigned num_sites) { ANGPU_CXX(angpu_es_create(num_sites, &h_)); }
};
struct ExactSummationPaulis : Ensemble {
    explicit ExactSummationPaulis(unsigned num_sites) { ANGPU_CXX(angpu_es_paulis_create(num_sites, &h_)); }
};
struct MonteCarloPaulis : Ensemble {
    MonteCarloPaulis(unsigned long long num_samples, unsigned num_sweeps, unsigned num_thermalization_sweeps, unsigned num_markov_chains,
                     uint64_t seed = 0xA11CEull) { ANGPU_CXX(angpu_mc_paulis_create(num_samples, num_sweeps, num_thermalization_sweeps, num_markov_chains, seed, &h_)); }
};
struct MonteCarloSpins : Ensemble {
    MonteCarloSpins(unsigned long long num_samples, unsigned num_sweeps, unsigned num_thermalization_sweeps, unsigned num_markov_chains,
                    uint64_t seed = 0xA11CEull) { ANGPU_CXX(angpu_mc_create(num_samples, num_sweeps, num_thermalization_sweeps, num_markov_chains, seed, &h_)); }
    double acceptance_rate() const {
        unsigned long long ar[2]; ANGPU_CXX(angpu_mc_acceptance(h_, ar));
        return ar[0] + ar[1] ? (double)ar[0] / (double)(ar[0] + ar[1]) : 0.0;
    }
};

class Psi {
protected:
    angpu_psi_t h_ = nullptr;
    Psi() = default;
public:
    Psi(const Psi& o) { ANGPU_CXX(angpu_psi_copy(o.h_, &h_)); }
    Psi(Psi&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Psi& operator=(Psi o) { std::swap(h_, o.h_); return *this; }
    ~Psi() { if(h_) angpu_psi_destroy(h_); }
    unsigned get_num_sites() const { unsigned n; ANGPU_CXX(angpu_psi_num_sites(h_, &n)); return n; }
    unsigned num_params() const { unsigned n; ANGPU_CXX(angpu_psi_num_params(h_, &n)); return n; }
    std::vector<complex_t> get_params() const { std::vector<complex_t> p(num_params()); ANGPU_CXX(angpu_psi_get_params(h_, dptr(p.data()))); return p; }
    void set_params(const std::vector<complex_t>& p) {
        if(p.size() != num_params()) throw std::runtime_error("set_params: wrong number of parameters");
        ANGPU_CXX(angpu_psi_set_params(h_, dptr(p.data())));
    }
    complex_t log_prefactor() const { double v[2]; ANGPU_CXX(angpu_psi_get_log_prefactor(h_, v)); return {v[0], v[1]}; }
    void set_log_prefactor(complex_t v) { const double d[2] = {v.real(), v.imag()}; ANGPU_CXX(angpu_psi_set_log_prefactor(h_, d)); }
    // free functions of the reference (PsiVector / PsiNorm / PsiOkVector .cu.template) as members
    complex_t log_psi_s(const std::vector<uint64_t>& conf) const { double v[2]; ANGPU_CXX(angpu_log_psi_s(h_, conf.data(), v)); return {v[0], v[1]}; }
    std::vector<complex_t> O_k(const std::vector<uint64_t>& conf) const { std::vector<complex_t> o(num_params()); ANGPU_CXX(angpu_psi_O_k(h_, conf.data(), dptr(o.data()))); return o; }
    double norm(const ExactSummationSpins& es) const { double n; ANGPU_CXX(angpu_psi_norm(h_, es.handle(), &n)); return n; }
    std::vector<complex_t> vector(const Ensemble& es) const { std::vector<complex_t> v(es.local_steps()); ANGPU_CXX(angpu_psi_vector(h_, es.handle(), dptr(v.data()))); return v; }
    angpu_psi_t handle() const { return h_; }
};
struct PsiRBM : Psi {
    // W row-major [N][M] (PsiRBM.hpp:169-175); parameters = W only
    PsiRBM(unsigned N, unsigned M, const std::vector<complex_t>& W, complex_t final_weight, complex_t log_prefactor = 0.0) {
        if(W.size() != (size_t)N * M) throw std::runtime_error("PsiRBM: W must hold N*M entries");
        const double fw[2] = {final_weight.real(), final_weight.imag()}, lp[2] = {log_prefactor.real(), log_prefactor.imag()};
        ANGPU_CXX(angpu_rbm_create(N, M, dptr(W.data()), fw, lp, &h_));
    }
};
struct PsiDeep : Psi {
    struct Layer { unsigned size, connectivity; std::vector<complex_t> biases; std::vector<unsigned> lhs_connections; std::vector<complex_t> lhs_weights; };
    PsiDeep(unsigned num_sites, const std::vector<complex_t>& input_weights, const std::vector<Layer>& layers,
            const std::vector<complex_t>& final_weights, complex_t log_prefactor = 0.0) {
        std::vector<unsigned> sizes, conn, lhs_c; std::vector<complex_t> bias, lhs_w;
        for(const Layer& l : layers) {
            sizes.push_back(l.size); conn.push_back(l.connectivity);
            bias.insert(bias.end(), l.biases.begin(), l.biases.end());
            lhs_c.insert(lhs_c.end(), l.lhs_connections.begin(), l.lhs_connections.end());
            lhs_w.insert(lhs_w.end(), l.lhs_weights.begin(), l.lhs_weights.end());
        }
        const double lp[2] = {log_prefactor.real(), log_prefactor.imag()};
        ANGPU_CXX(angpu_deep_create(num_sites, (unsigned)input_weights.size(), dptr(input_weights.data()), (unsigned)layers.size(), sizes.data(),
                                    conn.data(), dptr(bias.data()), lhs_c.data(), dptr(lhs_w.data()), dptr(final_weights.data()), lp, &h_));
    }
};
struct PsiCNN : Psi {
    PsiCNN(const unsigned (&extent)[3], const std::vector<unsigned>& num_channels_list, const std::vector<unsigned>& connectivity_list,
           const std::vector<unsigned>& symmetry_classes, const std::vector<complex_t>& params, double final_factor, complex_t log_prefactor = 0.0) {
        const double lp[2] = {log_prefactor.real(), log_prefactor.imag()};
        ANGPU_CXX(angpu_cnn_create(extent, (unsigned)num_channels_list.size(), num_channels_list.data(), connectivity_list.data(),
                                   symmetry_classes.data(), dptr(params.data()), (unsigned)params.size(), final_factor, lp, &h_));
    }
};
struct PsiClassical : Psi {
    // psi_ref == nullptr: PsiFullyPolarized (PsiClassicalFP_<order>); else a PsiCNN (PsiClassicalANN_<order>), copied
    PsiClassical(unsigned num_sites, unsigned order, const std::vector<const Operator*>& H_local, const std::vector<complex_t>& params,
                 const PsiCNN* psi_ref = nullptr, complex_t log_prefactor = 0.0) {
        std::vector<angpu_operator_t> ops;
        for(const Operator* o : H_local) ops.push_back(o->handle());
        const double lp[2] = {log_prefactor.real(), log_prefactor.imag()};
        ANGPU_CXX(angpu_classical_create(num_sites, order, (unsigned)ops.size(), ops.data(), dptr(params.data()), (unsigned)params.size(),
                                         psi_ref ? psi_ref->handle() : nullptr, lp, &h_));
    }
};

class ExpectationValue {
    angpu_expval_t h_ = nullptr;
public:
    ExpectationValue() { ANGPU_CXX(angpu_expval_create(&h_)); }
    ExpectationValue(const ExpectationValue&) = delete;
    ExpectationValue& operator=(const ExpectationValue&) = delete;
    ~ExpectationValue() { if(h_) angpu_expval_destroy(h_); }
    complex_t operator()(const Operator& op, const Psi& psi, Ensemble& ens) { double v[2]; ANGPU_CXX(angpu_expectation(h_, op.handle(), psi.handle(), ens.handle(), v)); return {v[0], v[1]}; }
    std::pair<double, complex_t> fluctuation(const Operator& op, const Psi& psi, Ensemble& ens) {
        double f, m[2]; ANGPU_CXX(angpu_fluctuation(h_, op.handle(), psi.handle(), ens.handle(), &f, m)); return {f, {m[0], m[1]}};
    }
    std::pair<std::vector<complex_t>, complex_t> gradient(const Operator& op, const Psi& psi, Ensemble& ens) {
        std::vector<complex_t> g(psi.num_params()); double m[2];
        ANGPU_CXX(angpu_gradient(h_, op.handle(), psi.handle(), ens.handle(), dptr(g.data()), m)); return {std::move(g), {m[0], m[1]}};
    }
};

class TDVP {
    angpu_tdvp_t h_ = nullptr;
    unsigned P_;
    std::vector<complex_t> vec(int (*get)(angpu_tdvp_t, double*), size_t n) const { std::vector<complex_t> v(n); check(get(h_, dptr(v.data())), "TDVP getter"); return v; }
public:
    explicit TDVP(unsigned num_params) : P_(num_params) { ANGPU_CXX(angpu_tdvp_create(num_params, &h_)); }
    TDVP(const TDVP&) = delete;
    TDVP& operator=(const TDVP&) = delete;
    ~TDVP() { if(h_) angpu_tdvp_destroy(h_); }
    // NOTE (lifetime, as in the reference where kernels capture psi.kernel() by value): S_matrix / O_k_samples /
    // solve() after eval may read psi again -- keep psi alive until the TDVP results have been fetched.
    void eval(const Operator& op, const Psi& psi, Ensemble& ens, double s_tolerance = 0.0) {
        if(s_tolerance == 0.0) ANGPU_CXX(angpu_tdvp_eval(h_, op.handle(), psi.handle(), ens.handle()));
        else ANGPU_CXX(angpu_tdvp_eval_tol(h_, op.handle(), psi.handle(), ens.handle(), s_tolerance));
    }
    void eval_F_vector(const Operator& op, const Psi& psi, Ensemble& ens) { ANGPU_CXX(angpu_tdvp_eval_F(h_, op.handle(), psi.handle(), ens.handle())); }
    std::vector<complex_t> S_matrix() const { return vec(angpu_tdvp_get_S, (size_t)P_ * P_); }
    std::vector<complex_t> F_vector() const { return vec(angpu_tdvp_get_F, P_); }
    std::vector<complex_t> O_k_vector() const { return vec(angpu_tdvp_get_O_k, P_); }
    complex_t E_local() const { double s[5]; ANGPU_CXX(angpu_tdvp_get_scalars(h_, s)); return {s[0], s[1]}; }
    double var_H() const { double s[5]; ANGPU_CXX(angpu_tdvp_get_scalars(h_, s)); return s[2]; }
    std::vector<complex_t> S_dot_vector(const std::vector<complex_t>& v) const {
        std::vector<complex_t> out(P_); ANGPU_CXX(angpu_tdvp_S_dot_vector(h_, dptr(v.data()), dptr(out.data()))); return out;
    }
    struct CgResult { std::vector<complex_t> x; unsigned iterations; double rel_residual; };
    // PsiRBM: CG search directions / S_dot_vector on the tcgen05 tensor cores (1 on, 0 off, -1 auto = default); the residual stays fp64
    void set_tensorcore_products(int enable) { ANGPU_CXX(angpu_tdvp_set_tensorcore_products(h_, enable)); }
    CgResult solve_cg(double tol = 1e-6, unsigned max_iter = 1000, double shift_abs = 0.0, double shift_rel = 1e-3, complex_t rhs_phase = 1.0) {
        CgResult r{std::vector<complex_t>(P_), 0u, 0.0};
        const double ph[2] = {rhs_phase.real(), rhs_phase.imag()};
        ANGPU_CXX(angpu_tdvp_solve_cg(h_, tol, max_iter, shift_abs, shift_rel, ph, dptr(r.x.data()), &r.iterations, &r.rel_residual));
        return r;
    }
    std::vector<complex_t> solve(double shift_abs = 0.0, double shift_rel = 1e-3, complex_t rhs_phase = 1.0) {
        std::vector<complex_t> x(P_);
        const double ph[2] = {rhs_phase.real(), rhs_phase.imag()};
        ANGPU_CXX(angpu_tdvp_solve_dense(h_, shift_abs, shift_rel, ph, dptr(x.data())));
        return x;
    }
    // SR / TDVP parameter step on the device: params(psi) += alpha * x of the last solve
    void apply_update(Psi& psi, complex_t alpha) { const double a[2] = {alpha.real(), alpha.imag()}; ANGPU_CXX(angpu_tdvp_apply_update(h_, psi.handle(), a)); }
};

} // namespace angpu_cxx

namespace ann_on_gpu_b200 = angpu_cxx;

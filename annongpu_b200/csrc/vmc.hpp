// Ensembles and "network functions" (ExpectationValue, TDVP, probes) of the VMC hot path.
//
// Structure (differs from the reference by design):  every functional is
//     generate samples -> per-sample kernels (E_loc, O_k) -> deterministic partial sums on the device
//     -> [all-reduce hook, one packed buffer] -> finalise on the device -> copy out on request.
// The reference instead fuses a consumer lambda into one kernel per ensemble.foreach and reduces with global
// fp64 atomics from every block into single addresses (ExpectationValue.cu.template:43,259-260;
// TDVP.cu.template:107-121, 260-263).
#pragma once
#include "psi.hpp"
#include "comm.hpp"

namespace angpu {

// x = A^{-1} b, A Hermitian positive definite P x P row-major (upper triangle read, overwritten by its Cholesky factor), b overwritten
// by x; info_dev[0] = 0 or the 1-based index of the first non-positive pivot (cholesky.cu)
void cholesky_solve(cplx* A, cplx* b, unsigned P, int* info_dev, DevBuf<cplx>& work);

struct Ensemble {
    bool is_mc = false;
    bool paulis = false;            // MonteCarloPaulis / ExactSummationPaulis: configurations are Pauli strings (pauli_basis.cuh)
    // ExactSummation (include/ensembles/ExactSummation.hpp)
    unsigned num_sites = 0;
    // MonteCarlo (include/ensembles/MonteCarlo.hpp, source/ensembles/MonteCarlo.cu:14-42)
    unsigned long long num_samples = 0;
    unsigned num_sweeps = 0, num_therm = 0, num_chains = 0, call = 0;
    uint64_t seed = 0xA11CE;
    // sharding: this process owns chains / basis indices [begin, begin+count) of the global range
    // (a new ensemble inherits the communicator's rank / world: angpu_comm_init, comm.hpp)
    unsigned rank = (unsigned)comm_rank(), world = (unsigned)comm_world();
    DevBuf<unsigned long long> d_acc_rej;

    size_t num_steps() const { return is_mc ? (size_t)num_samples : ((size_t)1 << (paulis ? 2u * num_sites : num_sites)); }
    void shard(size_t total, size_t& begin, size_t& count) const {
        begin = total * rank / world;
        count = total * (rank + 1) / world - begin;
    }
    size_t local_steps() const {
        size_t b, c;
        if(is_mc) { shard(num_chains, b, c); return c * (size_t)(num_samples / num_chains); }
        shard(num_steps(), b, c); return c;
    }
    // fills S.conf / S.log_psi / S.weight for this process' share
    void generate(Psi& psi, SampleSet& S);
    // importance reweighting (ExpectationValue.cu.template:127-172, TDVP.cu.template:15-74): configurations drawn from
    // |psi_sampling|^2, then S.log_psi = log psi(s) and S.weight *= exp(2 (Re log psi(s) - Re log psi_sampling(s)))
    void generate_reweighted(Psi& psi, Psi& psi_sampling, SampleSet& S);
    DevBuf<cplx> lp_sampling;
    void acceptance(unsigned long long out[2]);
};

struct ExpectationValue {
    SampleSet S;
    DevBuf<double> d_scal;
    // <A>  (ExpectationValue.cu.template:20-50)
    cplx value(const Operator& op, Psi& psi, Ensemble& ens);
    // (sqrt(<|A_loc|^2> - |<A>|^2), <A>)  (:176-216)
    void fluctuation(const Operator& op, Psi& psi, Ensemble& ens, double& fluct, cplx& mean);
    // sum_s w_s r_s A_loc(s) / sum_s w_s r_s with samples from psi_sampling (ExpectationValue.cu.template:127-172; the
    // reference never accumulates its denominator `prob_ratio`, i.e. divides by zero -- the evident intent is implemented)
    cplx value_reweighted(const Operator& op, Psi& psi, Psi& psi_sampling, Ensemble& ens);
    // sum_s w_s exp(fast_local_energy(op, s))  (ExpectationValue.cu.template:52-82)
    cplx exp_sigma_z(const Operator& op, Psi& psi, Ensemble& ens);
};

struct TDVP {
    unsigned P;
    double threshold = -1e6;                 // kept for API parity (include/network_functions/TDVP.hpp:35)
    SampleSet S;
    // packed partial sums, all-reduced as ONE buffer: [0] = sum w E, [1] = (sum w |E|^2, sum w),
    // [2 .. 2+P) = sum w O_k, [2+P .. 2+2P) = sum w E conj(O_k)
    DevBuf<cplx> packed;
    DevBuf<cplx> F;                          // F_k = <E O_k*> - <E><O_k>*
    DevBuf<cplx> Smat;                       // P x P row-major, S = <O_k* O_k'> - <O_k>*<O_k'>
    DevBuf<cplx> O;                          // dense O_k_samples [ns][P] (valid iff have_dense_O)
    DevBuf<cplx> T;                          // PsiRBM factorised form [ns][M] (valid iff factorised)
    DevBuf<cplx> chunk_buf, row_a, vec_in, vec_out, cg_buf, vb_part, ones;
    DevBuf<double> d_scal, diag_part;
    bool have_dense_O = false, factorised = false, have_S = false, evaluated = false;
    const cplx* last_x = nullptr;            // device solution of the last solve_cg / solve_dense (valid until the next solve)
    bool sharded = false;                    // the samples of the last eval are one rank's share: later products / solves sum over ranks
    unsigned rbm_N = 0, rbm_M = 0, words = 1;
    cplx E{0.0, 0.0}; double E2 = 0.0, total_weight = 0.0;
    unsigned long long num_steps_global = 0;

    explicit TDVP(unsigned P_) : P(P_) {}
    const cplx* Ok_dev() const { return packed.p + 2; }
    // TDVP::eval (TDVP.cu.template:182-302): E, E2, <O_k>, F, O_k_samples, S.
    // psi_sampling != null: TDVP::eval(..., true_t) = eval_with_psi_ref (TDVP.cu.template:15-74): samples from psi_sampling,
    // weights multiplied by |psi/psi_sampling|^2 and NOT normalised (total_weight is exposed, as in the reference)
    void eval(const Operator& op, Psi& psi, Ensemble& ens, bool want_S, Psi* psi_sampling = nullptr);
    // TDVP::eval_F_vector (:306-334).  For PsiRBM the samples are kept in factorised form unless dense rows are requested.
    void eval_F(const Operator& op, Psi& psi, Ensemble& ens) { eval(op, psi, ens, false); }
    double var_H() const { return E2 - abs2(E); }
    // log-derivative rows of psi at the configurations in S: dense O, or the factorised (conf, T) form for a PsiRBM
    void prepare_rows(Psi& psi, bool dense);
    // x_k = sum_s w_s X_s conj(O_sk) over the local samples (deterministic two-stage sum), x_out: [P] device
    void weighted_conj_column_sums(const cplx* X, cplx* x_out);
    void ensure_dense_O(Psi* psi);
    // out = S v using the samples of the last eval (TDVP::S_dot_vector, :337-443), O(ns*P) instead of the reference's O(ns*P^2)
    void S_dot_vector_dev(const cplx* v_dev, cplx* out_dev);
    void rowdot(const cplx* v_dev);
    void matvec(const cplx* v_dev, cplx* out_dev, const cplx* dot_dev, const double* diag, double shift_abs, double shift_rel, bool allow_S = false);
    void S_dot_vector(const cplx* v_host, cplx* out_host);
    // NEW (no reference counterpart): solve (S + shift_abs*I + shift_rel*diag(S)) x = rhs_phase * F
    int  solve_cg(double tol, unsigned max_iter, double shift_abs, double shift_rel, cplx rhs_phase, cplx* x_host, double* rel_res_out);
    void solve_dense(double shift_abs, double shift_rel, cplx rhs_phase, cplx* x_host);
    void build_S();
    // opt-in fast path: 3xTF32 on tcgen05 tensor cores (sbuild_tc.cu), ~1e-5 relative to ||S||
    void build_S_tensorcore();
    DevBuf<float> tc_planes;
    // factorised S.v on the tcgen05 tensor cores (sv_tc.cu): TF32 planes of sigma (per eval), of the vector / of w a conj(T) (per product)
    DevBuf<float> tc_sig; DevBuf<cplx> tc_apart; bool tc_ready = false;
    // tc_products: S_dot_vector and the search directions of solve_cg use the tensor-core product (angpu_tdvp_set_tensorcore_products;
    // ANGPU_CG_TC=0/1 sets the default of new objects); solve_cg still refreshes the residual with the exact product
    int tc_products = tc_products_default();             // 1 on, 0 off, -1 auto (solve_cg decides by size: tc_wanted)
    static int tc_products_default();
    bool tc_wanted() const;
    bool tc_available() const;
    void tc_prepare();
    void tc_rowdot(const cplx* v_dev);
    unsigned tc_col_partials(const cplx* X, const cplx* xbar_parts, unsigned nbar, cplx** px_out);
    DevBuf<cplx> tc_zero;                                  // zeros standing in for Obar . v where the mean is already removed
    DevBuf<cplx> solve_A, solve_b, solve_work;   // dense-solve workspace (grow-only)
    DevBuf<int> solve_info;
    Psi* last_psi = nullptr;
    // optional phase timing (bench.py): CUDA events on the library stream around sample / E_loc / O_k+reduce
    bool profile = false;
    cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float phase_ms[6] = {0, 0, 0, 0, 0, 0};     // sample, eloc, ok+reduce(+allreduce), total, S build, last solve
    void mark(int i);
    ~TDVP();
};

// HilbertSpaceDistance (include/network_functions/HilbertSpaceDistance.hpp:20-116,
// source/network_functions/HilbertSpaceDistance.cu.template:16-174): distance between U|psi> (or exp(A)|psi>) and
// |psi'> estimated on samples of psi, and its gradient with respect to the parameters of psi'.
struct HilbertSpaceDistance {
    unsigned P;                      // parameters of psi_prime
    SampleSet S;                     // samples of psi (conf, log psi, weight, E_loc)
    TDVP rows;                       // holds psi_prime's rows at the same configurations (rows.S shares conf / weight)
    DevBuf<cplx> omega, ratio, g;    // per-sample omega_s, probability ratio (as complex), [2P] column sums
    DevBuf<double> d_scal;
    explicit HilbertSpaceDistance(unsigned P_) : P(P_), rows(P_) {}
    double distance(Psi& psi, Psi& psi_prime, const Operator& op, bool is_unitary, Ensemble& ens);
    double gradient(cplx* result_host, Psi& psi, Psi& psi_prime, const Operator& op, bool is_unitary, Ensemble& ens, float nu);
private:
    void averages(Psi& psi, Psi& psi_prime, const Operator& op, bool is_unitary, Ensemble& ens, bool want_gradient, double h[5]);
};

// KullbackLeibler (include/network_functions/KullbackLeibler.hpp:52-124, source/network_functions/KullbackLeibler.cu.template):
// deviation statistics of log psi' - scale * log psi on samples of psi', and their gradient with respect to psi'.
struct KullbackLeibler {
    unsigned P;                                  // parameters of psi_prime
    double log_psi_scale = 1.0;
    cplx   last_mean_deviation{0.0, 0.0}, mean_deviation{0.0, 0.0};
    double total_weight = 0.0;
    TDVP rows;                                   // rows.S = samples of psi_prime with the reweighted weights
    SampleSet Sp;                                // log psi at the same configurations
    DevBuf<cplx> dev, aux, g;                    // per-sample masked deviation; scratch factor; [k][P] column sums
    DevBuf<double> d_scal, gabs;
    explicit KullbackLeibler(unsigned P_) : P(P_), rows(P_) {}
    double value(Psi& psi, Psi& psi_prime, Ensemble& ens, double threshold);
    double gradient(cplx* result_host, Psi& psi, Psi& psi_prime, Ensemble& ens, double nu, double threshold);
    double gradient_with_noise(cplx* result_host, double* noise_host, Psi& psi, Psi& psi_prime, Ensemble& ens, double nu, double threshold);
private:
    void averages(Psi& psi, Psi& psi_prime, Ensemble& ens, double threshold, int mode, double h[6]);
};

// free functions (source/network_functions/{PsiVector,PsiNorm,PsiOkVector,ApplyOperator}.cu.template)
cplx   log_psi_s(Psi& psi, const uint64_t* conf);
void   psi_O_k(Psi& psi, const uint64_t* conf, cplx* out_host);
void   log_psi_vector(Psi& psi, Ensemble& ens, cplx* out_host, bool exponentiate);
cplx   log_psi_mean(Psi& psi, Ensemble& ens);
double psi_norm(Psi& psi, Ensemble& es);
void   psi_O_k_vector(Psi& psi, Ensemble& es, cplx* out_host);
void   apply_operator(Psi& psi, const Operator& op, Ensemble& ens, cplx* out_host);
// helpers for the C ABI
void   scalar_sums_eloc(const SampleSet& S, double* out4_dev);
void   enumerate_probe(uint64_t index, unsigned words, uint64_t* host_out);
double measure_fp64_tflops();
void   local_energies(Psi& psi, const Operator& op, const uint64_t* confs_host, size_t ns, cplx* log_psi_out, cplx* eloc_out);

} // namespace angpu

/*
 * libangpu — C ABI of the B200-native variational-Monte-Carlo hot path (drop-in for that path of
 * heikoburau/ANNonGPU).
 *
 * The reference has no C ABI: its boundary is C++ templates explicitly instantiated per
 * (Psi, Ensemble, Basis) and a pybind11 module (pyANNonGPU/main.cpp.template:65-541).  Each entry point
 * below names the reference interface it replaces (paths relative to the reference root); INTEGRATION.md
 * shows the binding a maintainer of the reference would add on top of this header.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; angpu_last_error() gives the message
 *     (the reference throws std::runtime_error from CUDA_CHECK, include/types.h:139-145);
 *   - handles are opaque and owned by the caller (angpu_*_destroy);
 *   - arrays are caller-owned HOST pointers unless the name ends in _dev; complex numbers are interleaved
 *     (re, im) doubles — the memory layout of std::complex<double> / numpy complex128;
 *   - spin configurations are `words` little-endian uint64 words, bit i of the mask <-> site i, bit set <-> s_i = +1
 *     (include/basis/Spins.h:104-117); words = ceil(num_sites / 64), num_sites <= 256 (the reference: <= 64);
 *   - Pauli strings are (a, b) masks: X = (1,0), Y = (0,1), Z = (1,1) (include/basis/PauliString.hpp:37-56);
 *   - there is no `gpu` flag and no CPU fallback: everything runs on the CUDA device chosen by angpu_init.
 *   - not re-entrant per handle (as the reference: objects hold mutable accumulators); one stream per process.
 */
#ifndef ANGPU_H
#define ANGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct angpu_psi_s*      angpu_psi_t;        /* PsiRBM | PsiDeep | PsiCNN | PsiClassical */
typedef struct angpu_operator_s* angpu_operator_t;   /* Operator = StandartOperator<PauliString> */
typedef struct angpu_ensemble_s* angpu_ensemble_t;   /* MonteCarloSpins | ExactSummationSpins */
typedef struct angpu_expval_s*   angpu_expval_t;     /* ExpectationValue */
typedef struct angpu_tdvp_s*     angpu_tdvp_t;       /* TDVP */
typedef struct angpu_hsd_s*      angpu_hsd_t;        /* HilbertSpaceDistance */
typedef struct angpu_kl_s*       angpu_kl_t;         /* KullbackLeibler */

enum { ANGPU_PSI_RBM = 0, ANGPU_PSI_DEEP = 1, ANGPU_PSI_CNN = 2, ANGPU_PSI_CLASSICAL = 3 };

/* ---- runtime ------------------------------------------------------------------------------------------ */
/* setDevice (source/ANNonGPU.cu:7-9): selects the device and creates the library's stream. */
int angpu_init(int device);
/* Run on a caller-owned CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); NULL restores the own stream. */
int angpu_set_stream(void* cuda_stream);
int angpu_synchronize(void);
const char* angpu_last_error(void);
/* start_profiling / stop_profiling (source/ANNonGPU.cu:11-17) */
int angpu_profiler_start(void);
int angpu_profiler_stop(void);
/* number of kernels launched by the library since the last call with reset != 0 */
unsigned long long angpu_launch_count(int reset);
/* Multi-GPU (no reference counterpart; the reference is single-device).  One process per GPU.  The library owns an NCCL
 * communicator (libnccl.so.2 is opened at run time; single-GPU users never load it) and sums the partial results of
 * SHARDED ensembles (angpu_ensemble_set_shard / inherited rank, world) over ranks on its own stream:
 *   rank 0:      angpu_comm_unique_id(id)   -> ship the 128 bytes to every rank by any means (file, MPI, torch.distributed)
 *   every rank:  angpu_init(device); angpu_comm_init(id, rank, world)
 * Ensembles created afterwards own the chains / basis indices of `rank` out of `world`.  angpu_comm_init(.., 0, 1) or
 * angpu_comm_destroy() return to single-process operation. */
#define ANGPU_COMM_ID_BYTES 128
int angpu_comm_unique_id(unsigned char id_out[ANGPU_COMM_ID_BYTES]);
int angpu_comm_init(const unsigned char id[ANGPU_COMM_ID_BYTES], int rank, int world);
int angpu_comm_destroy(void);
int angpu_comm_rank(int* rank_out, int* world_out);
/* Alternative transport for hosts that own their communicator: a callback that sums `count` doubles at `dev_ptr` in
 * place over all ranks, ordered on the library's stream, and returns 0 on success (non-zero makes the calling entry
 * point fail instead of continuing with un-reduced partial sums).  Used only while no NCCL communicator is installed. */
typedef int (*angpu_allreduce_fn)(void* dev_ptr, unsigned long long count, void* user);
int angpu_set_allreduce(angpu_allreduce_fn fn, void* user);

/* ---- basis / operator primitives ----------------------------------------------------------------------- */
/* Spins::enumerate (include/basis/Spins.h:291-296) */
int angpu_spins_enumerate(uint64_t index, unsigned words, uint64_t* conf_out);
/* PauliString::apply(Spins) (include/basis/PauliString.hpp:242-255), evaluated on the device */
int angpu_pauli_apply(const uint64_t* a, const uint64_t* b, const uint64_t* conf, unsigned words,
                      double coeff_out[2], uint64_t* conf_out);
/* my_logcosh / my_tanh (include/quantum_state/psi_functions.hpp:11-49, 80-116; bound as activation_function,
 * pyANNonGPU/main.cpp.template:534-536), evaluated on the device */
int angpu_activation(const double z[2], unsigned layer, double logcosh_out[2], double tanh_out[2]);

/* Operator(expr, gpu) (source/operator/Operator.cpp:18-39) from raw (coefficient, a, b) arrays; a, b: n x words */
int angpu_operator_create(unsigned n, const double* coeffs, const uint64_t* a, const uint64_t* b, unsigned words,
                          angpu_operator_t* out);
int angpu_operator_destroy(angpu_operator_t op);
int angpu_operator_num_strings(angpu_operator_t op, unsigned* out);

/* ---- wavefunctions ------------------------------------------------------------------------------------- */
/* PsiRBM(W[N,M], final_weight, log_prefactor, gpu) (include/quantum_state/PsiRBM.hpp:223-245) */
int angpu_rbm_create(unsigned N, unsigned M, const double* W, const double final_weight[2], const double log_prefactor[2],
                     angpu_psi_t* out);
/* PsiDeep(num_sites, input_weights, biases[], lhs_connections[], lhs_weights[], final_weights, log_prefactor, gpu)
 * (include/quantum_state/PsiDeep.hpp:502-560).  Per hidden layer l: sizes[l] units with conn[l] inputs each;
 * biases / lhs_connections / lhs_weights are the concatenations over layers of [size], [conn x size], [conn x size]. */
int angpu_deep_create(unsigned num_sites, unsigned N, const double* input_weights, unsigned num_hidden,
                      const unsigned* sizes, const unsigned* conn, const double* biases,
                      const unsigned* lhs_connections, const double* lhs_weights, const double* final_weights,
                      const double log_prefactor[2], angpu_psi_t* out);
/* PsiCNN(extent[3], num_channels_list[L], connectivity_list[L,3], symmetry_classes[N], params[P], final_factor,
 * log_prefactor, gpu) (include/quantum_state/PsiCNN.hpp:306-342) */
int angpu_cnn_create(const unsigned extent[3], unsigned num_layers, const unsigned* num_channels,
                     const unsigned* connectivity, const unsigned* symmetry_classes, const double* params,
                     unsigned num_params, double final_factor, const double log_prefactor[2], angpu_psi_t* out);
/* PsiClassicalFP_<order> / PsiClassicalANN_<order>(num_sites, H_local, params, psi_ref, log_prefactor, gpu)
 * (include/quantum_state/PsiClassical.hpp:192-214).  psi_ref == NULL: PsiFullyPolarized; else a PsiCNN handle (copied). */
int angpu_classical_create(unsigned num_sites, unsigned order, unsigned num_ops, const angpu_operator_t* H_local,
                           const double* params, unsigned num_own_params, angpu_psi_t psi_ref,
                           const double log_prefactor[2], angpu_psi_t* out);
int angpu_psi_copy(angpu_psi_t psi, angpu_psi_t* out);                    /* Psi::copy() */
int angpu_psi_destroy(angpu_psi_t psi);
int angpu_psi_kind(angpu_psi_t psi, int* out);
int angpu_psi_num_sites(angpu_psi_t psi, unsigned* out);
int angpu_psi_num_params(angpu_psi_t psi, unsigned* out);
int angpu_psi_get_params(angpu_psi_t psi, double* out);                   /* get_params() */
int angpu_psi_set_params(angpu_psi_t psi, const double* in);              /* set_params() */
int angpu_psi_get_log_prefactor(angpu_psi_t psi, double out[2]);
int angpu_psi_set_log_prefactor(angpu_psi_t psi, const double in[2]);

/* ---- ensembles ----------------------------------------------------------------------------------------- */
/* ExactSummationSpins(num_sites, gpu) (include/ensembles/ExactSummation.hpp:82-114) */
int angpu_es_create(unsigned num_sites, angpu_ensemble_t* out);
/* MonteCarloSpins(num_samples, num_sweeps, num_thermalization_sweeps, num_markov_chains, gpu)
 * (include/ensembles/MonteCarlo.hpp:195-259) + seed (additive: counter-based Philox replaces the fixed-seed XORWOW
 * states of source/RNGStates.cu:13-19) */
int angpu_mc_create(unsigned long long num_samples, unsigned num_sweeps, unsigned num_thermalization_sweeps,
                    unsigned num_markov_chains, uint64_t seed, angpu_ensemble_t* out);
/* ExactSummationPaulis(num_sites, gpu) / MonteCarloPaulis(...) (include/ensembles/ExactSummation.hpp:124-126,
 * include/ensembles/MonteCarlo.hpp:265-283; bindings pyANNonGPU/main.cpp.template:369-377, 400-405): the Pauli-string
 * (density-matrix) basis.  Configurations are Pauli strings (4^num_sites of them); the wavefunction is a PsiDeep with
 * N = 3 num_sites input units (include/quantum_state/PsiDeep.hpp:282-308); operators act by Pauli multiplication
 * (include/basis/PauliString.hpp:257-277).  Configurations cross the boundary as "units" masks: bit 3 s + t set iff site s
 * carries Pauli type t + 1 (X, Y, Z) -- PauliString::network_unit_at (include/basis/PauliString.hpp:84-90). */
int angpu_es_paulis_create(unsigned num_sites, angpu_ensemble_t* out);
int angpu_mc_paulis_create(unsigned long long num_samples, unsigned num_sweeps, unsigned num_thermalization_sweeps,
                           unsigned num_markov_chains, uint64_t seed, angpu_ensemble_t* out);
int angpu_ensemble_copy(angpu_ensemble_t ens, angpu_ensemble_t* out);
int angpu_ensemble_destroy(angpu_ensemble_t ens);
int angpu_ensemble_num_steps(angpu_ensemble_t ens, unsigned long long* out);          /* get_num_steps() (global) */
int angpu_ensemble_local_steps(angpu_ensemble_t ens, unsigned long long* out);        /* this process' share */
/* Multi-GPU: this process owns chains / basis indices [rank*n/world, (rank+1)*n/world). */
int angpu_ensemble_set_shard(angpu_ensemble_t ens, unsigned rank, unsigned world);
/* acceptances_ar / rejections_ar of the last call (include/ensembles/MonteCarlo.hpp:164-169), this process' chains */
int angpu_mc_acceptance(angpu_ensemble_t ens, unsigned long long out[2]);
/* Position of the Monte-Carlo random streams: the number of sampling calls made so far (the Philox counter block; the
 * reference's persistent per-chain RNG state).  Setting it replays / aligns streams between ensembles. */
int angpu_mc_get_call_index(angpu_ensemble_t ens, unsigned* out);
int angpu_mc_set_call_index(angpu_ensemble_t ens, unsigned call_index);
/* sampler counters of the last call: out[0..1] as angpu_mc_acceptance, out[2] = proposals whose accept/reject decision
 * needed the fp64 evaluation (the fp32-screened PsiRBM sampler, csrc/rbm_sampler.cuh; 0 for the all-fp64 samplers),
 * out[3] reserved (0).  No reference counterpart. */
int angpu_mc_counters(angpu_ensemble_t ens, unsigned long long out[4]);
/* Runs the sampler once and copies this process' configurations [local_steps][words] and log psi out (testing aid). */
int angpu_ensemble_sample(angpu_ensemble_t ens, angpu_psi_t psi, uint64_t* confs_out, double* log_psi_out);

/* ---- probes and whole-ensemble vectors (source/network_functions/{PsiVector,PsiNorm,PsiOkVector,ApplyOperator}.cu.template) */
int angpu_log_psi_s(angpu_psi_t psi, const uint64_t* conf, double out[2]);
int angpu_psi_O_k(angpu_psi_t psi, const uint64_t* conf, double* out);
int angpu_log_psi_vector(angpu_psi_t psi, angpu_ensemble_t ens, double* out);          /* [local_steps] */
int angpu_psi_vector(angpu_psi_t psi, angpu_ensemble_t ens, double* out);
int angpu_log_psi_mean(angpu_psi_t psi, angpu_ensemble_t ens, double out[2]);          /* log_psi(psi, ens) */
int angpu_psi_norm(angpu_psi_t psi, angpu_ensemble_t es, double* out);
int angpu_psi_O_k_vector(angpu_psi_t psi, angpu_ensemble_t es, double* out);
int angpu_apply_operator(angpu_psi_t psi, angpu_operator_t op, angpu_ensemble_t ens, double* out);
/* E_loc and log psi on caller-given configurations [ns][words] (StandartOperator::local_energy, Operator.hpp:88-121) */
int angpu_local_energies(angpu_psi_t psi, angpu_operator_t op, const uint64_t* confs, unsigned long long ns,
                         double* log_psi_out, double* eloc_out);

/* ---- ExpectationValue (source/network_functions/ExpectationValue.cu.template) --------------------------- */
int angpu_expval_create(angpu_expval_t* out);
int angpu_expval_destroy(angpu_expval_t ev);
int angpu_expectation(angpu_expval_t ev, angpu_operator_t op, angpu_psi_t psi, angpu_ensemble_t ens, double out[2]);     /* :20-50 */
int angpu_expectation_many(angpu_expval_t ev, unsigned num_ops, const angpu_operator_t* ops, angpu_psi_t psi,
                           angpu_ensemble_t ens, double* out);                                                            /* :81-130 */
/* importance-reweighted <A>: configurations from |psi_sampling|^2, weights times |psi/psi_sampling|^2, result normalised
 * by the summed weights (operator()(op, psi, psi_sampling, ens), :127-172; the reference divides by a never-accumulated
 * `prob_ratio`, i.e. by zero -- the intended quotient is returned here) */
int angpu_expectation_reweighted(angpu_expval_t ev, angpu_operator_t op, angpu_psi_t psi, angpu_psi_t psi_sampling,
                                 angpu_ensemble_t ens, double out[2]);
/* sum_s w_s exp(sum_n c_n <s|P_n|s'>_coefficient)  (exp_sigma_z, :52-82; Operator.hpp:138-157) */
int angpu_exp_sigma_z(angpu_expval_t ev, angpu_operator_t op, angpu_psi_t psi, angpu_ensemble_t ens, double out[2]);
int angpu_fluctuation(angpu_expval_t ev, angpu_operator_t op, angpu_psi_t psi, angpu_ensemble_t ens,
                      double* fluctuation_out, double mean_out[2]);                                                       /* :176-216 */
int angpu_gradient(angpu_expval_t ev, angpu_operator_t op, angpu_psi_t psi, angpu_ensemble_t ens,
                   double* gradient_out, double mean_out[2]);                                                             /* :220-275 */

/* ---- TDVP (include/network_functions/TDVP.hpp, source/network_functions/TDVP.cu.template) --------------- */
int angpu_tdvp_create(unsigned num_params, angpu_tdvp_t* out);
int angpu_tdvp_destroy(angpu_tdvp_t tdvp);
int angpu_tdvp_eval(angpu_tdvp_t tdvp, angpu_operator_t op, angpu_psi_t psi, angpu_ensemble_t ens);    /* eval, :182-302 */
int angpu_tdvp_eval_F(angpu_tdvp_t tdvp, angpu_operator_t op, angpu_psi_t psi, angpu_ensemble_t ens);  /* eval_F_vector, :306-334 */
/* eval with an accuracy budget for S: s_tolerance = 0 builds S in exact fp64 (= angpu_tdvp_eval); s_tolerance >= 1e-5 (relative
 * to ||S||) selects the tcgen05 tensor-core build (3xTF32 split, fp32 accumulation in TMEM; measured ~2e-6), 8x faster at C4.
 * E, F, <O_k> and O_k_samples are fp64 either way.  Values in (0, 1e-5) are an error.  Additive, no reference counterpart. */
int angpu_tdvp_eval_tol(angpu_tdvp_t tdvp, angpu_operator_t op, angpu_psi_t psi, angpu_ensemble_t ens, double s_tolerance);
/* eval_with_psi_ref = TDVP::eval(..., true_t) (TDVP.hpp:90-93, TDVP.cu.template:15-74): samples from psi_sampling (the
 * reference passes psi.psi_ref of a PsiClassical), weights w_s |psi(s)/psi_sampling(s)|^2, sums NOT normalised;
 * total_weight (angpu_tdvp_get_scalars) = sum of those weights for THIS call (the reference never clears it).
 * psi_sampling == NULL: psi must be a PsiClassical and its own reference state is used. */
int angpu_tdvp_eval_reweighted(angpu_tdvp_t tdvp, angpu_operator_t op, angpu_psi_t psi, angpu_psi_t psi_sampling, angpu_ensemble_t ens);
int angpu_tdvp_get_S(angpu_tdvp_t tdvp, double* out);                 /* S_matrix [P][P] row-major */
int angpu_tdvp_get_F(angpu_tdvp_t tdvp, double* out);                 /* F_vector */
int angpu_tdvp_get_O_k(angpu_tdvp_t tdvp, double* out);               /* O_k_ar */
int angpu_tdvp_get_scalars(angpu_tdvp_t tdvp, double out[5]);         /* Re E, Im E, E2, var_H (TDVP.hpp:69-71), total_weight */
int angpu_tdvp_num_local_samples(angpu_tdvp_t tdvp, unsigned long long* out);
int angpu_tdvp_get_O_k_samples(angpu_tdvp_t tdvp, double* out);       /* O_k_samples [local_steps][P] */
int angpu_tdvp_get_weights(angpu_tdvp_t tdvp, double* out);           /* weight_samples */
int angpu_tdvp_get_E_local_samples(angpu_tdvp_t tdvp, double* out);   /* the reference's never-allocated E_local_samples, TDVP.hpp:30 */
int angpu_tdvp_S_dot_vector(angpu_tdvp_t tdvp, const double* vec, double* out);                        /* :337-443 */
/* PsiRBM only (no reference counterpart): run the factorised S.v -- angpu_tdvp_S_dot_vector and the search-direction products of
 * angpu_tdvp_solve_cg -- on the tcgen05 tensor cores (sigma exact in TF32, the vector / w a conj(T) as TF32 hi + lo planes, fp32
 * accumulation in TMEM: ~1e-6 relative).  angpu_tdvp_solve_cg then recomputes the true residual with the exact FP64-tensor-core
 * product every 32 iterations and before it accepts a residual: the reported residual is the fp64 one.
 * enable: 1 on, 0 off (every product exact), -1 auto = the default: angpu_tdvp_S_dot_vector exact, angpu_tdvp_solve_cg on the tensor
 * cores when ns N M >= 1e9 (where the pipeline wins: BASELINE configuration 5).  ANGPU_CG_TC=0/1 sets the default of new objects. */
int angpu_tdvp_set_tensorcore_products(angpu_tdvp_t tdvp, int enable);

/* NEW (the reference contains no solver): x solves
 * (S + shift_abs*I + shift_rel*diag(S)) x = rhs_phase * F.  CG is matrix-free on the samples of the last eval / eval_F. */
int angpu_tdvp_solve_cg(angpu_tdvp_t tdvp, double tol, unsigned max_iter, double shift_abs, double shift_rel,
                        const double rhs_phase[2], double* x_out, unsigned* iterations_out, double* rel_residual_out);
int angpu_tdvp_solve_dense(angpu_tdvp_t tdvp, double shift_abs, double shift_rel, const double rhs_phase[2], double* x_out);
/* x_out may be NULL in both solvers: the solution then stays on the device, and
 * angpu_tdvp_apply_update performs the SR / TDVP step  params(psi) += alpha * x  there (no host round trip of P numbers). */
int angpu_tdvp_apply_update(angpu_tdvp_t tdvp, angpu_psi_t psi, const double alpha[2]);
/* NEW, opt-in: rebuild S from the samples of the last eval on the tcgen05 tensor cores (3xTF32 split, fp32 accumulation in
 * TMEM; ~1e-5 relative to ||S||).  angpu_tdvp_eval itself builds S in exact fp64 (the reference: fp64 atomics, :216-279). */
int angpu_tdvp_build_S_tensorcore(angpu_tdvp_t tdvp);

/* The solver behind angpu_tdvp_solve_dense on caller-given data: x = A^{-1} b for a Hermitian positive definite A
 * (n x n complex, row-major; only the upper triangle is read), blocked Cholesky + triangular solves on the device
 * (csrc/cholesky.cu).  Fails when a pivot is not positive. */
int angpu_hpd_solve(unsigned n, const double* A, const double* b, double* x_out);

/* ---- measurement aids (no reference counterpart) -------------------------------------------------------- */
/* CUDA-event timing on the library stream, ms: {sampling, E_loc, O_k + reductions, eval total} of the last eval / eval_F,
 * {S build, last solve_cg / solve_dense} */
int angpu_tdvp_set_profile(angpu_tdvp_t tdvp, int enable);
int angpu_tdvp_phase_ms(angpu_tdvp_t tdvp, double out[6]);
/* measured FP64 FMA throughput of the device (TFLOP/s), the roofline denominator of the FP64-pipe-bound kernels */
int angpu_measure_fp64_tflops(double* out);

/* ---- HilbertSpaceDistance(num_params, gpu)  (include/network_functions/HilbertSpaceDistance.hpp:55-116,
 * source/network_functions/HilbertSpaceDistance.cu.template:16-174; pyANNonGPU/main.cpp.template:436-440).
 * Samples s ~ |psi|^2; with A_loc = local energy of `op` on psi:
 *   is_unitary: omega_s = exp(conj(log psi'(s) - log psi(s))) A_loc(s),  next-state norm = <|A_loc|^2>
 *   else:       omega_s = exp(A_loc(s) + conj(log psi'(s) - log psi(s))), next-state norm = <exp(2 Re A_loc)>
 *   distance = sqrt(max(1 - |<omega>|^2 / (norm <|psi'/psi|^2>), 1e-8));
 * gradient: d distance / d conj(theta'_k) of psi_prime's parameters divided by distance^nu; returns the distance too. */
int angpu_hsd_create(unsigned num_params, angpu_hsd_t* out);
int angpu_hsd_destroy(angpu_hsd_t hsd);
int angpu_hsd_distance(angpu_hsd_t hsd, angpu_psi_t psi, angpu_psi_t psi_prime, angpu_operator_t op, int is_unitary,
                       angpu_ensemble_t ens, double* distance_out);
int angpu_hsd_gradient(angpu_hsd_t hsd, angpu_psi_t psi, angpu_psi_t psi_prime, angpu_operator_t op, int is_unitary,
                       angpu_ensemble_t ens, float nu, double* gradient_out, double* distance_out);

/* ---- KullbackLeibler(num_params, gpu)  (include/network_functions/KullbackLeibler.hpp:52-124,
 * source/network_functions/KullbackLeibler.cu.template:16-321; pyANNonGPU/main.cpp.template:445-461).
 * Samples s ~ |psi'|^2 (psi_prime); weight_s = w'_s |psi(s)^scale / psi'(s)|^2; deviation_s = log psi'(s) - scale log psi(s)
 * - last_mean_deviation (the weighted mean of the previous call, kept in the object); samples with |deviation| <= threshold
 * are left out of the deviation sums.  value = sqrt(max(1e-8, <|dev|^2> - |<dev>|^2)); gradient_k = (<dev conj(O'_k)> -
 * <dev> conj(<O'_k>)) / value^nu with respect to psi_prime's parameters; noise_k = its sampling error estimate
 * (dense rows of psi_prime).  The reference binds psi = PsiClassical*, psi_prime = PsiDeep | PsiCNN; any pair works here. */
int angpu_kl_create(unsigned num_params, angpu_kl_t* out);
int angpu_kl_destroy(angpu_kl_t kl);
int angpu_kl_set_log_psi_scale(angpu_kl_t kl, double scale);
int angpu_kl_get_state(angpu_kl_t kl, double out[4]);        /* total_weight, Re/Im mean_deviation, log_psi_scale */
int angpu_kl_value(angpu_kl_t kl, angpu_psi_t psi, angpu_psi_t psi_prime, angpu_ensemble_t ens, double threshold, double* value_out);
int angpu_kl_gradient(angpu_kl_t kl, angpu_psi_t psi, angpu_psi_t psi_prime, angpu_ensemble_t ens, double nu, double threshold,
                      double* gradient_out, double* value_out);
int angpu_kl_gradient_with_noise(angpu_kl_t kl, angpu_psi_t psi, angpu_psi_t psi_prime, angpu_ensemble_t ens, double nu,
                                 double threshold, double* gradient_out, double* noise_out, double* value_out);

#ifdef __cplusplus
}
#endif
#endif /* ANGPU_H */

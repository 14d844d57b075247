// extern "C" boundary of libangpu (declared in include/angpu.h).  Thin: argument checks, handle bookkeeping,
// exception -> error-code translation.  All work happens in psi.cu / vmc.cu.
#include "../../include/angpu.h"
#include "vmc.hpp"
#include <cuda_profiler_api.h>
#include <memory>
#include <string>

using namespace angpu;

struct angpu_psi_s      { std::unique_ptr<Psi> p; };
struct angpu_operator_s { std::unique_ptr<Operator> p; };
struct angpu_ensemble_s { Ensemble e; };
struct angpu_expval_s   { ExpectationValue ev; std::unique_ptr<TDVP> grad; };
struct angpu_tdvp_s     { std::unique_ptr<TDVP> t; };
struct angpu_hsd_s      { std::unique_ptr<HilbertSpaceDistance> h; };
struct angpu_kl_s       { std::unique_ptr<KullbackLeibler> k; };

static thread_local std::string g_err;

#define API_BEGIN try { if(ctx().device < 0) ctx_init(0);
#define API_END   return 0; } catch(const std::exception& e) { g_err = e.what(); return 1; } catch(...) { g_err = "unknown error"; return 1; }
#define NOTNULL(x) ANGPU_REQUIRE((x) != nullptr, "null argument: " #x)

static inline cplx c2(const double* p) { return cplx(p[0], p[1]); }
static inline const cplx* cp(const double* p) { return reinterpret_cast<const cplx*>(p); }
static inline cplx* cp(double* p) { return reinterpret_cast<cplx*>(p); }

// ---- tiny device probes so that the primitives under test are the DEVICE implementations
__global__ void k_probe_pauli(OpDev op, const uint64_t* conf, cplx* coeff_out, uint64_t* conf_out) {
    // one string, original (a,b) semantics: coefficient = prefactor (folded into op.coef) * sign; s' = s ^ flip
    if(threadIdx.x == 0) {
        const bool diag = op.num_diag == 1u;
        *coeff_out = string_sign(op, 0u, conf) * op.coef[0];
        for(unsigned w = 0; w < op.words; w++) conf_out[w] = conf[w] ^ (diag ? 0ull : op.flip[w]);
    }
}
__global__ void k_probe_activation(cplx z, unsigned layer, cplx* out) {
    if(threadIdx.x == 0) { out[0] = act_lc(z, layer); out[1] = act_th(z, layer); }
}

extern "C" {
#pragma GCC visibility push(default)

const char* angpu_last_error(void) { return g_err.c_str(); }

int angpu_init(int device) { try { ctx_init(device); return 0; } catch(const std::exception& e) { g_err = e.what(); return 1; } }
int angpu_set_stream(void* cuda_stream) {
    API_BEGIN
    Ctx& c = ctx();
    ANGPU_CUDA(cudaStreamSynchronize(c.stream));
    if(cuda_stream == nullptr) {
        if(!c.own_stream) { ANGPU_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking)); c.own_stream = true; }
    } else {
        if(c.own_stream && c.stream) cudaStreamDestroy(c.stream);
        c.stream = static_cast<cudaStream_t>(cuda_stream); c.own_stream = false;
    }
    API_END
}
int angpu_synchronize(void) { API_BEGIN ANGPU_CUDA(cudaStreamSynchronize(stream())); API_END }
int angpu_profiler_start(void) { API_BEGIN ANGPU_CUDA(cudaProfilerStart()); API_END }
int angpu_profiler_stop(void) { API_BEGIN ANGPU_CUDA(cudaProfilerStop()); API_END }
unsigned long long angpu_launch_count(int reset) { const unsigned long long n = ctx().launches; if(reset) ctx().launches = 0; return n; }
int angpu_set_allreduce(angpu_allreduce_fn fn, void* user) { set_allreduce(fn, user); return 0; }
int angpu_comm_unique_id(unsigned char id_out[ANGPU_COMM_ID_BYTES]) { API_BEGIN NOTNULL(id_out); comm_unique_id(id_out); API_END }
int angpu_comm_init(const unsigned char id[ANGPU_COMM_ID_BYTES], int rank, int world) { API_BEGIN NOTNULL(id); comm_init(id, rank, world); API_END }
int angpu_comm_destroy(void) { API_BEGIN comm_destroy(); API_END }
int angpu_comm_rank(int* rank_out, int* world_out) { API_BEGIN NOTNULL(rank_out); NOTNULL(world_out); *rank_out = comm_rank(); *world_out = comm_world(); API_END }

// ---- primitives
int angpu_spins_enumerate(uint64_t index, unsigned words, uint64_t* conf_out) {
    API_BEGIN NOTNULL(conf_out);
    ANGPU_REQUIRE(words >= 1 && words <= (unsigned)MAXW, "words must be in 1..4");
    enumerate_probe(index, words, conf_out);
    API_END
}
int angpu_pauli_apply(const uint64_t* a, const uint64_t* b, const uint64_t* conf, unsigned words, double coeff_out[2], uint64_t* conf_out) {
    API_BEGIN NOTNULL(a); NOTNULL(b); NOTNULL(conf); NOTNULL(coeff_out); NOTNULL(conf_out);
    const double one[2] = {1.0, 0.0};
    Operator op(1, one, a, b, words);
    DevBuf<uint64_t> dc; dc.upload(conf, words);
    DevBuf<uint64_t> dout(words); DevBuf<cplx> dco(1);
    k_probe_pauli<<<1, 32, 0, stream()>>>(op.dev, dc.p, dco.p, dout.p);
    ANGPU_CHECK_LAUNCH(); count_launch();
    dco.download(cp(coeff_out), 1); dout.download(conf_out, words);
    API_END
}
int angpu_activation(const double z[2], unsigned layer, double logcosh_out[2], double tanh_out[2]) {
    API_BEGIN NOTNULL(z); NOTNULL(logcosh_out); NOTNULL(tanh_out);
    DevBuf<cplx> d(2);
    k_probe_activation<<<1, 32, 0, stream()>>>(c2(z), layer, d.p);
    ANGPU_CHECK_LAUNCH(); count_launch();
    cplx h[2]; d.download(h, 2);
    logcosh_out[0] = h[0].re; logcosh_out[1] = h[0].im; tanh_out[0] = h[1].re; tanh_out[1] = h[1].im;
    API_END
}

int angpu_operator_create(unsigned n, const double* coeffs, const uint64_t* a, const uint64_t* b, unsigned words, angpu_operator_t* out) {
    API_BEGIN NOTNULL(out);
    if(n) { NOTNULL(coeffs); NOTNULL(a); NOTNULL(b); }
    *out = new angpu_operator_s{std::unique_ptr<Operator>(new Operator(n, coeffs, a, b, words))};
    API_END
}
int angpu_operator_destroy(angpu_operator_t op) { API_BEGIN delete op; API_END }
int angpu_operator_num_strings(angpu_operator_t op, unsigned* out) { API_BEGIN NOTNULL(op); NOTNULL(out); *out = op->p->num_strings; API_END }

// ---- wavefunctions
int angpu_rbm_create(unsigned N, unsigned M, const double* W, const double fw[2], const double lp[2], angpu_psi_t* out) {
    API_BEGIN NOTNULL(W); NOTNULL(fw); NOTNULL(lp); NOTNULL(out);
    *out = new angpu_psi_s{std::unique_ptr<Psi>(new PsiRBM(N, M, cp(W), c2(fw), c2(lp)))};
    API_END
}
int angpu_deep_create(unsigned num_sites, unsigned N, const double* input_weights, unsigned num_hidden, const unsigned* sizes,
                      const unsigned* conn, const double* biases, const unsigned* lhs_connections, const double* lhs_weights,
                      const double* final_weights, const double lp[2], angpu_psi_t* out) {
    API_BEGIN NOTNULL(input_weights); NOTNULL(sizes); NOTNULL(conn); NOTNULL(biases); NOTNULL(lhs_connections);
    NOTNULL(lhs_weights); NOTNULL(final_weights); NOTNULL(lp); NOTNULL(out);
    *out = new angpu_psi_s{std::unique_ptr<Psi>(new PsiDeep(num_sites, N, cp(input_weights), num_hidden, sizes, conn, cp(biases),
                                                             lhs_connections, cp(lhs_weights), cp(final_weights), c2(lp)))};
    API_END
}
int angpu_cnn_create(const unsigned extent[3], unsigned num_layers, const unsigned* num_channels, const unsigned* connectivity,
                     const unsigned* symmetry_classes, const double* params, unsigned num_params, double final_factor,
                     const double lp[2], angpu_psi_t* out) {
    API_BEGIN NOTNULL(extent); NOTNULL(num_channels); NOTNULL(connectivity); NOTNULL(symmetry_classes); NOTNULL(params); NOTNULL(lp); NOTNULL(out);
    *out = new angpu_psi_s{std::unique_ptr<Psi>(new PsiCNN(extent, num_layers, num_channels, connectivity, symmetry_classes,
                                                            cp(params), num_params, final_factor, c2(lp)))};
    API_END
}
int angpu_classical_create(unsigned num_sites, unsigned order, unsigned num_ops, const angpu_operator_t* H_local, const double* params,
                           unsigned num_own_params, angpu_psi_t psi_ref, const double lp[2], angpu_psi_t* out) {
    API_BEGIN NOTNULL(lp); NOTNULL(out);
    if(num_ops) { NOTNULL(H_local); NOTNULL(params); }
    std::vector<const Operator*> ops;
    for(unsigned i = 0; i < num_ops; i++) { NOTNULL(H_local[i]); ops.push_back(H_local[i]->p.get()); }
    const PsiCNN* ref = nullptr;
    if(psi_ref) {
        ANGPU_REQUIRE(psi_ref->p->kind == Psi::CNN, "PsiClassical: psi_ref must be a PsiCNN (or NULL for PsiFullyPolarized)");
        ref = static_cast<const PsiCNN*>(psi_ref->p.get());
    }
    *out = new angpu_psi_s{std::unique_ptr<Psi>(new PsiClassical(num_sites, order, num_ops, ops.data(), cp(params), num_own_params, ref, c2(lp)))};
    API_END
}
int angpu_psi_copy(angpu_psi_t psi, angpu_psi_t* out) { API_BEGIN NOTNULL(psi); NOTNULL(out); *out = new angpu_psi_s{std::unique_ptr<Psi>(psi->p->clone())}; API_END }
int angpu_psi_destroy(angpu_psi_t psi) { API_BEGIN delete psi; API_END }
int angpu_psi_kind(angpu_psi_t psi, int* out) { API_BEGIN NOTNULL(psi); NOTNULL(out); *out = (int)psi->p->kind; API_END }
int angpu_psi_num_sites(angpu_psi_t psi, unsigned* out) { API_BEGIN NOTNULL(psi); NOTNULL(out); *out = psi->p->N; API_END }
int angpu_psi_num_params(angpu_psi_t psi, unsigned* out) { API_BEGIN NOTNULL(psi); NOTNULL(out); *out = psi->p->P; API_END }
int angpu_psi_get_params(angpu_psi_t psi, double* out) { API_BEGIN NOTNULL(psi); NOTNULL(out); psi->p->get_params(cp(out)); API_END }
int angpu_psi_set_params(angpu_psi_t psi, const double* in) { API_BEGIN NOTNULL(psi); NOTNULL(in); psi->p->set_params(cp(in)); API_END }
int angpu_psi_get_log_prefactor(angpu_psi_t psi, double out[2]) { API_BEGIN NOTNULL(psi); NOTNULL(out); out[0] = psi->p->lp.re; out[1] = psi->p->lp.im; API_END }
int angpu_psi_set_log_prefactor(angpu_psi_t psi, const double in[2]) { API_BEGIN NOTNULL(psi); NOTNULL(in); psi->p->set_log_prefactor(c2(in)); API_END }

// ---- ensembles
int angpu_es_create(unsigned num_sites, angpu_ensemble_t* out) {
    API_BEGIN NOTNULL(out);
    ANGPU_REQUIRE(num_sites >= 1 && num_sites <= 40u, "ExactSummation: 1 <= num_sites <= 40");
    auto* e = new angpu_ensemble_s(); e->e.is_mc = false; e->e.num_sites = num_sites; *out = e;
    API_END
}
int angpu_mc_create(unsigned long long num_samples, unsigned num_sweeps, unsigned num_therm, unsigned num_chains, uint64_t seed, angpu_ensemble_t* out) {
    API_BEGIN NOTNULL(out);
    ANGPU_REQUIRE(num_chains >= 1, "MonteCarlo: num_markov_chains must be >= 1");
    ANGPU_REQUIRE(num_samples >= 1, "MonteCarlo: num_samples must be >= 1");
    auto* e = new angpu_ensemble_s(); e->e.is_mc = true;
    e->e.num_samples = num_samples; e->e.num_sweeps = num_sweeps; e->e.num_therm = num_therm; e->e.num_chains = num_chains; e->e.seed = seed;
    *out = e;
    API_END
}
// MonteCarloPaulis / ExactSummationPaulis (pyANNonGPU/main.cpp.template:369-377, 400-405): configurations are Pauli strings
int angpu_es_paulis_create(unsigned num_sites, angpu_ensemble_t* out) {
    API_BEGIN NOTNULL(out);
    ANGPU_REQUIRE(num_sites >= 1 && num_sites <= 20u, "ExactSummationPaulis: 1 <= num_sites <= 20");
    auto* e = new angpu_ensemble_s(); e->e.is_mc = false; e->e.paulis = true; e->e.num_sites = num_sites; *out = e;
    API_END
}
int angpu_mc_paulis_create(unsigned long long num_samples, unsigned num_sweeps, unsigned num_therm, unsigned num_chains, uint64_t seed, angpu_ensemble_t* out) {
    API_BEGIN NOTNULL(out);
    ANGPU_REQUIRE(num_chains >= 1, "MonteCarlo: num_markov_chains must be >= 1");
    ANGPU_REQUIRE(num_samples >= 1, "MonteCarlo: num_samples must be >= 1");
    auto* e = new angpu_ensemble_s(); e->e.is_mc = true; e->e.paulis = true;
    e->e.num_samples = num_samples; e->e.num_sweeps = num_sweeps; e->e.num_therm = num_therm; e->e.num_chains = num_chains; e->e.seed = seed;
    *out = e;
    API_END
}
int angpu_ensemble_copy(angpu_ensemble_t ens, angpu_ensemble_t* out) {
    API_BEGIN NOTNULL(ens); NOTNULL(out);
    auto* e = new angpu_ensemble_s(); const Ensemble& s = ens->e;
    e->e.is_mc = s.is_mc; e->e.paulis = s.paulis; e->e.num_sites = s.num_sites; e->e.num_samples = s.num_samples; e->e.num_sweeps = s.num_sweeps;
    e->e.num_therm = s.num_therm; e->e.num_chains = s.num_chains; e->e.call = s.call; e->e.seed = s.seed; e->e.rank = s.rank; e->e.world = s.world;
    *out = e;
    API_END
}
int angpu_ensemble_destroy(angpu_ensemble_t ens) { API_BEGIN delete ens; API_END }
int angpu_ensemble_num_steps(angpu_ensemble_t ens, unsigned long long* out) { API_BEGIN NOTNULL(ens); NOTNULL(out); *out = ens->e.num_steps(); API_END }
int angpu_ensemble_local_steps(angpu_ensemble_t ens, unsigned long long* out) { API_BEGIN NOTNULL(ens); NOTNULL(out); *out = ens->e.local_steps(); API_END }
int angpu_ensemble_set_shard(angpu_ensemble_t ens, unsigned rank, unsigned world) {
    API_BEGIN NOTNULL(ens);
    ANGPU_REQUIRE(world >= 1 && rank < world, "set_shard: need rank < world");
    ens->e.rank = rank; ens->e.world = world;
    API_END
}
int angpu_mc_acceptance(angpu_ensemble_t ens, unsigned long long out[2]) { API_BEGIN NOTNULL(ens); NOTNULL(out); ens->e.acceptance(out); API_END }
int angpu_mc_get_call_index(angpu_ensemble_t ens, unsigned* out) { API_BEGIN NOTNULL(ens); NOTNULL(out); *out = ens->e.call; API_END }
int angpu_mc_set_call_index(angpu_ensemble_t ens, unsigned call_index) { API_BEGIN NOTNULL(ens); ens->e.call = call_index; API_END }
int angpu_mc_counters(angpu_ensemble_t ens, unsigned long long out[4]) {
    API_BEGIN NOTNULL(ens); NOTNULL(out);
    out[0] = out[1] = out[2] = out[3] = 0;
    if(ens->e.d_acc_rej.n >= 4) ens->e.d_acc_rej.download(out, 4);
    API_END
}
int angpu_ensemble_sample(angpu_ensemble_t ens, angpu_psi_t psi, uint64_t* confs_out, double* log_psi_out) {
    API_BEGIN NOTNULL(ens); NOTNULL(psi);
    SampleSet S;
    ens->e.generate(*psi->p, S);
    if(confs_out) S.conf.download(confs_out, S.ns * S.words);
    if(log_psi_out) S.log_psi.download(cp(log_psi_out), S.ns);
    API_END
}

// ---- probes / vectors
int angpu_log_psi_s(angpu_psi_t psi, const uint64_t* conf, double out[2]) {
    API_BEGIN NOTNULL(psi); NOTNULL(conf); NOTNULL(out);
    const cplx r = log_psi_s(*psi->p, conf); out[0] = r.re; out[1] = r.im;
    API_END
}
int angpu_psi_O_k(angpu_psi_t psi, const uint64_t* conf, double* out) { API_BEGIN NOTNULL(psi); NOTNULL(conf); NOTNULL(out); psi_O_k(*psi->p, conf, cp(out)); API_END }
int angpu_log_psi_vector(angpu_psi_t psi, angpu_ensemble_t ens, double* out) { API_BEGIN NOTNULL(psi); NOTNULL(ens); NOTNULL(out); log_psi_vector(*psi->p, ens->e, cp(out), false); API_END }
int angpu_psi_vector(angpu_psi_t psi, angpu_ensemble_t ens, double* out) { API_BEGIN NOTNULL(psi); NOTNULL(ens); NOTNULL(out); log_psi_vector(*psi->p, ens->e, cp(out), true); API_END }
int angpu_log_psi_mean(angpu_psi_t psi, angpu_ensemble_t ens, double out[2]) {
    API_BEGIN NOTNULL(psi); NOTNULL(ens); NOTNULL(out);
    const cplx r = log_psi_mean(*psi->p, ens->e); out[0] = r.re; out[1] = r.im;
    API_END
}
int angpu_psi_norm(angpu_psi_t psi, angpu_ensemble_t es, double* out) { API_BEGIN NOTNULL(psi); NOTNULL(es); NOTNULL(out); *out = psi_norm(*psi->p, es->e); API_END }
int angpu_psi_O_k_vector(angpu_psi_t psi, angpu_ensemble_t es, double* out) { API_BEGIN NOTNULL(psi); NOTNULL(es); NOTNULL(out); psi_O_k_vector(*psi->p, es->e, cp(out)); API_END }
int angpu_apply_operator(angpu_psi_t psi, angpu_operator_t op, angpu_ensemble_t ens, double* out) {
    API_BEGIN NOTNULL(psi); NOTNULL(op); NOTNULL(ens); NOTNULL(out); apply_operator(*psi->p, *op->p, ens->e, cp(out)); API_END
}
int angpu_local_energies(angpu_psi_t psi, angpu_operator_t op, const uint64_t* confs, unsigned long long ns, double* log_psi_out, double* eloc_out) {
    API_BEGIN NOTNULL(psi); NOTNULL(op); if(ns) NOTNULL(confs);
    local_energies(*psi->p, *op->p, confs, (size_t)ns, log_psi_out ? cp(log_psi_out) : nullptr, eloc_out ? cp(eloc_out) : nullptr);
    API_END
}

// ---- ExpectationValue
int angpu_expval_create(angpu_expval_t* out) { API_BEGIN NOTNULL(out); *out = new angpu_expval_s(); API_END }
int angpu_expval_destroy(angpu_expval_t ev) { API_BEGIN delete ev; API_END }
int angpu_expectation(angpu_expval_t ev, angpu_operator_t op, angpu_psi_t psi, angpu_ensemble_t ens, double out[2]) {
    API_BEGIN NOTNULL(ev); NOTNULL(op); NOTNULL(psi); NOTNULL(ens); NOTNULL(out);
    const cplx r = ev->ev.value(*op->p, *psi->p, ens->e); out[0] = r.re; out[1] = r.im;
    API_END
}
int angpu_expectation_many(angpu_expval_t ev, unsigned num_ops, const angpu_operator_t* ops, angpu_psi_t psi, angpu_ensemble_t ens, double* out) {
    API_BEGIN NOTNULL(ev); NOTNULL(psi); NOTNULL(ens);
    if(num_ops) { NOTNULL(ops); NOTNULL(out); }
    // one set of samples for all operators, as the reference's single foreach (ExpectationValue.cu.template:102-124)
    SampleSet& S = ev->ev.S;
    ens->e.generate(*psi->p, S);
    DevBuf<double> d((size_t)4 * std::max(1u, num_ops));
    for(unsigned i = 0; i < num_ops; i++) {
        NOTNULL(ops[i]);
        psi->p->eloc(*ops[i]->p, S);
        scalar_sums_eloc(S, d.p + 4 * i);
    }
    allreduce_sum(d.p, (size_t)4 * num_ops);
    std::vector<double> h((size_t)4 * num_ops); d.download(h.data(), h.size());
    for(unsigned i = 0; i < num_ops; i++) { out[2 * i] = h[4 * i]; out[2 * i + 1] = h[4 * i + 1]; }
    API_END
}
int angpu_expectation_reweighted(angpu_expval_t ev, angpu_operator_t op, angpu_psi_t psi, angpu_psi_t psi_sampling, angpu_ensemble_t ens, double out[2]) {
    API_BEGIN NOTNULL(ev); NOTNULL(op); NOTNULL(psi); NOTNULL(psi_sampling); NOTNULL(ens); NOTNULL(out);
    const cplx r = ev->ev.value_reweighted(*op->p, *psi->p, *psi_sampling->p, ens->e); out[0] = r.re; out[1] = r.im;
    API_END
}
int angpu_exp_sigma_z(angpu_expval_t ev, angpu_operator_t op, angpu_psi_t psi, angpu_ensemble_t ens, double out[2]) {
    API_BEGIN NOTNULL(ev); NOTNULL(op); NOTNULL(psi); NOTNULL(ens); NOTNULL(out);
    const cplx r = ev->ev.exp_sigma_z(*op->p, *psi->p, ens->e); out[0] = r.re; out[1] = r.im;
    API_END
}
int angpu_fluctuation(angpu_expval_t ev, angpu_operator_t op, angpu_psi_t psi, angpu_ensemble_t ens, double* fluctuation_out, double mean_out[2]) {
    API_BEGIN NOTNULL(ev); NOTNULL(op); NOTNULL(psi); NOTNULL(ens); NOTNULL(fluctuation_out); NOTNULL(mean_out);
    cplx m; ev->ev.fluctuation(*op->p, *psi->p, ens->e, *fluctuation_out, m); mean_out[0] = m.re; mean_out[1] = m.im;
    API_END
}
// gradient_k = <O_k* E_loc> - <E_loc><O_k*> (ExpectationValue.cu.template:220-275) == TDVP's F_k (TDVP.cu.template:300)
int angpu_gradient(angpu_expval_t ev, angpu_operator_t op, angpu_psi_t psi, angpu_ensemble_t ens, double* gradient_out, double mean_out[2]) {
    API_BEGIN NOTNULL(ev); NOTNULL(op); NOTNULL(psi); NOTNULL(ens); NOTNULL(gradient_out); NOTNULL(mean_out);
    if(!ev->grad || ev->grad->P != psi->p->P) ev->grad.reset(new TDVP(psi->p->P));
    ev->grad->eval_F(*op->p, *psi->p, ens->e);
    ev->grad->F.download(cp(gradient_out), psi->p->P);
    mean_out[0] = ev->grad->E.re; mean_out[1] = ev->grad->E.im;
    API_END
}

// ---- HilbertSpaceDistance
int angpu_hsd_create(unsigned num_params, angpu_hsd_t* out) { API_BEGIN NOTNULL(out); *out = new angpu_hsd_s{std::unique_ptr<HilbertSpaceDistance>(new HilbertSpaceDistance(num_params))}; API_END }
int angpu_hsd_destroy(angpu_hsd_t hsd) { API_BEGIN delete hsd; API_END }
int angpu_hsd_distance(angpu_hsd_t hsd, angpu_psi_t psi, angpu_psi_t psi_prime, angpu_operator_t op, int is_unitary, angpu_ensemble_t ens, double* distance_out) {
    API_BEGIN NOTNULL(hsd); NOTNULL(psi); NOTNULL(psi_prime); NOTNULL(op); NOTNULL(ens); NOTNULL(distance_out);
    *distance_out = hsd->h->distance(*psi->p, *psi_prime->p, *op->p, is_unitary != 0, ens->e);
    API_END
}
int angpu_hsd_gradient(angpu_hsd_t hsd, angpu_psi_t psi, angpu_psi_t psi_prime, angpu_operator_t op, int is_unitary, angpu_ensemble_t ens,
                       float nu, double* gradient_out, double* distance_out) {
    API_BEGIN NOTNULL(hsd); NOTNULL(psi); NOTNULL(psi_prime); NOTNULL(op); NOTNULL(ens); NOTNULL(gradient_out); NOTNULL(distance_out);
    *distance_out = hsd->h->gradient(cp(gradient_out), *psi->p, *psi_prime->p, *op->p, is_unitary != 0, ens->e, nu);
    API_END
}

// ---- KullbackLeibler
int angpu_kl_create(unsigned num_params, angpu_kl_t* out) { API_BEGIN NOTNULL(out); *out = new angpu_kl_s{std::unique_ptr<KullbackLeibler>(new KullbackLeibler(num_params))}; API_END }
int angpu_kl_destroy(angpu_kl_t kl) { API_BEGIN delete kl; API_END }
int angpu_kl_set_log_psi_scale(angpu_kl_t kl, double scale) { API_BEGIN NOTNULL(kl); kl->k->log_psi_scale = scale; API_END }
int angpu_kl_get_state(angpu_kl_t kl, double out[4]) {
    API_BEGIN NOTNULL(kl); NOTNULL(out);
    out[0] = kl->k->total_weight; out[1] = kl->k->mean_deviation.re; out[2] = kl->k->mean_deviation.im; out[3] = kl->k->log_psi_scale;
    API_END
}
int angpu_kl_value(angpu_kl_t kl, angpu_psi_t psi, angpu_psi_t psi_prime, angpu_ensemble_t ens, double threshold, double* value_out) {
    API_BEGIN NOTNULL(kl); NOTNULL(psi); NOTNULL(psi_prime); NOTNULL(ens); NOTNULL(value_out);
    *value_out = kl->k->value(*psi->p, *psi_prime->p, ens->e, threshold);
    API_END
}
int angpu_kl_gradient(angpu_kl_t kl, angpu_psi_t psi, angpu_psi_t psi_prime, angpu_ensemble_t ens, double nu, double threshold,
                      double* gradient_out, double* value_out) {
    API_BEGIN NOTNULL(kl); NOTNULL(psi); NOTNULL(psi_prime); NOTNULL(ens); NOTNULL(gradient_out); NOTNULL(value_out);
    *value_out = kl->k->gradient(cp(gradient_out), *psi->p, *psi_prime->p, ens->e, nu, threshold);
    API_END
}
int angpu_kl_gradient_with_noise(angpu_kl_t kl, angpu_psi_t psi, angpu_psi_t psi_prime, angpu_ensemble_t ens, double nu, double threshold,
                                 double* gradient_out, double* noise_out, double* value_out) {
    API_BEGIN NOTNULL(kl); NOTNULL(psi); NOTNULL(psi_prime); NOTNULL(ens); NOTNULL(gradient_out); NOTNULL(noise_out); NOTNULL(value_out);
    *value_out = kl->k->gradient_with_noise(cp(gradient_out), noise_out, *psi->p, *psi_prime->p, ens->e, nu, threshold);
    API_END
}

// ---- TDVP
int angpu_tdvp_create(unsigned num_params, angpu_tdvp_t* out) { API_BEGIN NOTNULL(out); *out = new angpu_tdvp_s{std::unique_ptr<TDVP>(new TDVP(num_params))}; API_END }
int angpu_tdvp_destroy(angpu_tdvp_t tdvp) { API_BEGIN delete tdvp; API_END }
int angpu_tdvp_eval_reweighted(angpu_tdvp_t tdvp, angpu_operator_t op, angpu_psi_t psi, angpu_psi_t psi_sampling, angpu_ensemble_t ens) {
    API_BEGIN NOTNULL(tdvp); NOTNULL(op); NOTNULL(psi); NOTNULL(ens);
    Psi* sampling = psi_sampling ? psi_sampling->p.get() : nullptr;
    if(!sampling) {
        ANGPU_REQUIRE(psi->p->kind == Psi::CLASSICAL, "eval_with_psi_ref: psi_sampling may be NULL only for a PsiClassical (its psi_ref is used)");
        sampling = static_cast<PsiClassical*>(psi->p.get())->sampling_ref();
    }
    tdvp->t->eval(*op->p, *psi->p, ens->e, true, sampling);
    API_END
}
int angpu_tdvp_eval(angpu_tdvp_t tdvp, angpu_operator_t op, angpu_psi_t psi, angpu_ensemble_t ens) {
    API_BEGIN NOTNULL(tdvp); NOTNULL(op); NOTNULL(psi); NOTNULL(ens); tdvp->t->eval(*op->p, *psi->p, ens->e, true); API_END
}
int angpu_tdvp_eval_tol(angpu_tdvp_t tdvp, angpu_operator_t op, angpu_psi_t psi, angpu_ensemble_t ens, double s_tolerance) {
    API_BEGIN NOTNULL(tdvp); NOTNULL(op); NOTNULL(psi); NOTNULL(ens);
    ANGPU_REQUIRE(s_tolerance == 0.0 || s_tolerance >= 1e-5, "s_tolerance: 0 (exact fp64 S) or >= 1e-5 (tensor-core S)");
    if(s_tolerance == 0.0) tdvp->t->eval(*op->p, *psi->p, ens->e, true);
    else { tdvp->t->eval(*op->p, *psi->p, ens->e, false); tdvp->t->build_S_tensorcore(); }
    API_END
}
int angpu_tdvp_eval_F(angpu_tdvp_t tdvp, angpu_operator_t op, angpu_psi_t psi, angpu_ensemble_t ens) {
    API_BEGIN NOTNULL(tdvp); NOTNULL(op); NOTNULL(psi); NOTNULL(ens); tdvp->t->eval_F(*op->p, *psi->p, ens->e); API_END
}
int angpu_tdvp_get_S(angpu_tdvp_t tdvp, double* out) {
    API_BEGIN NOTNULL(tdvp); NOTNULL(out);
    TDVP& t = *tdvp->t;
    ANGPU_REQUIRE(t.evaluated, "TDVP: call eval first");
    if(!t.have_S) t.build_S();
    t.Smat.download(cp(out), (size_t)t.P * t.P);
    API_END
}
int angpu_tdvp_get_F(angpu_tdvp_t tdvp, double* out) { API_BEGIN NOTNULL(tdvp); NOTNULL(out); ANGPU_REQUIRE(tdvp->t->evaluated, "TDVP: call eval first"); tdvp->t->F.download(cp(out), tdvp->t->P); API_END }
int angpu_tdvp_get_O_k(angpu_tdvp_t tdvp, double* out) { API_BEGIN NOTNULL(tdvp); NOTNULL(out); ANGPU_REQUIRE(tdvp->t->evaluated, "TDVP: call eval first"); tdvp->t->packed.download(cp(out), tdvp->t->P, 2); API_END }
int angpu_tdvp_get_scalars(angpu_tdvp_t tdvp, double out[5]) {
    API_BEGIN NOTNULL(tdvp); NOTNULL(out);
    const TDVP& t = *tdvp->t;
    out[0] = t.E.re; out[1] = t.E.im; out[2] = t.E2; out[3] = t.var_H(); out[4] = t.total_weight;
    API_END
}
int angpu_tdvp_num_local_samples(angpu_tdvp_t tdvp, unsigned long long* out) { API_BEGIN NOTNULL(tdvp); NOTNULL(out); *out = tdvp->t->S.ns; API_END }
int angpu_tdvp_get_O_k_samples(angpu_tdvp_t tdvp, double* out) {
    API_BEGIN NOTNULL(tdvp); NOTNULL(out);
    TDVP& t = *tdvp->t;
    t.ensure_dense_O(t.last_psi);
    t.O.download(cp(out), t.S.ns * (size_t)t.P);
    API_END
}
int angpu_tdvp_get_weights(angpu_tdvp_t tdvp, double* out) { API_BEGIN NOTNULL(tdvp); NOTNULL(out); tdvp->t->S.weight.download(out, tdvp->t->S.ns); API_END }
int angpu_tdvp_get_E_local_samples(angpu_tdvp_t tdvp, double* out) { API_BEGIN NOTNULL(tdvp); NOTNULL(out); tdvp->t->S.eloc.download(cp(out), tdvp->t->S.ns); API_END }
int angpu_tdvp_S_dot_vector(angpu_tdvp_t tdvp, const double* vec, double* out) { API_BEGIN NOTNULL(tdvp); NOTNULL(vec); NOTNULL(out); tdvp->t->S_dot_vector(cp(vec), cp(out)); API_END }
int angpu_tdvp_solve_cg(angpu_tdvp_t tdvp, double tol, unsigned max_iter, double shift_abs, double shift_rel, const double rhs_phase[2],
                        double* x_out, unsigned* iterations_out, double* rel_residual_out) {
    API_BEGIN NOTNULL(tdvp); NOTNULL(rhs_phase);
    double rr = 0.0;
    const int it = tdvp->t->solve_cg(tol, max_iter, shift_abs, shift_rel, c2(rhs_phase), cp(x_out), &rr);
    if(iterations_out) *iterations_out = (unsigned)it;
    if(rel_residual_out) *rel_residual_out = rr;
    API_END
}
int angpu_tdvp_build_S_tensorcore(angpu_tdvp_t tdvp) { API_BEGIN NOTNULL(tdvp); tdvp->t->build_S_tensorcore(); API_END }
int angpu_tdvp_set_tensorcore_products(angpu_tdvp_t tdvp, int enable) { API_BEGIN NOTNULL(tdvp); tdvp->t->tc_products = enable < 0 ? -1 : (enable != 0 ? 1 : 0); API_END }
int angpu_tdvp_set_profile(angpu_tdvp_t tdvp, int enable) { API_BEGIN NOTNULL(tdvp); tdvp->t->profile = enable != 0; API_END }
int angpu_tdvp_phase_ms(angpu_tdvp_t tdvp, double out[6]) { API_BEGIN NOTNULL(tdvp); NOTNULL(out); for(int i = 0; i < 6; i++) out[i] = tdvp->t->phase_ms[i]; API_END }
int angpu_measure_fp64_tflops(double* out) { API_BEGIN NOTNULL(out); *out = measure_fp64_tflops(); API_END }
int angpu_tdvp_solve_dense(angpu_tdvp_t tdvp, double shift_abs, double shift_rel, const double rhs_phase[2], double* x_out) {
    API_BEGIN NOTNULL(tdvp); NOTNULL(rhs_phase); tdvp->t->solve_dense(shift_abs, shift_rel, c2(rhs_phase), cp(x_out)); API_END
}
int angpu_hpd_solve(unsigned n, const double* A, const double* b, double* x_out) {
    API_BEGIN NOTNULL(A); NOTNULL(b); NOTNULL(x_out);
    ANGPU_REQUIRE(n >= 1, "angpu_hpd_solve: n >= 1");
    DevBuf<cplx> dA, db, work; DevBuf<int> info(2);
    dA.upload(cp(A), (size_t)n * n); db.upload(cp(b), n);
    cholesky_solve(dA.p, db.p, n, info.p, work);
    int hinfo = 0; info.download(&hinfo, 1);
    if(hinfo != 0) throw Error("angpu_hpd_solve: the matrix is not positive definite (pivot " + std::to_string(hinfo) + ")");
    db.download(cp(x_out), n);
    API_END
}
int angpu_tdvp_apply_update(angpu_tdvp_t tdvp, angpu_psi_t psi, const double alpha[2]) {
    API_BEGIN NOTNULL(tdvp); NOTNULL(psi); NOTNULL(alpha);
    ANGPU_REQUIRE(tdvp->t->last_x != nullptr, "TDVP: no solution on the device (call solve_cg / solve_dense first)");
    ANGPU_REQUIRE(psi->p->P == tdvp->t->P, "TDVP: num_params differs from psi's");
    psi->p->add_params_dev(tdvp->t->last_x, c2(alpha));
    API_END
}

#pragma GCC visibility pop
} // extern "C"

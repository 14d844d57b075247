/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Plain-C (C11 + OpenMP) restatement of the variational-Monte-Carlo hot path of
 * heikoburau/ANNonGPU, generalised to multi-word spin / Pauli masks (N <= 256) and to hidden
 * layers wider than the reference's compiled-in limits, so that it can check the CUDA path at
 * the BASELINE.json shapes the unmodified reference cannot run.
 *
 * Pinning: tests/test_oracle_pinned.py checks every function below against the compiled,
 * unmodified reference (oracle/_ref/liboracle_ref.so) where that is present, and against the
 * golden vectors in tests/golden/ (generated from the compiled reference by
 * tests/golden/make_golden.py) everywhere else.
 *
 * Each function cites the reference file:line whose arithmetic it follows (paths relative to
 * the reference root).  Complex numbers cross the C ABI as interleaved (re, im) doubles.
 *
 * The Monte-Carlo random stream is NOT the reference's (XORWOW on GPU, mt19937 on CPU — not
 * reproducible across its own back-ends): it is the counter-based Philox4x32-10 stream that the
 * CUDA path uses, keyed (seed | call, tag | chain | step), so that CPU and GPU chains can be
 * compared configuration by configuration.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cx;

#define MAXW 4          /* 64-bit words per configuration: N <= 256 */
#define MAX_LAYERS 8

/* ------------------------------------------------------------------ activations
 * include/quantum_state/psi_functions.hpp:43-48 (my_logcosh), :106-115 (my_tanh). */
static inline cx act_lc(cx z, unsigned layer) {
    const cx z2 = z * z, z4 = z2 * z2;
    if(layer == 0u) return 0.5 * z2 - (1.0 / 12.0) * z4 + (1.0 / 45.0) * z4 * z2;
    return z - (1.0 / 3.0) * z2 * z + (2.0 / 15.0) * z4 * z;
}
static inline cx act_th(cx z, unsigned layer) {
    const cx z2 = z * z, z4 = z2 * z2;
    if(layer == 0u) return z - (1.0 / 3.0) * z2 * z + (2.0 / 15.0) * z4 * z;
    return 1.0 - z2 + (2.0 / 3.0) * z4;
}
void port_activation(double re, double im, unsigned layer, double* lc_out, double* th_out) {
    const cx a = act_lc(re + im * I, layer), b = act_th(re + im * I, layer);
    lc_out[0] = creal(a); lc_out[1] = cimag(a); th_out[0] = creal(b); th_out[1] = cimag(b);
}

/* ------------------------------------------------------------------ spins
 * include/basis/Spins.h:104-108 (operator[]: bit set <=> +1), :291-296 (enumerate: index == mask). */
static inline unsigned words_for_c(unsigned n) { return (n + 63u) / 64u; }
static inline double spin_at(const uint64_t* conf, unsigned i) {
    return (conf[i >> 6] >> (i & 63u)) & 1u ? 1.0 : -1.0;
}
static inline int conf_equal(const uint64_t* a, const uint64_t* b, unsigned words) {
    for(unsigned w = 0; w < words; w++) if(a[w] != b[w]) return 0;
    return 1;
}

/* ------------------------------------------------------------------ operator
 * include/operator/Operator.hpp:25-31 (SoA of coefficients + strings). */
typedef struct {
    unsigned  n, words;
    cx*       coef;
    uint64_t* a;      /* n x words */
    uint64_t* b;
} op_t;

op_t* port_op_create(unsigned n, const double* coeffs, const uint64_t* a, const uint64_t* b, unsigned words) {
    op_t* op = (op_t*)calloc(1, sizeof(op_t));
    op->n = n; op->words = words;
    op->coef = (cx*)malloc(sizeof(cx) * (n ? n : 1));
    op->a = (uint64_t*)malloc(sizeof(uint64_t) * (n ? n : 1) * words);
    op->b = (uint64_t*)malloc(sizeof(uint64_t) * (n ? n : 1) * words);
    for(unsigned i = 0; i < n; i++) op->coef[i] = coeffs[2 * i] + coeffs[2 * i + 1] * I;
    memcpy(op->a, a, sizeof(uint64_t) * n * words);
    memcpy(op->b, b, sizeof(uint64_t) * n * words);
    return op;
}
void port_op_destroy(op_t* op) { free(op->coef); free(op->a); free(op->b); free(op); }

/* PauliString::apply(Spins) and complex_prefactor, include/basis/PauliString.hpp:193-208, 242-255.
 * (a,b): X=(1,0) Y=(0,1) Z=(1,1).  prefactor (-i)^{n_Y}; sign (-1)^{popc(~s & (Y|Z))}; flip mask a^b. */
static inline cx pauli_apply(const uint64_t* a, const uint64_t* b, const uint64_t* conf, unsigned words, uint64_t* out) {
    unsigned ny = 0, nneg = 0;
    for(unsigned w = 0; w < words; w++) {
        const uint64_t is_y = ~a[w] & b[w], is_z = a[w] & b[w];
        ny += (unsigned)__builtin_popcountll(is_y);
        nneg += (unsigned)__builtin_popcountll(~conf[w] & (is_z | is_y));
        out[w] = conf[w] ^ (a[w] ^ b[w]);
    }
    cx f = 1.0;
    if((ny & 3u) > 1u) f *= -1.0;
    if(ny & 1u) f *= -1.0 * I;
    if(nneg & 1u) f *= -1.0;
    return f;
}
void port_pauli_apply(const uint64_t* a, const uint64_t* b, const uint64_t* conf, unsigned words, double* coeff_out, uint64_t* conf_out) {
    const cx f = pauli_apply(a, b, conf, words, conf_out);
    coeff_out[0] = creal(f); coeff_out[1] = cimag(f);
}

/* StandartOperator::fast_local_energy, include/operator/Operator.hpp:125-136. */
static cx op_fast_local_energy(const op_t* op, const uint64_t* conf) {
    cx r = 0.0; uint64_t tmp[MAXW];
    for(unsigned n = 0; n < op->n; n++)
        r += op->coef[n] * pauli_apply(op->a + n * op->words, op->b + n * op->words, conf, op->words, tmp);
    return r;
}

/* ------------------------------------------------------------------ Pauli-string basis (density-matrix ensembles)
 * A configuration is a Pauli string x = (a, b) over num_sites sites (I=(0,0) X=(1,0) Y=(0,1) Z=(1,1)).  The network sees it
 * through 3 num_sites input units, unit 3 s + t = +1 iff x[s] - 1 == t, else -1 (PauliString::network_unit_at,
 * include/basis/PauliString.hpp:84-90; PsiDeep::update_angles, PsiDeep.hpp:283-307).  Stored form here and on the GPU: the
 * "units" bit mask (bit 3 s + t set iff the unit is +1), on which the spin-basis network code applies unchanged. */
void port_paulis_to_units(const uint64_t* a, const uint64_t* b, unsigned num_sites, uint64_t* units) {
    for(unsigned w = 0; w < words_for_c(3u * num_sites); w++) units[w] = 0;
    for(unsigned s = 0; s < num_sites; s++) {
        const unsigned t = (unsigned)((a[s >> 6] >> (s & 63u)) & 1u) | ((unsigned)((b[s >> 6] >> (s & 63u)) & 1u) << 1);
        if(t) { const unsigned u = 3u * s + t - 1u; units[u >> 6] |= 1ull << (u & 63u); }
    }
}
void port_units_to_paulis(const uint64_t* units, unsigned num_sites, uint64_t* a, uint64_t* b) {
    for(unsigned w = 0; w < words_for_c(num_sites); w++) a[w] = b[w] = 0;
    for(unsigned s = 0; s < num_sites; s++) {
        unsigned t = 0;
        for(unsigned k = 0; k < 3u; k++) { const unsigned u = 3u * s + k; if((units[u >> 6] >> (u & 63u)) & 1u) t = k + 1u; }
        if(t & 1u) a[s >> 6] |= 1ull << (s & 63u);
        if(t & 2u) b[s >> 6] |= 1ull << (s & 63u);
    }
}
/* PauliString::apply(PauliString), include/basis/PauliString.hpp:257-277: P x = factor * (P xor x) */
static inline cx pauli_mul(const uint64_t* Pa, const uint64_t* Pb, const uint64_t* xa, const uint64_t* xb, unsigned words,
                           uint64_t* oa, uint64_t* ob) {
    unsigned nneg = 0, neps = 0;
    for(unsigned w = 0; w < words; w++) {
        const uint64_t px = Pa[w] & ~Pb[w], py = ~Pa[w] & Pb[w], pz = Pa[w] & Pb[w];
        const uint64_t xx = xa[w] & ~xb[w], xy = ~xa[w] & xb[w], xz = xa[w] & xb[w];
        nneg += (unsigned)__builtin_popcountll((px & xz) | (py & xx) | (pz & xy));
        neps += (unsigned)__builtin_popcountll((Pa[w] | Pb[w]) & (xa[w] | xb[w]) & ((Pa[w] ^ xa[w]) | (Pb[w] ^ xb[w])));
        oa[w] = Pa[w] ^ xa[w]; ob[w] = Pb[w] ^ xb[w];
    }
    cx f = 1.0;
    if(nneg & 1u) f *= -1.0;
    if((neps & 3u) > 1u) f *= -1.0;
    if(neps & 1u) f *= -1.0 * I;
    return f;
}
void port_pauli_mul(const uint64_t* Pa, const uint64_t* Pb, const uint64_t* xa, const uint64_t* xb, unsigned words,
                    double* coeff_out, uint64_t* a_out, uint64_t* b_out) {
    const cx f = pauli_mul(Pa, Pb, xa, xb, words, a_out, b_out);
    coeff_out[0] = creal(f); coeff_out[1] = cimag(f);
}
/* PauliString::enumerate, :33-38: site s takes type (index >> 2 s) & 3 */
void port_paulis_enumerate(uint64_t index, unsigned num_sites, uint64_t* a, uint64_t* b) {
    for(unsigned w = 0; w < words_for_c(num_sites); w++) a[w] = b[w] = 0;
    for(unsigned s = 0; s < num_sites && s < 32u; s++) {
        if((index >> (2u * s)) & 1u) a[s >> 6] |= 1ull << (s & 63u);
        if((index >> (2u * s + 1u)) & 1u) b[s >> 6] |= 1ull << (s & 63u);
    }
}

/* ------------------------------------------------------------------ wavefunctions */
enum { K_RBM = 0, K_DEEP = 1, K_CNN = 2, K_CLASSICAL = 3 };

typedef struct {
    unsigned size, conn, rhs_conn, begin_params, begin_deep;
    unsigned *lhs_connections, *rhs_connections;   /* conn x size ; size x rhs_conn */
    cx *lhs_weights, *rhs_weights, *biases;
} deep_layer_t;

typedef struct { unsigned begin_params; } cnn_link_t;
typedef struct {
    unsigned num_channels, num_links, connectivity[3], vol, angle_offset;
    cnn_link_t links[64];
} cnn_layer_t;

typedef struct psi_s {
    int       kind;
    unsigned  N, words, num_params;
    unsigned  num_sites;   /* physical sites; N == 3 num_sites marks a PsiDeep on the Pauli-string basis (one input unit per site and Pauli type) */
    cx        log_prefactor;
    /* RBM  (include/quantum_state/PsiRBM.hpp:42-67) */
    unsigned  M; cx* W; cx final_weight;
    /* Deep (include/quantum_state/PsiDeep.hpp:71-138) */
    unsigned  num_layers, width, num_deep; deep_layer_t layers[MAX_LAYERS]; cx *input_weights, *final_weights;
    /* CNN  (include/quantum_state/PsiCNN.hpp:33-92) */
    unsigned  extent[3], cnn_layers, num_sym, num_angles; unsigned* sym; cnn_layer_t cl[MAX_LAYERS]; cx* params; double final_factor;
    /* Classical (include/quantum_state/PsiClassical.hpp:48-70) */
    unsigned  order, num_ops, num_own; op_t** ops; cx* cparams; struct psi_s* ref;
} psi_t;

static unsigned words_for(unsigned N) { return (N + 63u) / 64u; }

psi_t* port_rbm_create(unsigned N, unsigned M, const double* W, const double* fw, const double* lp) {
    psi_t* p = (psi_t*)calloc(1, sizeof(psi_t));
    p->kind = K_RBM; p->N = N; p->M = M; p->words = words_for(N); p->num_params = N * M;
    p->W = (cx*)malloc(sizeof(cx) * N * M);
    for(unsigned k = 0; k < N * M; k++) p->W[k] = W[2 * k] + W[2 * k + 1] * I;
    p->final_weight = fw[0] + fw[1] * I; p->log_prefactor = lp[0] + lp[1] * I;
    return p;
}

/* rhs tables: source/quantum_state/PsiDeep.cu:214-242 (compile_rhs_connections_and_weights). */
static void deep_compile_rhs(psi_t* p) {
    for(unsigned l = 0; l + 1 < p->num_layers; l++) {
        deep_layer_t* lo = &p->layers[l]; deep_layer_t* hi = &p->layers[l + 1];
        lo->rhs_conn = hi->size * hi->conn / lo->size;
        free(lo->rhs_connections); free(lo->rhs_weights);
        lo->rhs_connections = (unsigned*)calloc((size_t)lo->size * (lo->rhs_conn ? lo->rhs_conn : 1), sizeof(unsigned));
        lo->rhs_weights = (cx*)calloc((size_t)lo->size * (lo->rhs_conn ? lo->rhs_conn : 1), sizeof(cx));
        unsigned* fill = (unsigned*)calloc(lo->size, sizeof(unsigned));
        for(unsigned j = 0; j < hi->size; j++)
            for(unsigned i = 0; i < hi->conn; i++) {
                const unsigned lhs = hi->lhs_connections[i * hi->size + j];
                lo->rhs_connections[lhs * lo->rhs_conn + fill[lhs]] = j;
                lo->rhs_weights[lhs * lo->rhs_conn + fill[lhs]] = hi->lhs_weights[i * hi->size + j];
                fill[lhs]++;
            }
        free(fill);
    }
    p->layers[p->num_layers - 1].rhs_conn = 0;
}

/* Parameter layout: source/quantum_state/PsiDeep.cu:143-183 (init_kernel), :246-270 (get_params):
 * [input_weights (N)] then per hidden layer [biases (size)] [lhs_weights (conn x size)]. */
psi_t* port_deep_create(unsigned num_sites, unsigned N, const double* input_weights, unsigned num_hidden,
                        const unsigned* sizes, const unsigned* conn, const double* biases,
                        const unsigned* lhs_connections, const double* lhs_weights,
                        const double* final_weights, const double* lp) {
    psi_t* p = (psi_t*)calloc(1, sizeof(psi_t));
    p->kind = K_DEEP; p->N = N; p->num_sites = num_sites; p->words = words_for(N); p->num_layers = num_hidden + 1;
    p->log_prefactor = lp[0] + lp[1] * I;
    p->input_weights = (cx*)malloc(sizeof(cx) * N);
    for(unsigned i = 0; i < N; i++) p->input_weights[i] = input_weights[2 * i] + input_weights[2 * i + 1] * I;
    p->layers[0].size = N; p->width = N;
    p->num_params = N;
    size_t off_b = 0, off_w = 0; unsigned deep = 0;
    for(unsigned l = 1; l <= num_hidden; l++) {
        deep_layer_t* L = &p->layers[l];
        L->size = sizes[l - 1]; L->conn = conn[l - 1];
        if(L->size > p->width) p->width = L->size;
        const size_t nw = (size_t)L->size * L->conn;
        L->biases = (cx*)malloc(sizeof(cx) * L->size);
        L->lhs_weights = (cx*)malloc(sizeof(cx) * nw);
        L->lhs_connections = (unsigned*)malloc(sizeof(unsigned) * nw);
        for(unsigned j = 0; j < L->size; j++) L->biases[j] = biases[2 * (off_b + j)] + biases[2 * (off_b + j) + 1] * I;
        for(size_t k = 0; k < nw; k++) {
            L->lhs_weights[k] = lhs_weights[2 * (off_w + k)] + lhs_weights[2 * (off_w + k) + 1] * I;
            L->lhs_connections[k] = lhs_connections[off_w + k];
        }
        L->begin_params = p->num_params;
        p->num_params += L->size + (unsigned)nw;
        if(l > 1) { L->begin_deep = deep; deep += L->size; }
        off_b += L->size; off_w += nw;
    }
    p->num_deep = deep;
    const unsigned nf = p->layers[num_hidden].size;
    p->final_weights = (cx*)malloc(sizeof(cx) * nf);
    for(unsigned j = 0; j < nf; j++) p->final_weights[j] = final_weights[2 * j] + final_weights[2 * j + 1] * I;
    deep_compile_rhs(p);
    return p;
}

/* source/quantum_state/PsiCNN.cpp:10-55 (init_kernel): params flat, channel-link major. */
psi_t* port_cnn_create(const unsigned* extent, unsigned num_layers, const unsigned* num_channels,
                       const unsigned* connectivity, const unsigned* symmetry_classes,
                       const double* params, unsigned num_params, double final_factor, const double* lp) {
    psi_t* p = (psi_t*)calloc(1, sizeof(psi_t));
    p->kind = K_CNN; p->cnn_layers = num_layers; p->final_factor = final_factor;
    p->N = extent[0] * extent[1] * extent[2]; p->words = words_for(p->N);
    for(int d = 0; d < 3; d++) p->extent[d] = extent[d];
    p->log_prefactor = lp[0] + lp[1] * I; p->num_params = num_params;
    p->sym = (unsigned*)malloc(sizeof(unsigned) * p->N);
    unsigned maxsym = 0;
    for(unsigned i = 0; i < p->N; i++) { p->sym[i] = symmetry_classes[i]; if(p->sym[i] > maxsym) maxsym = p->sym[i]; }
    {   /* num_symmetry_classes = number of distinct values (PsiCNN.hpp:334-336) */
        unsigned char* seen = (unsigned char*)calloc(maxsym + 1, 1); unsigned cnt = 0;
        for(unsigned i = 0; i < p->N; i++) if(!seen[p->sym[i]]) { seen[p->sym[i]] = 1; cnt++; }
        free(seen); p->num_sym = cnt;
    }
    p->params = (cx*)malloc(sizeof(cx) * num_params);
    for(unsigned k = 0; k < num_params; k++) p->params[k] = params[2 * k] + params[2 * k + 1] * I;
    unsigned off = 0; p->num_angles = 0;
    for(unsigned l = 0; l < num_layers; l++) {
        cnn_layer_t* L = &p->cl[l];
        L->num_channels = num_channels[l];
        L->num_links = L->num_channels * (l > 0 ? p->cl[l - 1].num_channels : 1u);
        L->vol = 1;
        for(int d = 0; d < 3; d++) { L->connectivity[d] = connectivity[l * 3 + d]; L->vol *= L->connectivity[d]; }
        for(unsigned c = 0; c < L->num_links; c++) { L->links[c].begin_params = off; off += p->num_sym * L->vol; }
        L->angle_offset = p->num_angles;
        p->num_angles += L->num_channels * p->N;
    }
    return p;
}

/* include/quantum_state/PsiClassical.hpp:192-214 + source/quantum_state/PsiClassical.cu:12-34. */
psi_t* port_classical_create(unsigned num_sites, unsigned order, unsigned num_ops, op_t** ops,
                             const double* params, unsigned num_own, psi_t* ref, const double* lp) {
    psi_t* p = (psi_t*)calloc(1, sizeof(psi_t));
    p->kind = K_CLASSICAL; p->N = num_sites; p->words = words_for(num_sites); p->order = order;
    p->num_ops = num_ops; p->num_own = num_own; p->ref = ref;
    p->ops = (op_t**)malloc(sizeof(op_t*) * (num_ops ? num_ops : 1));
    for(unsigned i = 0; i < num_ops; i++) p->ops[i] = ops[i];
    p->cparams = (cx*)malloc(sizeof(cx) * (num_own ? num_own : 1));
    for(unsigned i = 0; i < num_own; i++) p->cparams[i] = params[2 * i] + params[2 * i + 1] * I;
    p->log_prefactor = lp[0] + lp[1] * I;
    p->num_params = num_own + ((order > 1u && ref) ? ref->num_params : 0u);
    return p;
}

void port_psi_destroy(psi_t* p) {
    free(p->W); free(p->input_weights); free(p->final_weights); free(p->sym); free(p->params);
    free(p->ops); free(p->cparams);
    for(unsigned l = 0; l < MAX_LAYERS; l++) {
        free(p->layers[l].lhs_connections); free(p->layers[l].rhs_connections);
        free(p->layers[l].lhs_weights); free(p->layers[l].rhs_weights); free(p->layers[l].biases);
    }
    free(p);
}
unsigned port_psi_num_params(const psi_t* p) { return p->num_params; }
void port_psi_set_log_prefactor(psi_t* p, double re, double im) { p->log_prefactor = re + im * I; }

void port_psi_get_params(const psi_t* p, double* out) {
    cx* o = (cx*)out;
    if(p->kind == K_RBM) memcpy(o, p->W, sizeof(cx) * p->num_params);
    else if(p->kind == K_CNN) memcpy(o, p->params, sizeof(cx) * p->num_params);
    else if(p->kind == K_DEEP) {
        memcpy(o, p->input_weights, sizeof(cx) * p->N); o += p->N;
        for(unsigned l = 1; l < p->num_layers; l++) {
            const deep_layer_t* L = &p->layers[l];
            memcpy(o, L->biases, sizeof(cx) * L->size); o += L->size;
            memcpy(o, L->lhs_weights, sizeof(cx) * L->size * L->conn); o += L->size * L->conn;
        }
    } else {
        memcpy(o, p->cparams, sizeof(cx) * p->num_own);
        if(p->order > 1u && p->ref) port_psi_get_params(p->ref, out + 2 * p->num_own);
    }
}
void port_psi_set_params(psi_t* p, const double* in) {
    const cx* s = (const cx*)in;
    if(p->kind == K_RBM) memcpy(p->W, s, sizeof(cx) * p->num_params);
    else if(p->kind == K_CNN) memcpy(p->params, s, sizeof(cx) * p->num_params);
    else if(p->kind == K_DEEP) {
        memcpy(p->input_weights, s, sizeof(cx) * p->N); s += p->N;
        for(unsigned l = 1; l < p->num_layers; l++) {
            deep_layer_t* L = &p->layers[l];
            memcpy(L->biases, s, sizeof(cx) * L->size); s += L->size;
            memcpy(L->lhs_weights, s, sizeof(cx) * L->size * L->conn); s += L->size * L->conn;
        }
        deep_compile_rhs(p);
    } else {
        memcpy(p->cparams, s, sizeof(cx) * p->num_own);
        if(p->order > 1u && p->ref) port_psi_set_params(p->ref, in + 2 * p->num_own);
    }
}

/* ------------------------------------------------------------------ payloads (cached angles)
 * RBM : include/quantum_state/PsiRBM.hpp:71-80 (compute_angles), :122-157 (update_input_units).
 * Deep: include/quantum_state/PsiDeep.hpp:140-157, :269-280, :311-343.
 * CNN : no cache (PsiCNN.hpp:177-181).  Classical: delegates to psi_ref for order 2. */
typedef struct payload_s { cx* angles; cx* act; cx* deep; cx* cnn_in; cx* cnn_out; cx* cnn_angles; struct payload_s* refp; } payload_t;

static payload_t* payload_new(const psi_t* p) {
    payload_t* pl = (payload_t*)calloc(1, sizeof(payload_t));
    if(p->kind == K_RBM) pl->angles = (cx*)malloc(sizeof(cx) * p->M);
    else if(p->kind == K_DEEP) {
        pl->angles = (cx*)malloc(sizeof(cx) * p->width);
        pl->act = (cx*)malloc(sizeof(cx) * p->width * 2);
        pl->deep = (cx*)malloc(sizeof(cx) * (p->num_deep ? p->num_deep : 1));
    } else if(p->kind == K_CNN) {
        unsigned maxc = 1;
        for(unsigned l = 0; l < p->cnn_layers; l++) if(p->cl[l].num_channels > maxc) maxc = p->cl[l].num_channels;
        pl->cnn_in = (cx*)malloc(sizeof(cx) * maxc * p->N);
        pl->cnn_out = (cx*)malloc(sizeof(cx) * maxc * p->N);
        pl->cnn_angles = (cx*)malloc(sizeof(cx) * p->num_angles);
    } else if(p->ref) pl->refp = payload_new(p->ref);
    return pl;
}
static void payload_free(payload_t* pl) {
    if(!pl) return;
    free(pl->angles); free(pl->act); free(pl->deep); free(pl->cnn_in); free(pl->cnn_out); free(pl->cnn_angles);
    payload_free(pl->refp); free(pl);
}

static void payload_init(const psi_t* p, payload_t* pl, const uint64_t* conf) {
    if(p->kind == K_RBM) {
        for(unsigned j = 0; j < p->M; j++) {
            cx a = 0.0;
            for(unsigned i = 0; i < p->N; i++) a += p->W[i * p->M + j] * spin_at(conf, i);
            pl->angles[j] = a;
        }
    } else if(p->kind == K_DEEP) {
        const deep_layer_t* L = &p->layers[1];
        for(unsigned j = 0; j < L->size; j++) {
            cx a = 0.0;
            for(unsigned i = 0; i < L->conn; i++) a += L->lhs_weights[i * L->size + j] * spin_at(conf, L->lhs_connections[i * L->size + j]);
            pl->angles[j] = a + L->biases[j];
        }
    } else if(p->kind == K_CLASSICAL && p->order > 1u && p->ref) payload_init(p->ref, pl->refp, conf);
}

static void payload_update(const psi_t* p, payload_t* pl, const uint64_t* old_conf, const uint64_t* new_conf) {
    if(p->kind == K_CLASSICAL) { if(p->order > 1u && p->ref) payload_update(p->ref, pl->refp, old_conf, new_conf); return; }
    if(p->kind == K_CNN) return;
    for(unsigned w = 0; w < p->words; w++) {
        uint64_t diff = old_conf[w] ^ new_conf[w];
        while(diff) {
            const unsigned pos = w * 64u + (unsigned)__builtin_ctzll(diff);
            const double delta = spin_at(new_conf, pos) - spin_at(old_conf, pos);
            if(p->kind == K_RBM) {
                for(unsigned j = 0; j < p->M; j++) pl->angles[j] += delta * p->W[pos * p->M + j];
            } else {
                const deep_layer_t* L0 = &p->layers[0];
                for(unsigned j = 0; j < L0->rhs_conn; j++)
                    pl->angles[L0->rhs_connections[pos * L0->rhs_conn + j]] += delta * L0->rhs_weights[pos * L0->rhs_conn + j];
            }
            diff &= diff - 1;
        }
    }
}

/* CNN periodic forward-shift connections, include/quantum_state/detail/Convolve.hpp:122-149. */
static inline unsigned cnn_input_idx(const psi_t* p, unsigned idx, unsigned k, unsigned i, unsigned j) {
    const unsigned page_size = p->extent[1] * p->extent[2];
    const unsigned page = idx / page_size, row = (idx % page_size) / p->extent[2], col = (idx % page_size) % p->extent[2];
    return ((page + k) % p->extent[0]) * page_size + ((row + i) % p->extent[1]) * p->extent[2] + ((col + j) % p->extent[2]);
}

/* PsiCNN forward_pass, include/quantum_state/PsiCNN.hpp:99-160. Returns sum(last activations)*final_factor. */
static cx cnn_forward(const psi_t* p, payload_t* pl, const uint64_t* conf) {
    const unsigned N = p->N;
    for(unsigned j = 0; j < N; j++) pl->cnn_in[j] = spin_at(conf, j);
    cx result = 0.0;
    for(unsigned l = 0; l < p->cnn_layers; l++) {
        const cnn_layer_t* L = &p->cl[l];
        const unsigned prev = l > 0 ? p->cl[l - 1].num_channels : 1u;
        for(unsigned cj = 0; cj < L->num_channels; cj++) {
            for(unsigned x = 0; x < N; x++) {
                cx acc = 0.0;
                for(unsigned ci = 0; ci < prev; ci++) {
                    const cx* w = p->params + L->links[ci * L->num_channels + cj].begin_params;
                    cx part = 0.0; unsigned c = 0;
                    for(unsigned k = 0; k < L->connectivity[0]; k++)
                        for(unsigned i = 0; i < L->connectivity[1]; i++)
                            for(unsigned j = 0; j < L->connectivity[2]; j++, c++)
                                part += w[p->sym[x] * L->vol + c] * pl->cnn_in[ci * N + cnn_input_idx(p, x, k, i, j)];
                    acc += part;
                }
                pl->cnn_angles[L->angle_offset + cj * N + x] = acc;
                pl->cnn_out[cj * N + x] = act_lc(acc, l);
            }
        }
        for(unsigned j = 0; j < L->num_channels * N; j++) {
            if(l + 1 < p->cnn_layers) pl->cnn_in[j] = pl->cnn_out[j];
            else result += pl->cnn_out[j] * p->final_factor;
        }
    }
    return result;
}

/* PsiDeep forward_pass, include/quantum_state/PsiDeep.hpp:173-215. */
static cx deep_forward(const psi_t* p, payload_t* pl) {
    cx* act = pl->act; cx* nxt = pl->act + p->width;
    for(unsigned i = 0; i < p->layers[1].size; i++) act[i] = act_lc(pl->angles[i], 0u);
    for(unsigned l = 2; l < p->num_layers; l++) {
        const deep_layer_t* L = &p->layers[l];
        for(unsigned j = 0; j < L->size; j++) {
            cx a = 0.0;
            for(unsigned i = 0; i < L->conn; i++) a += L->lhs_weights[i * L->size + j] * act[L->lhs_connections[i * L->size + j]];
            a += L->biases[j];
            pl->deep[L->begin_deep + j] = a;
            nxt[j] = act_lc(a, l - 1);
        }
        memcpy(act, nxt, sizeof(cx) * L->size);
    }
    cx r = 0.0;
    const unsigned nf = p->layers[p->num_layers - 1].size;
    for(unsigned j = 0; j < nf; j++) r += act[j] * p->final_weights[j];
    return r;
}

/* log_psi_s of each model: PsiRBM.hpp:110-119 (+ :92-106), PsiDeep.hpp:219-266, PsiCNN.hpp:164-175,
 * PsiClassical.hpp:85-112 (PsiFullyPolarized.hpp:41-49 gives 0 for the reference state). */
static cx psi_log_psi(const psi_t* p, payload_t* pl, const uint64_t* conf) {
    cx r = p->log_prefactor;
    if(p->kind == K_RBM) {
        for(unsigned j = 0; j < p->M; j++) r += act_lc(pl->angles[j], 0u) * p->final_weight;
    } else if(p->kind == K_DEEP) r += deep_forward(p, pl);
    else if(p->kind == K_CNN) r += cnn_forward(p, pl, conf);
    else {
        for(unsigned n = 0; n < p->num_ops; n++) r += p->cparams[n] * op_fast_local_energy(p->ops[n], conf);
        if(p->order > 1u && p->ref) r += psi_log_psi(p->ref, pl->refp, conf);
    }
    return r;
}

/* foreach_O_k of each model, accumulated into O (caller zeroes): PsiRBM.hpp:161-176,
 * PsiDeep.hpp:347-445, PsiCNN.hpp:185-266, PsiClassical.hpp:124-145. */
static void psi_O_k(const psi_t* p, payload_t* pl, const uint64_t* conf, cx* O) {
    if(p->kind == K_RBM) {
        for(unsigned j = 0; j < p->M; j++) {
            const cx a = p->final_weight * act_th(pl->angles[j], 0u);
            for(unsigned i = 0; i < p->N; i++) O[i * p->M + j] += a * spin_at(conf, i);
        }
    } else if(p->kind == K_DEEP) {
        for(unsigned i = 0; i < p->N; i++) O[i] += spin_at(conf, i);
        payload_init(p, pl, conf);
        (void)deep_forward(p, pl);
        cx* act = pl->act; cx* tmp = pl->act + p->width;
        for(int l = (int)p->num_layers - 1; l > 0; l--) {
            const deep_layer_t* L = &p->layers[l];
            if(l == (int)p->num_layers - 1) {
                for(unsigned j = 0; j < L->size; j++)
                    act[j] = p->final_weights[j] * (p->num_layers == 2u ? act_th(pl->angles[j], 0u)
                                                                        : act_th(pl->deep[L->begin_deep + j], p->num_layers - 2u));
            } else {
                for(unsigned i = 0; i < L->size; i++) {
                    cx u = 0.0;
                    for(unsigned j = 0; j < L->rhs_conn; j++) u += L->rhs_weights[i * L->rhs_conn + j] * act[L->rhs_connections[i * L->rhs_conn + j]];
                    u *= (l == 1 ? act_th(pl->angles[i], 0u) : act_th(pl->deep[L->begin_deep + i], (unsigned)l - 1u));
                    tmp[i] = u;
                }
                memcpy(act, tmp, sizeof(cx) * L->size);
            }
            for(unsigned j = 0; j < L->size; j++) {
                O[L->begin_params + j] += act[j];
                for(unsigned i = 0; i < L->conn; i++) {
                    const unsigned lhs = L->lhs_connections[i * L->size + j];
                    const cx in = (l == 1) ? (cx)spin_at(conf, lhs)
                                : (l == 2) ? act_lc(pl->angles[lhs], 0u)
                                           : act_lc(pl->deep[p->layers[l - 1].begin_deep + lhs], (unsigned)l - 1u);
                    O[L->begin_params + L->size + i * L->size + j] += act[j] * in;
                }
            }
        }
    } else if(p->kind == K_CNN) {
        const unsigned N = p->N;
        (void)cnn_forward(p, pl, conf);
        cx* in_act = pl->cnn_in; cx* out_act = pl->cnn_out;
        const unsigned lastc = p->cl[p->cnn_layers - 1].num_channels;
        for(unsigned j = 0; j < lastc * N; j++) out_act[j] = p->final_factor;
        for(int l = (int)p->cnn_layers - 1; l >= 0; l--) {
            const cnn_layer_t* L = &p->cl[l];
            const unsigned prev = l > 0 ? p->cl[l - 1].num_channels : 1u;
            for(unsigned cj = 0; cj < L->num_channels; cj++)
                for(unsigned j = 0; j < N; j++)
                    in_act[cj * N + j] = out_act[cj * N + j] * act_th(pl->cnn_angles[L->angle_offset + cj * N + j], (unsigned)l);
            for(unsigned ci = 0; ci < prev; ci++) {
                for(unsigned i = 0; i < N; i++) out_act[ci * N + i] = 0.0;
                for(unsigned cj = 0; cj < L->num_channels; cj++) {
                    const unsigned bp = L->links[ci * L->num_channels + cj].begin_params;
                    const cx* w = p->params + bp;
                    for(unsigned x = 0; x < N; x++) {
                        unsigned c = 0;
                        for(unsigned k = 0; k < L->connectivity[0]; k++)
                            for(unsigned i = 0; i < L->connectivity[1]; i++)
                                for(unsigned j = 0; j < L->connectivity[2]; j++, c++) {
                                    const unsigned src = cnn_input_idx(p, x, k, i, j);
                                    const cx in = (l == 0) ? (cx)spin_at(conf, src)
                                                           : act_lc(pl->cnn_angles[p->cl[l - 1].angle_offset + ci * N + src], (unsigned)l - 1u);
                                    O[bp + p->sym[x] * L->vol + c] += in_act[cj * N + x] * in;
                                    if(l > 0) out_act[ci * N + src] += w[p->sym[x] * L->vol + c] * in_act[cj * N + x];
                                }
                    }
                }
            }
        }
    } else {
        for(unsigned n = 0; n < p->num_ops; n++) O[n] += op_fast_local_energy(p->ops[n], conf);
        if(p->order > 1u && p->ref) { payload_init(p->ref, pl->refp, conf); psi_O_k(p->ref, pl->refp, conf, O + p->num_ops); }
    }
}

/* StandartOperator::local_energy / nth_local_energy, include/operator/Operator.hpp:38-121:
 * strictly serial over strings; off-diagonal strings go through update -> log psi' -> restore. */
static cx op_local_energy(const op_t* op, const psi_t* p, payload_t* pl, const uint64_t* conf, cx log_psi) {
    cx result = 0.0; uint64_t prime[MAXW];
    for(unsigned n = 0; n < op->n; n++) {
        const cx me = pauli_apply(op->a + n * op->words, op->b + n * op->words, conf, op->words, prime) * op->coef[n];
        if(!conf_equal(conf, prime, op->words)) {
            payload_update(p, pl, conf, prime);
            const cx lp = psi_log_psi(p, pl, prime);
            result += me * cexp(lp - log_psi);
            payload_update(p, pl, prime, conf);
        } else result += me;
    }
    return result;
}

/* The same on the Pauli-string basis: `units` is the network-side mask of the Pauli string x; a string P of the operator maps it
 * to factor * (P xor x) (pauli_mul); only the identity string is diagonal (PauliString::is_diagonal_on_basis(PauliString), :163-165). */
static cx op_local_energy_paulis(const op_t* op, const psi_t* p, payload_t* pl, const uint64_t* units, cx log_psi) {
    cx result = 0.0;
    uint64_t xa[MAXW], xb[MAXW], na[MAXW], nb[MAXW], prime[MAXW];
    port_units_to_paulis(units, p->num_sites, xa, xb);
    for(unsigned n = 0; n < op->n; n++) {
        const cx me = pauli_mul(op->a + n * op->words, op->b + n * op->words, xa, xb, op->words, na, nb) * op->coef[n];
        port_paulis_to_units(na, nb, p->num_sites, prime);
        if(!conf_equal(units, prime, p->words)) {
            payload_update(p, pl, units, prime);
            const cx lp = psi_log_psi(p, pl, prime);
            result += me * cexp(lp - log_psi);
            payload_update(p, pl, prime, units);
        } else result += me;
    }
    return result;
}
static inline int psi_on_paulis(const psi_t* p) { return p->kind == K_DEEP && p->num_sites && p->N == 3u * p->num_sites; }
static inline cx local_energy_any(const op_t* op, const psi_t* p, payload_t* pl, const uint64_t* conf, cx log_psi) {
    return psi_on_paulis(p) ? op_local_energy_paulis(op, p, pl, conf, log_psi) : op_local_energy(op, p, pl, conf, log_psi);
}

/* ------------------------------------------------------------------ single-configuration probes
 * source/network_functions/PsiVector.cu.template:102-133, PsiOkVector.cu.template:41-74. */
void port_log_psi_s(const psi_t* p, const uint64_t* conf, double* out) {
    payload_t* pl = payload_new(p); payload_init(p, pl, conf);
    const cx r = psi_log_psi(p, pl, conf); out[0] = creal(r); out[1] = cimag(r);
    payload_free(pl);
}
void port_psi_O_k(const psi_t* p, const uint64_t* conf, double* out) {
    payload_t* pl = payload_new(p); payload_init(p, pl, conf);
    memset(out, 0, sizeof(cx) * p->num_params);
    psi_O_k(p, pl, conf, (cx*)out);
    payload_free(pl);
}
void port_local_energy(const psi_t* p, const op_t* op, const uint64_t* conf, double* out) {
    payload_t* pl = payload_new(p); payload_init(p, pl, conf);
    const cx lp = psi_log_psi(p, pl, conf);
    const cx e = local_energy_any(op, p, pl, conf, lp); out[0] = creal(e); out[1] = cimag(e);
    payload_free(pl);
}

/* ------------------------------------------------------------------ batch evaluation on given configurations
 * The per-sample body shared by every consumer lambda of the reference (ExpectationValue.cu.template:236-264,
 * TDVP.cu.template:91-125): log psi, E_loc, then init_payload + foreach_O_k.  Reductions over samples are
 * done by the caller (oracle/vmc_oracle.py) with the reference's formulas.
 * confs: ns x words.  Any of log_psi_out / eloc_out / O_out (ns x P, row-major) may be NULL. */
void port_eval_samples(const psi_t* p, const op_t* op, const uint64_t* confs, unsigned long ns,
                       double* log_psi_out, double* eloc_out, double* O_out, int nthreads) {
#ifdef _OPENMP
    if(nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel
    {
        payload_t* pl = payload_new(p);
        #pragma omp for schedule(dynamic, 16)
        for(unsigned long s = 0; s < ns; s++) {
            const uint64_t* conf = confs + s * p->words;
            payload_init(p, pl, conf);
            const cx lp = psi_log_psi(p, pl, conf);
            if(log_psi_out) { log_psi_out[2 * s] = creal(lp); log_psi_out[2 * s + 1] = cimag(lp); }
            if(eloc_out && op) {
                const cx e = local_energy_any(op, p, pl, conf, lp);
                eloc_out[2 * s] = creal(e); eloc_out[2 * s + 1] = cimag(e);
            }
            if(O_out) {
                cx* row = (cx*)O_out + s * (unsigned long)p->num_params;
                memset(row, 0, sizeof(cx) * p->num_params);
                payload_init(p, pl, conf);
                psi_O_k(p, pl, conf, row);
            }
        }
        payload_free(pl);
    }
}

/* ------------------------------------------------------------------ Philox4x32-10 (Salmon et al., SC'11)
 * Counter layout shared with the CUDA path (annongpu_b200/csrc/philox.cuh):
 *   ctr = (step_lo, step_hi, chain, (call << 1) | tag), key = (seed_lo, seed_hi);
 *   tag 0: initial configuration (step = word index); tag 1: Metropolis proposals. */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for(int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void port_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

/* ------------------------------------------------------------------ Monte-Carlo sampling
 * MonteCarlo_t::kernel_foreach / mc_update, include/ensembles/MonteCarlo.hpp:57-177;
 * Init_Policy (policies/Init_Policy.hpp:16-24), Update_Policy (policies/Update_Policy.hpp:20-28).
 * Differences from the reference, by design: every chain runs (the reference's
 * CPU path runs chain 0 only), and the random stream is Philox (see above).
 * Sample index = step * num_chains + chain (MonteCarlo.hpp:107-108); num_mc_steps_per_chain =
 * num_samples / num_chains (source/ensembles/MonteCarlo.cu:35).
 * chain0: global id of this call's first chain (multi-GPU sharding uses global ids). */
typedef struct {
    unsigned long num_samples; unsigned num_sweeps, num_therm; unsigned long num_chains;
    uint64_t seed; uint32_t call; unsigned long chain0;
} mc_params_t;

static void mc_chain(const psi_t* p, const mc_params_t* mc, unsigned long chain, unsigned long steps_per_chain,
                     uint64_t* confs_out, double* log_psi_out, unsigned long* acc, unsigned long* rej) {
    payload_t* pl = payload_new(p);
    uint64_t conf[MAXW] = {0}, next[MAXW];
    const uint32_t k0 = (uint32_t)mc->seed, k1 = (uint32_t)(mc->seed >> 32);
    const uint32_t gchain = (uint32_t)(mc->chain0 + chain);
    uint32_t r[4];
    const int paulis = psi_on_paulis(p);
    if(paulis) {
        /* Init_Policy<PauliString> = PauliString::set_randomly (policies/Init_Policy.hpp:32-42, PauliString.hpp:59-65): two random
         * masks, both cut with (1 << (num_sites % 64)) - 1 -- which is 0 for num_sites = 64 (kept) */
        uint64_t a[MAXW] = {0}, b[MAXW] = {0};
        const unsigned sw = words_for_c(p->num_sites);
        for(unsigned w = 0; w < sw; w++) {
            philox4x32_10(w, 0u, gchain, (mc->call << 1) | 0u, k0, k1, r);
            a[w] = (uint64_t)r[0] | ((uint64_t)r[1] << 32); b[w] = (uint64_t)r[2] | ((uint64_t)r[3] << 32);
        }
        const uint64_t cut = (1ull << (p->num_sites % 64u)) - 1ull;
        a[sw - 1] &= cut; b[sw - 1] &= cut;
        port_paulis_to_units(a, b, p->num_sites, conf);
    } else {
    for(unsigned w = 0; w < p->words; w++) {
        philox4x32_10(w, 0u, gchain, (mc->call << 1) | 0u, k0, k1, r);
        conf[w] = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
    }
    if(p->N % 64u) conf[p->words - 1] &= (1ull << (p->N % 64u)) - 1ull;
    }
    payload_init(p, pl, conf);
    cx log_psi = psi_log_psi(p, pl, conf);
    uint64_t t = 0;
    const unsigned long therm = (unsigned long)mc->num_therm * p->N, per_sample = (unsigned long)mc->num_sweeps * p->N;
    for(unsigned long s = 0; s <= steps_per_chain; s++) {
        /* s == 0: thermalisation; s >= 1: the sweeps before recorded sample s-1 */
        const unsigned long nsteps = (s == 0) ? therm : per_sample;
        for(unsigned long i = 0; i < nsteps; i++, t++) {
            philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), gchain, (mc->call << 1) | 1u, k0, k1, r);
            memcpy(next, conf, sizeof(conf));
            if(paulis) {
                /* Update_Policy<PauliString> (policies/Update_Policy.hpp:38-55): one random number x, site x % num_sites takes type x >> 30 */
                const unsigned site = r[0] % p->num_sites, type = r[0] >> 30;
                for(unsigned k = 0; k < 3u; k++) { const unsigned u = 3u * site + k; next[u >> 6] &= ~(1ull << (u & 63u)); }
                if(type) { const unsigned u = 3u * site + type - 1u; next[u >> 6] |= 1ull << (u & 63u); }
            } else {
                const unsigned site = r[0] % p->N;
                next[site >> 6] ^= 1ull << (site & 63u);
            }
            payload_update(p, pl, conf, next);
            const cx nlp = psi_log_psi(p, pl, next);
            const double ratio = exp(2.0 * (creal(nlp) - creal(log_psi)));
            const double u = (double)((((uint64_t)r[1] | ((uint64_t)r[2] << 32)) >> 11) + 1ull) * 0x1.0p-53;
            if(ratio > 1.0 || u <= ratio) { log_psi = nlp; memcpy(conf, next, sizeof(conf)); (*acc)++; }
            else { payload_update(p, pl, next, conf); (*rej)++; }
        }
        if(s == 0) continue;
        const unsigned long idx = (s - 1) * mc->num_chains + chain;
        memcpy(confs_out + idx * p->words, conf, sizeof(uint64_t) * p->words);
        if(log_psi_out) {
            /* the reference hands the consumer the running log_psi (MonteCarlo.hpp:107-113) */
            log_psi_out[2 * idx] = creal(log_psi); log_psi_out[2 * idx + 1] = cimag(log_psi);
        }
    }
    payload_free(pl);
}

void port_mc_sample(const psi_t* p, unsigned long num_samples, unsigned num_sweeps, unsigned num_therm,
                    unsigned long num_chains, uint64_t seed, uint32_t call, unsigned long chain0,
                    uint64_t* confs_out, double* log_psi_out, unsigned long* acc_rej_out, int nthreads) {
    mc_params_t mc = {num_samples, num_sweeps, num_therm, num_chains, seed, call, chain0};
    const unsigned long steps_per_chain = num_samples / num_chains;
    unsigned long acc = 0, rej = 0;
#ifdef _OPENMP
    if(nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel for schedule(dynamic, 1) reduction(+:acc, rej)
    for(unsigned long c = 0; c < num_chains; c++) {
        unsigned long a = 0, r = 0;
        mc_chain(p, &mc, c, steps_per_chain, confs_out, log_psi_out, &a, &r);
        acc += a; rej += r;
    }
    acc_rej_out[0] = acc; acc_rej_out[1] = rej;
}

/* ------------------------------------------------------------------ timed CPU baseline
 * One call of ExpectationValue::gradient over a Monte-Carlo ensemble
 * (source/network_functions/ExpectationValue.cu.template:220-275 over MonteCarlo.hpp:57-131):
 * sampling + E_loc + O_k with <O_k*>, <O_k* E_loc> accumulation, threaded over chains.
 * out_grad: P complex; out_E: 1 complex. */
void port_mc_gradient(const psi_t* p, const op_t* op, unsigned long num_samples, unsigned num_sweeps, unsigned num_therm,
                      unsigned long num_chains, uint64_t seed, uint32_t call,
                      double* out_grad, double* out_E, unsigned long* acc_rej_out, int nthreads) {
    const unsigned long P = p->num_params, ns = (num_samples / num_chains) * num_chains;
    uint64_t* confs = (uint64_t*)malloc(sizeof(uint64_t) * p->words * (ns ? ns : 1));
    port_mc_sample(p, num_samples, num_sweeps, num_therm, num_chains, seed, call, 0, confs, NULL, acc_rej_out, nthreads);
    const double weight = 1.0 / (double)num_samples;
    cx* grad = (cx*)out_grad; memset(grad, 0, sizeof(cx) * P);
    cx* Ok_mean = (cx*)calloc(P, sizeof(cx));
    cx E = 0.0;
    #pragma omp parallel
    {
        payload_t* pl = payload_new(p);
        cx* row = (cx*)malloc(sizeof(cx) * P);
        cx* g_loc = (cx*)calloc(P, sizeof(cx)); cx* o_loc = (cx*)calloc(P, sizeof(cx)); cx e_loc_sum = 0.0;
        #pragma omp for schedule(dynamic, 4)
        for(unsigned long s = 0; s < ns; s++) {
            const uint64_t* conf = confs + s * p->words;
            payload_init(p, pl, conf);
            const cx lp = psi_log_psi(p, pl, conf);
            const cx e = local_energy_any(op, p, pl, conf, lp);
            e_loc_sum += weight * e;
            memset(row, 0, sizeof(cx) * P);
            payload_init(p, pl, conf);
            psi_O_k(p, pl, conf, row);
            for(unsigned long k = 0; k < P; k++) { o_loc[k] += weight * conj(row[k]); g_loc[k] += weight * conj(row[k]) * e; }
        }
        #pragma omp critical
        { E += e_loc_sum; for(unsigned long k = 0; k < P; k++) { grad[k] += g_loc[k]; Ok_mean[k] += o_loc[k]; } }
        free(row); free(g_loc); free(o_loc); payload_free(pl);
    }
    for(unsigned long k = 0; k < P; k++) grad[k] -= E * Ok_mean[k];
    out_E[0] = creal(E); out_E[1] = cimag(E);
    free(Ok_mean); free(confs);
}

int port_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Host-side wavefunction objects: own the parameters (host copy + device buffers), hand by-value device
// views to the kernels, and launch the per-model kernels.  Mirrors the reference's host classes
// PsiRBM / PsiDeep / PsiCNN / PsiClassical (include/quantum_state/*.hpp) behind one virtual interface instead
// of the reference's template-instantiation matrix (template_engine/config.json).
#pragma once
#include "runtime.hpp"
#include "operator.hpp"
#include "psi_dev.cuh"
#include "kernels.cuh"
#include <memory>

namespace angpu {

// Device-resident batch of configurations with everything the consumers need per sample.
struct SampleSet {
    size_t   ns = 0;
    unsigned words = 1;
    DevBuf<uint64_t> conf;      // [ns][words]
    DevBuf<cplx>     log_psi;   // [ns]
    DevBuf<double>   weight;    // [ns]
    DevBuf<cplx>     eloc;      // [ns]
    DevBuf<cplx>     angles;    // [ns][A]  cached first-layer angles (PsiRBM fast path), valid iff has_angles
    bool     has_angles = false;
    unsigned pauli_sites = 0;   // != 0: the configurations are Pauli strings of that many sites, stored as units masks (pauli_basis.cuh)
    void resize(size_t ns_, unsigned words_) {
        ns = ns_; words = words_;
        conf.resize(ns * words); log_psi.resize(ns); weight.resize(ns); eloc.resize(ns);
        has_angles = false;
    }
};

struct Psi {
  private:
    static bool registry(const Psi* p, int op);     // +1 insert, -1 erase, 0 query (psi.cu)
  public:
    enum Kind { RBM = 0, DEEP = 1, CNN = 2, CLASSICAL = 3 };
    Kind     kind;
    unsigned N = 0, words = 1, P = 0;
    unsigned pauli_sites = 0;   // != 0: a network on the Pauli-string basis (PsiDeep with N = 3 num_sites input units)
    cplx     lp{0.0, 0.0};

    // live-object registry: TDVP / HilbertSpaceDistance / KullbackLeibler keep a Psi* for the lazily materialised dense O rows;
    // a C-ABI caller may destroy the psi in between, which must surface as an error instead of a use-after-free
    Psi() { registry(this, +1); }
    Psi(const Psi& o) : kind(o.kind), N(o.N), words(o.words), P(o.P), pauli_sites(o.pauli_sites), lp(o.lp) { registry(this, +1); }
    virtual ~Psi() { registry(this, -1); }
    static bool is_live(const Psi* p) { return registry(p, 0); }
    virtual Psi* clone() const = 0;
    virtual void get_params(cplx* out) const = 0;
    virtual void set_params(const cplx* in) = 0;
    virtual void set_log_prefactor(cplx v) { lp = v; }
    // params += alpha * x with x resident on the device (SR / TDVP update); the default goes through the host copy
    virtual void add_params_dev(const cplx* x_dev, cplx alpha);

    // fills S.log_psi (and S.weight = exp(2 Re log psi) when es_weights) for S.conf
    virtual void log_psi(SampleSet& S, bool es_weights) = 0;
    // fills S.eloc for S.conf / S.log_psi
    virtual void eloc(const Operator& op, SampleSet& S) = 0;
    // dense rows O[s0 .. s0+cnt) -> out (cnt x P)
    virtual void ok_rows(SampleSet& S, size_t s0, size_t cnt, cplx* out) = 0;
    // Metropolis sampling into S (conf, log_psi); S must be sized steps_per_chain * num_chains_local
    virtual void mc_sample(const McParams& mc, SampleSet& S, unsigned long long* acc_rej_dev) = 0;
};

// unit pairs per lane of the warp-per-chain RBM samplers (rows padded to 64 KK hidden units)
inline unsigned rbm_sampler_KK(unsigned M) { return M <= 64u ? 1u : M <= 128u ? 2u : M <= 256u ? 4u : 8u; }

struct PsiRBM : Psi {
    unsigned M = 0;
    cplx fw{1.0, 0.0};
    mutable std::vector<cplx> hW;
    DevBuf<cplx> dW, dWpad;
    DevBuf<cplx> T_scratch;             // factorised rows for ok_rows (dense O on request)
    DevBuf<float4> dWf;                 // fp32 copy of W in the screened sampler's layout (rbm_sampler.cuh), M <= 512
    unsigned Mpad = 0;
    const cplx* Wpad() const { return Mpad == M ? dW.p : dWpad.p; }

    PsiRBM(unsigned N_, unsigned M_, const cplx* W, cplx fw_, cplx lp_);
    RbmDev dev() const { return RbmDev{N, M, words, P, lp, fw, dW.p, (float)(2.0 * fw.re)}; }
    void upload(const cplx* src = nullptr);      // src: the caller's buffer (copied to the device straight from there), else hW
    // the device copies are updated in place by add_params_dev; the host copy is refreshed lazily
    mutable bool host_stale = false;
    void sync_host() const;
    Psi* clone() const override { sync_host(); return new PsiRBM(N, M, hW.data(), fw, lp); }
    void get_params(cplx* out) const override { sync_host(); std::memcpy(out, hW.data(), sizeof(cplx) * P); }
    void set_params(const cplx* in) override { upload(in); }
    void add_params_dev(const cplx* x_dev, cplx alpha) override;
    void log_psi(SampleSet& S, bool es_weights) override;
    void eloc(const Operator& op, SampleSet& S) override;
    void ok_rows(SampleSet& S, size_t s0, size_t cnt, cplx* out) override;
    void mc_sample(const McParams& mc, SampleSet& S, unsigned long long* acc_rej_dev) override;
    // factorised log-derivative T[s][j] = fw * th0(theta_sj) for the samples of S
    void compute_T(SampleSet& S, DevBuf<cplx>& T);
    void ensure_angles(SampleSet& S);
};

struct PsiDeep : Psi {
    struct Layer {
        unsigned size = 0, conn = 0, rhs_conn = 0, begin_params = 0, begin_deep = 0;
        std::vector<unsigned> lhs_c, rhs_c;
        std::vector<cplx> lhs_w, rhs_w, bias;
        DevBuf<unsigned> d_lhs_c, d_rhs_c;
        DevBuf<cplx> d_lhs_w, d_rhs_w, d_bias;
    };
    unsigned num_sites = 0, num_layers = 0, width = 0, num_deep = 0;
    std::vector<Layer> layers;          // [0] = input layer
    std::vector<cplx> input_weights, final_weights;
    DevBuf<cplx> d_final;
    // block-per-chain sampler (deep_kernels.cuh): eligible with 2 or 3 hidden layers, each <= 64 wide, no unit reading
    // the same input twice; d_w1dense = the layers as dense tables (first layer [N][64], then [64][64] per deep layer)
    bool block_sampler_ok = false;
    DevBuf<cplx> d_w1dense;

    PsiDeep(unsigned num_sites_, unsigned N_, const cplx* input_weights_, unsigned num_hidden, const unsigned* sizes,
            const unsigned* conn, const cplx* biases, const unsigned* lhs_connections, const cplx* lhs_weights,
            const cplx* final_weights_, cplx lp_);
    PsiDeep(const PsiDeep& o);
    void compile_rhs();
    void upload();
    DeepDev dev() const;
    Psi* clone() const override { return new PsiDeep(*this); }
    void get_params(cplx* out) const override;
    void set_params(const cplx* in) override;
    void log_psi(SampleSet& S, bool es_weights) override;
    void eloc(const Operator& op, SampleSet& S) override;
    void ok_rows(SampleSet& S, size_t s0, size_t cnt, cplx* out) override;
    void mc_sample(const McParams& mc, SampleSet& S, unsigned long long* acc_rej_dev) override;
};

struct PsiCNN : Psi {
    unsigned extent[3] = {1, 1, 1}, num_layers = 0, num_sym = 1, num_angles = 0, maxch = 1;
    double final_factor = 1.0;
    std::vector<unsigned> num_channels, connectivity, sym;
    std::vector<cplx> params;
    std::vector<std::vector<unsigned>> h_nbr, h_inv;
    DevBuf<unsigned> d_sym; DevBuf<cplx> d_params;
    std::vector<DevBuf<unsigned>> d_nbr, d_inv;
    CnnLayerDev layer_dev[CNN_MAX_LAYERS];
    // incremental sampler (cnn_kernels.cuh): per layer and flipped site, the output sites inside the receptive cone
    std::vector<DevBuf<unsigned>> d_aff, d_aff_cnt;
    unsigned aff_max[CNN_MAX_LAYERS] = {0, 0, 0, 0};
    std::vector<std::vector<std::vector<unsigned>>> h_cone;     // [layer][site] -> sorted affected output sites
    // per-operator union cones of the flip groups (cone-based E_loc), cached for the last operator seen
    std::vector<DevBuf<unsigned>> d_gaff, d_gaff_cnt;
    unsigned gaff_max[CNN_MAX_LAYERS] = {0, 0, 0, 0};
    uint64_t gaff_hash = 0; unsigned gaff_groups = 0;

    PsiCNN(const unsigned* extent_, unsigned num_layers_, const unsigned* num_channels_, const unsigned* connectivity_,
           const unsigned* symmetry_classes, const cplx* params_, unsigned num_params, double final_factor_, cplx lp_);
    void build();
    CnnDev dev(bool keep_angles = true) const;
    Psi* clone() const override {
        return new PsiCNN(extent, num_layers, num_channels.data(), connectivity.data(), sym.data(), params.data(), P, final_factor, lp);
    }
    void get_params(cplx* out) const override { std::memcpy(out, params.data(), sizeof(cplx) * P); }
    void set_params(const cplx* in) override { params.assign(in, in + P); d_params.upload(params); }
    void log_psi(SampleSet& S, bool es_weights) override;
    void eloc(const Operator& op, SampleSet& S) override;
    void ok_rows(SampleSet& S, size_t s0, size_t cnt, cplx* out) override;
    void mc_sample(const McParams& mc, SampleSet& S, unsigned long long* acc_rej_dev) override;
};

struct PsiClassical : Psi {
    unsigned order = 1, num_ops = 0;
    std::vector<std::unique_ptr<Operator>> ops;
    std::vector<cplx> own_params;
    std::unique_ptr<PsiCNN> ref;        // null => PsiFullyPolarized
    DevBuf<OpDev> d_ops; DevBuf<cplx> d_params;
    // the state eval_with_psi_ref samples from (kernel().psi_ref in the reference): the CNN, or for the FP variants a
    // parameter-free PsiClassical with log psi = 0 (PsiFullyPolarized.hpp:41-49), created on first use
    std::unique_ptr<PsiClassical> polarized;
    Psi* sampling_ref() {
        if(ref) return ref.get();
        if(!polarized) polarized.reset(new PsiClassical(N, 1u, 0u, nullptr, nullptr, 0u, nullptr, cplx(0.0, 0.0)));
        return polarized.get();
    }

    PsiClassical(unsigned num_sites, unsigned order_, unsigned num_ops_, const Operator* const* ops_, const cplx* params_,
                 unsigned num_own, const PsiCNN* ref_, cplx lp_);
    void upload();
    ClassicalDev dev() const;
    Psi* clone() const override;
    void get_params(cplx* out) const override;
    void set_params(const cplx* in) override;
    void log_psi(SampleSet& S, bool es_weights) override;
    void eloc(const Operator& op, SampleSet& S) override;
    void ok_rows(SampleSet& S, size_t s0, size_t cnt, cplx* out) override;
    void mc_sample(const McParams& mc, SampleSet& S, unsigned long long* acc_rej_dev) override;
};

} // namespace angpu

// Host-side wavefunction objects: parameter bookkeeping + kernel launches.
#include <unordered_set>
#include <mutex>
#include "psi.hpp"
#include "rbm_kernels.cuh"
#include "rbm_sampler.cuh"
#include "deep_kernels.cuh"
#include "cnn_kernels.cuh"
#include "pauli_basis.cuh"
#include <algorithm>
#include <set>
#include <string>

namespace angpu {

bool Psi::registry(const Psi* p, int op) {
    static std::mutex mu;
    static std::unordered_set<const Psi*> live;
    std::lock_guard<std::mutex> g(mu);
    if(op > 0) { live.insert(p); return true; }
    if(op < 0) { live.erase(p); return false; }
    return live.count(p) != 0;
}


static Ctx g_ctx;
Ctx& ctx() { return g_ctx; }

void ctx_init(int device) {
    if(g_ctx.device == device && g_ctx.stream) return;
    ANGPU_CUDA(cudaSetDevice(device));
    if(g_ctx.own_stream && g_ctx.stream) cudaStreamDestroy(g_ctx.stream);
    ANGPU_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
    g_ctx.own_stream = true;
    g_ctx.device = device;
    cudaDeviceProp prop;
    ANGPU_CUDA(cudaGetDeviceProperties(&prop, device));
    g_ctx.num_sms = prop.multiProcessorCount;
    g_ctx.smem_optin = prop.sharedMemPerBlockOptin;
}

// ---------------------------------------------------------------------------------------- generic launches

namespace {

struct WarpCfg { unsigned wpb, grid; size_t smem; };

// one warp per item; warps per block limited by the per-warp scratch slice (+ the model's block-level scratch)
WarpCfg warp_cfg(unsigned pl_elems, size_t items, size_t block_bytes = 0) {
    const size_t slice = warp_slice_bytes(pl_elems);
    const size_t budget = std::min<size_t>(ctx().smem_optin, 200 * 1024) - block_bytes;
    ANGPU_REQUIRE(slice <= budget, "model scratch does not fit in shared memory");
    unsigned wpb = (unsigned)std::min<size_t>(8, budget / slice);
    // keep >= 2 blocks per SM resident when the slice allows it
    while(wpb > 1 && wpb * slice > budget / 2) wpb--;
    const size_t blocks_needed = (items + wpb - 1) / wpb;
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>(blocks_needed, (size_t)ctx().num_sms * 16));
    return WarpCfg{wpb, grid, wpb * slice + block_bytes};
}

template<class F>
void set_smem(F* kernel, size_t smem) {
    if(smem > 48 * 1024) ANGPU_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
}

template<class Dev>
void generic_log_psi(const Dev& d, SampleSet& S, bool es_weights) {
    if(S.ns == 0) return;
    const WarpCfg c = warp_cfg(d.payload_elems(), S.ns, d.block_scratch_bytes());
    set_smem(k_log_psi<Dev>, c.smem);
    k_log_psi<Dev><<<c.grid, c.wpb * 32, c.smem, stream()>>>(d, S.conf.p, S.ns, S.log_psi.p, es_weights ? S.weight.p : nullptr);
    ANGPU_CHECK_LAUNCH(); count_launch();
}
template<class Dev>
void generic_eloc(const Dev& d, const Operator& op, SampleSet& S) {
    if(S.ns == 0) return;
    ANGPU_REQUIRE(op.words == d.words, "operator / wavefunction word count mismatch");
    const WarpCfg c = warp_cfg(d.payload_elems(), S.ns, d.block_scratch_bytes());
    set_smem(k_eloc<Dev>, c.smem);
    k_eloc<Dev><<<c.grid, c.wpb * 32, c.smem, stream()>>>(d, op.dev, S.conf.p, S.log_psi.p, S.ns, S.eloc.p);
    ANGPU_CHECK_LAUNCH(); count_launch();
}
template<class Dev>
void generic_ok(const Dev& d, SampleSet& S, size_t s0, size_t cnt, cplx* out) {
    if(cnt == 0) return;
    const WarpCfg c = warp_cfg(d.payload_elems(), cnt, d.block_scratch_bytes());
    set_smem(k_ok<Dev>, c.smem);
    k_ok<Dev><<<c.grid, c.wpb * 32, c.smem, stream()>>>(d, S.conf.p + s0 * d.words, cnt, out);
    ANGPU_CHECK_LAUNCH(); count_launch();
}
template<class Dev>
void generic_mc(const Dev& d, const McParams& mc, SampleSet& S, unsigned long long* acc_rej_dev) {
    if(mc.num_chains_local == 0) return;
    const size_t slice = warp_slice_bytes(d.payload_elems());
    const size_t bb = d.block_scratch_bytes();
    const size_t budget = std::min<size_t>(ctx().smem_optin, 200 * 1024) - bb;
    ANGPU_REQUIRE(slice <= budget, "model scratch does not fit in shared memory");
    unsigned wpb = (unsigned)std::min<size_t>(4, budget / slice);
    while(wpb > 1 && wpb * slice > budget / 2) wpb--;
    const unsigned grid = ceil_div(mc.num_chains_local, wpb);
    set_smem(k_mc<Dev>, wpb * slice + bb);
    k_mc<Dev><<<grid, wpb * 32, wpb * slice + bb, stream()>>>(d, mc, S.conf.p, S.log_psi.p, acc_rej_dev);
    ANGPU_CHECK_LAUNCH(); count_launch();
}

} // namespace

void Psi::add_params_dev(const cplx* x_dev, cplx alpha) {
    std::vector<cplx> x(P), p(P);
    ANGPU_CUDA(cudaMemcpyAsync(x.data(), x_dev, sizeof(cplx) * P, cudaMemcpyDeviceToHost, stream()));
    ANGPU_CUDA(cudaStreamSynchronize(stream()));
    get_params(p.data());
    for(unsigned k = 0; k < P; k++) p[k] += alpha * x[k];
    set_params(p.data());
}

// ---------------------------------------------------------------------------------------- PsiRBM

// ANGPU_MC_SCREEN=1 selects the fp32-screened sampler (rbm_sampler.cuh); its fp32 weight table is only kept when it is on
static bool rbm_screen_on() { static const bool v = [] { const char* e = getenv("ANGPU_MC_SCREEN"); return e && atoi(e) != 0; }(); return v; }

// W += alpha x on every device copy of the weights: W itself, the row-padded copy and the fp32 copy of the screened sampler
__global__ void k_rbm_add_params(cplx* __restrict__ W, cplx* __restrict__ Wpad, float4* __restrict__ Wf, const cplx* __restrict__ x,
                                 cplx alpha, unsigned N, unsigned M, unsigned Mpad, unsigned KK) {
    for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < (size_t)N * M; k += (size_t)gridDim.x * blockDim.x) {
        const unsigned i = (unsigned)(k / M), j = (unsigned)(k - (size_t)i * M);
        cplx w = W[k];
        cfma(w, alpha, x[k]);
        W[k] = w;
        if(Wpad) Wpad[(size_t)i * Mpad + j] = w;
        if(Wf) {
            float* f = reinterpret_cast<float*>(&Wf[((size_t)i * KK + (j >> 6)) * 32u + ((j & 63u) >> 1)]);
            f[j & 1u] = (float)w.re; f[2u + (j & 1u)] = (float)w.im;
        }
    }
}
void PsiRBM::add_params_dev(const cplx* x_dev, cplx alpha) {
    const unsigned grid = (unsigned)std::min<size_t>(((size_t)P + 255) / 256, (size_t)ctx().num_sms * 16);
    k_rbm_add_params<<<grid, 256, 0, stream()>>>(dW.p, Mpad != M && Mpad ? dWpad.p : nullptr, dWf.n ? dWf.p : nullptr, x_dev, alpha, N, M, Mpad,
                                                 M <= 512u ? rbm_sampler_KK(M) : 0u);
    ANGPU_CHECK_LAUNCH(); count_launch();
    host_stale = true;
}
void PsiRBM::sync_host() const {
    if(!host_stale) return;
    dW.download(hW.data(), hW.size());
    host_stale = false;
}

PsiRBM::PsiRBM(unsigned N_, unsigned M_, const cplx* W, cplx fw_, cplx lp_) {
    kind = RBM; N = N_; M = M_; words = words_for(N); P = N * M; lp = lp_; fw = fw_;
    ANGPU_REQUIRE(N >= 1 && N <= 64u * MAXW, "PsiRBM: 1 <= N <= 256");
    ANGPU_REQUIRE(M >= 1, "PsiRBM: M >= 1");
    hW.assign(W, W + (size_t)N * M);
    upload();
}
void PsiRBM::upload(const cplx* src) {
    // the host -> device copy is issued first and straight from the caller's buffer (pinned in the end-to-end path); the host
    // mirror is refreshed while it is in flight
    dW.resize((size_t)N * M);
    ANGPU_CUDA(cudaMemcpyAsync(dW.p, src ? src : hW.data(), sizeof(cplx) * (size_t)N * M, cudaMemcpyHostToDevice, stream()));
    if(src) { hW.assign(src, src + (size_t)N * M); host_stale = false; }
    // rows padded to a multiple of 32*K (warp sampler) / 256 (block sampler) complex for the register-resident samplers
    // (rbm_kernels.cuh); when M already is such a multiple (C2: 256) the samplers read W itself
    Mpad = 0;
    if(M <= 2048u) {
        Mpad = (M <= 512u) ? 64u * rbm_sampler_KK(M) : (unsigned)MC_BLOCK_T * ((M + MC_BLOCK_T - 1) / MC_BLOCK_T);
        if(Mpad != M) {
            std::vector<cplx> wp((size_t)N * Mpad, cplx(0.0, 0.0));
            for(unsigned i = 0; i < N; i++) std::memcpy(&wp[(size_t)i * Mpad], &hW[(size_t)i * M], sizeof(cplx) * M);
            dWpad.upload(wp);
        }
        if(M <= 512u && rbm_screen_on()) {
            // fp32 copy for the (opt-in) screened sampler (rbm_sampler.cuh): float4 (Re W_pj0, Re W_pj1, Im W_pj0, Im W_pj1) at
            // [(p KK + kk) 32 + lane], j0 = 64 kk + 2 lane, zeros beyond M
            const unsigned KK = rbm_sampler_KK(M);
            std::vector<float4> wf((size_t)N * KK * 32u);
            for(unsigned i = 0; i < N; i++)
                for(unsigned kk = 0; kk < KK; kk++)
                    for(unsigned l = 0; l < 32u; l++) {
                        const unsigned j0 = 64u * kk + 2u * l;
                        const cplx a = j0 < M ? hW[(size_t)i * M + j0] : cplx(0.0, 0.0), b = j0 + 1u < M ? hW[(size_t)i * M + j0 + 1u] : cplx(0.0, 0.0);
                        wf[((size_t)i * KK + kk) * 32u + l] = make_float4((float)a.re, (float)b.re, (float)a.im, (float)b.im);
                    }
            dWf.upload(wf);
        }
    }
    ANGPU_CUDA(cudaStreamSynchronize(stream()));          // src may be short-lived
}
// theta = sigma W for a batch of configurations: the FP64 tensor-core GEMM (k_rbm_angles_dmma) for batches that fill the
// m8 tiles, the warp-per-configuration kernel for probes of a few configurations; ANGPU_ANGLES=fma forces the latter
static void rbm_angles(const RbmDev& d, const uint64_t* confs, size_t ns, cplx* angles, cplx* log_psi, double* weight) {
    static const bool force_fma = [] { const char* e = getenv("ANGPU_ANGLES"); return e && std::string(e) == "fma"; }();
    if(ns >= 64 && !force_fma) {
        const bool wide = 2u * d.M > 64u;
        static bool attr = false;
        if(!attr) {
            ANGPU_CUDA(cudaFuncSetAttribute(k_rbm_angles_dmma<8, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ag_smem(128)));
            ANGPU_CUDA(cudaFuncSetAttribute(k_rbm_angles_dmma<8, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ag_smem(64)));
            attr = true;
        }
        if(wide) k_rbm_angles_dmma<8, 128><<<ceil_div(ns, 64), 256, ag_smem(128), stream()>>>(d, confs, ns, angles, log_psi, weight);
        else k_rbm_angles_dmma<8, 64><<<ceil_div(ns, 64), 256, ag_smem(64), stream()>>>(d, confs, ns, angles, log_psi, weight);
    } else {
        const unsigned wpb = 8, grid = (unsigned)std::min<size_t>((ns + wpb - 1) / wpb, (size_t)ctx().num_sms * 16);
        k_rbm_angles<<<grid, wpb * 32, 0, stream()>>>(d, confs, ns, angles, log_psi, weight);
    }
    ANGPU_CHECK_LAUNCH(); count_launch();
}
void PsiRBM::log_psi(SampleSet& S, bool es_weights) {
    if(S.ns == 0) return;
    S.angles.resize(S.ns * M);
    rbm_angles(dev(), S.conf.p, S.ns, S.angles.p, S.log_psi.p, es_weights ? S.weight.p : nullptr);
    S.has_angles = true;
}
void PsiRBM::ensure_angles(SampleSet& S) {
    if(S.has_angles || S.ns == 0) return;
    S.angles.resize(S.ns * M);
    rbm_angles(dev(), S.conf.p, S.ns, S.angles.p, nullptr, nullptr);
    S.has_angles = true;
}
template<int WPS>
static void launch_eloc_rbm(const RbmDev& d, const Operator& op, SampleSet& S) {
    constexpr unsigned TEAMS = RBM_ELOC_WARPS / WPS;
    const size_t smem = TEAMS * rbm_eloc_slice_bytes(d.M, op.dev.num_groups, WPS);
    const unsigned grid = (unsigned)std::min<size_t>((S.ns + TEAMS - 1) / TEAMS, (size_t)ctx().num_sms * 32);
    auto launch = [&](auto kernel) {
        set_smem(kernel, smem);
        kernel<<<grid, RBM_ELOC_WARPS * 32, smem, stream()>>>(d, op.dev, S.conf.p, S.angles.p, S.ns, S.eloc.p);
    };
    switch(op.dev.max_flips) {
        case 0: case 1: launch(k_eloc_rbm<1, WPS>); break;
        case 2: launch(k_eloc_rbm<2, WPS>); break;
        case 3: launch(k_eloc_rbm<3, WPS>); break;
        default: launch(k_eloc_rbm<4, WPS>); break;
    }
    ANGPU_CHECK_LAUNCH(); count_launch();
}
// the tile kernel (k_eloc_rbm_tile): M <= 256, <= 2 flips per group; the tile size makes the tiles fill the resident blocks in whole
// rounds (ET_BLOCKS_PER_SM blocks per SM): the smallest number of rounds R with ceil(ns / (slots R)) <= 16 samples per tile
template<int K>
static bool launch_eloc_rbm_tile(const RbmDev& d, const Operator& op, SampleSet& S) {
    const size_t slots = (size_t)ctx().num_sms * ET_BLOCKS_PER_SM;
    unsigned st = 0;
    for(size_t R = 1; R <= 64 && !st; R++) { const size_t t = (S.ns + slots * R - 1) / (slots * R); if(t <= 16) st = (unsigned)std::max<size_t>(t, 1); }
    if(!st) st = 16;
    const size_t smem = eloc_tile_smem(32u * K, st, d.words);
    if(smem > (size_t)ctx().smem_optin / ET_BLOCKS_PER_SM) return false;
    set_smem(k_eloc_rbm_tile<K>, smem);
    const unsigned grid = (unsigned)std::min<size_t>((S.ns + st - 1) / st, slots * 8);
    k_eloc_rbm_tile<K><<<grid, ET_WARPS * 32, smem, stream()>>>(d, op.dev, S.conf.p, S.angles.p, S.ns, st, S.eloc.p);
    ANGPU_CHECK_LAUNCH(); count_launch();
    return true;
}
void PsiRBM::eloc(const Operator& op, SampleSet& S) {
    require_operator_fits(op, N, words);
    if(S.ns == 0) return;
    const size_t budget = std::min<size_t>(ctx().smem_optin, 200 * 1024);
    if(op.dev.max_flips > (unsigned)RBM_ELOC_MAXF || rbm_eloc_slice_bytes(M, op.dev.num_groups, RBM_ELOC_WARPS) > budget) { generic_eloc(dev(), op, S); return; }
    ensure_angles(S);
    {
        // M <= 256, <= 2 flips per group, enough samples to fill the GPU: W rows in registers, a tile of samples per block
        static const bool tile_on = [] { const char* e = getenv("ANGPU_ELOC_TILE"); return !(e && atoi(e) == 0); }();
        // (narrow layers, M <= 64, stay on the team kernel: measured at C1, M = 32, the tile kernel costs 0.08 ms of 0.31)
        if(tile_on && M > 64u && M <= 256u && op.dev.max_flips <= 2u && op.dev.num_groups >= 1u && S.ns >= (size_t)ctx().num_sms * 8) {
            const RbmDev d = dev();
            bool done = false;
            if(M <= 128u) done = launch_eloc_rbm_tile<4>(d, op, S);
            else done = launch_eloc_rbm_tile<8>(d, op, S);
            if(done) return;
        }
    }
    // warps per sample: as few as keep >= 4 blocks (32 warps) per SM resident, but at least 2 (finer scheduling units);
    // ANGPU_ELOC_WPS overrides (A/B timing)
    const char* env = getenv("ANGPU_ELOC_WPS");
    unsigned wps = env ? (unsigned)atoi(env) : 0u;
    if(wps != 1u && wps != 2u && wps != 4u && wps != 8u) {
        wps = 2u;
        while(wps < (unsigned)RBM_ELOC_WARPS && (RBM_ELOC_WARPS / wps) * rbm_eloc_slice_bytes(M, op.dev.num_groups, wps) > (size_t)ctx().smem_optin / 4) wps *= 2u;
    }
    const RbmDev d = dev();
    switch(wps) {
        case 1: launch_eloc_rbm<1>(d, op, S); break;
        case 2: launch_eloc_rbm<2>(d, op, S); break;
        case 4: launch_eloc_rbm<4>(d, op, S); break;
        default: launch_eloc_rbm<8>(d, op, S); break;
    }
}
void PsiRBM::compute_T(SampleSet& S, DevBuf<cplx>& T) {
    ensure_angles(S);
    T.resize(S.ns * M);
    if(S.ns == 0) return;
    const size_t total = S.ns * M;
    const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)ctx().num_sms * 32);
    k_rbm_T<<<grid, 256, 0, stream()>>>(dev(), S.angles.p, total, T.p);
    ANGPU_CHECK_LAUNCH(); count_launch();
}
void PsiRBM::ok_rows(SampleSet& S, size_t s0, size_t cnt, cplx* out) {
    if(cnt == 0) return;
    DevBuf<cplx>& T = T_scratch;                   // grow-only member: no allocation / stream sync per call
    compute_T(S, T);
    k_rbm_dense_O<<<(unsigned)cnt, 256, 0, stream()>>>(dev(), S.conf.p + s0 * words, T.p + s0 * M, cnt, out);
    ANGPU_CHECK_LAUNCH(); count_launch();
}

// initial configurations + theta_0 = sigma_0 W (tensor-core GEMM) into the chains' first sample slots; false when there is
// no slot to use (no recorded sample), in which case the samplers initialise themselves
static bool rbm_preinit(const RbmDev& d, const McParams& mc, SampleSet& S) {
    if(mc.steps_per_chain == 0u) return false;
    k_mc_init_conf<<<ceil_div(mc.num_chains_local, 128), 128, 0, stream()>>>(mc, d.N, d.words, S.conf.p);
    ANGPU_CHECK_LAUNCH(); count_launch();
    rbm_angles(d, S.conf.p, mc.num_chains_local, S.angles.p, nullptr, nullptr);
    return true;
}

template<int K, int WORDS>
static void launch_mc_rbm(const RbmDev& d, const cplx* Wp, const McParams& mc, SampleSet& S, unsigned long long* acc_rej_dev) {
    const unsigned wpb = MC_RBM_THREADS / 32, grid = ceil_div(mc.num_chains_local, wpb);
    // the fp64 sampler (complex final weight, or ANGPU_MC_SCREEN=0): resident blocks per SM requested from ptxas: 8
    // (<= 128 registers); for K = 8 with one-word configurations 10 blocks = 20 warps per SM fit in 96 registers
    constexpr int MINB = (K == 8 && WORDS == 1) ? 10 : (K <= 8) ? 8 : 1;
    cplx* angles = rbm_preinit(d, mc, S) ? S.angles.p : nullptr;      // non-null: the chains start from the pre-computed theta_0
    if(d.fw.im == 0.0) k_mc_rbm<K, WORDS, true, MINB><<<grid, wpb * 32, 0, stream()>>>(d, Wp, mc, S.conf.p, S.log_psi.p, angles, acc_rej_dev);
    else k_mc_rbm<K, WORDS, false, MINB><<<grid, wpb * 32, 0, stream()>>>(d, Wp, mc, S.conf.p, S.log_psi.p, angles, acc_rej_dev);
    ANGPU_CHECK_LAUNCH(); count_launch();
}
// the fp32-screened sampler (rbm_sampler.cuh): real final weight
template<int KK, int WORDS>
static void launch_mc_rbm_scr(const RbmDev& d, const cplx* Wp, const float4* Wf, const McParams& mc, SampleSet& S, unsigned long long* stats) {
    const unsigned wpb = MC_RBM_THREADS / 32, grid = ceil_div(mc.num_chains_local, wpb);
    static const int refresh = [] { const char* e = getenv("ANGPU_MC_REFRESH"); return e ? atoi(e) : 16; }();   // 16 or 32 accepted flips between syncs
    static const int minb = [] { const char* e = getenv("ANGPU_MC_MINB"); return e ? atoi(e) : 8; }();
    auto go = [&](auto kern) { kern<<<grid, wpb * 32, 0, stream()>>>(d, Wp, Wf, mc, S.conf.p, S.log_psi.p, S.angles.p, stats); };
    if(KK <= 4 && minb >= 10) { if(refresh >= 32) go(k_mc_rbm_scr<KK, WORDS, (KK <= 4) ? 10 : 4, 32>); else go(k_mc_rbm_scr<KK, WORDS, (KK <= 4) ? 10 : 4, 16>); }
    else { if(refresh >= 32) go(k_mc_rbm_scr<KK, WORDS, (KK <= 4) ? 8 : 4, 32>); else go(k_mc_rbm_scr<KK, WORDS, (KK <= 4) ? 8 : 4, 16>); }
    ANGPU_CHECK_LAUNCH(); count_launch();
}
template<int KK>
static void launch_mc_rbm_kk(const RbmDev& d, const cplx* Wp, const float4* Wf, bool screened, const McParams& mc, SampleSet& S, unsigned long long* a) {
    if(screened) {
        switch(d.words) {
            case 1: launch_mc_rbm_scr<KK, 1>(d, Wp, Wf, mc, S, a); break;
            case 2: launch_mc_rbm_scr<KK, 2>(d, Wp, Wf, mc, S, a); break;
            case 3: launch_mc_rbm_scr<KK, 3>(d, Wp, Wf, mc, S, a); break;
            default: launch_mc_rbm_scr<KK, 4>(d, Wp, Wf, mc, S, a); break;
        }
        return;
    }
    switch(d.words) {
        case 1: launch_mc_rbm<2 * KK, 1>(d, Wp, mc, S, a); break;
        case 2: launch_mc_rbm<2 * KK, 2>(d, Wp, mc, S, a); break;
        case 3: launch_mc_rbm<2 * KK, 3>(d, Wp, mc, S, a); break;
        default: launch_mc_rbm<2 * KK, 4>(d, Wp, mc, S, a); break;
    }
}

void PsiRBM::mc_sample(const McParams& mc, SampleSet& S, unsigned long long* acc_rej_dev) {
    if(mc.num_chains_local == 0) return;
    const RbmDev d = dev();
    if(M <= 512u) {
        S.angles.resize(S.ns * M);
        // ANGPU_MC_SCREEN=1 selects the fp32-screened sampler (rbm_sampler.cuh): identical chains, but measured SLOWER than
        // the all-fp64 sampler on B200 (C2: 2.3 vs 1.75 ms; FP32 runs at only 2x the FP64 rate, F2F at a quarter of it, and
        // the kernel is issue-bound -- DESIGN.md 9.1), so it is opt-in
        const bool screen_on = rbm_screen_on();
        const unsigned long long steps = (unsigned long long)N * ((unsigned long long)mc.num_therm + (unsigned long long)mc.num_sweeps * mc.steps_per_chain);
        const bool screened = screen_on && fw.im == 0.0 && steps < 0xffffffffull;
        switch(rbm_sampler_KK(M)) {
            case 1: launch_mc_rbm_kk<1>(d, Wpad(), dWf.p, screened, mc, S, acc_rej_dev); break;
            case 2: launch_mc_rbm_kk<2>(d, Wpad(), dWf.p, screened, mc, S, acc_rej_dev); break;
            case 4: launch_mc_rbm_kk<4>(d, Wpad(), dWf.p, screened, mc, S, acc_rej_dev); break;
            default: launch_mc_rbm_kk<8>(d, Wpad(), dWf.p, screened, mc, S, acc_rej_dev); break;
        }
        S.has_angles = true;
    } else if(M <= 2048u) {
        S.angles.resize(S.ns * M);
        const unsigned K = (M + MC_BLOCK_T - 1) / MC_BLOCK_T;          // 3..8
        cplx* angles = rbm_preinit(d, mc, S) ? S.angles.p : nullptr;  // non-null: the chains start from the pre-computed theta_0
        auto launch = [&](auto kr, auto kc) {
            if(d.fw.im == 0.0) kr<<<mc.num_chains_local, MC_BLOCK_T, 0, stream()>>>(d, Wpad(), mc, S.conf.p, S.log_psi.p, angles, acc_rej_dev);
            else kc<<<mc.num_chains_local, MC_BLOCK_T, 0, stream()>>>(d, Wpad(), mc, S.conf.p, S.log_psi.p, angles, acc_rej_dev);
        };
        switch(K) {
            case 3: launch(k_mc_rbm_block<3, true>, k_mc_rbm_block<3, false>); break;
            case 4: launch(k_mc_rbm_block<4, true>, k_mc_rbm_block<4, false>); break;
            case 5: launch(k_mc_rbm_block<5, true>, k_mc_rbm_block<5, false>); break;
            case 6: launch(k_mc_rbm_block<6, true>, k_mc_rbm_block<6, false>); break;
            case 7: launch(k_mc_rbm_block<7, true>, k_mc_rbm_block<7, false>); break;
            default: launch(k_mc_rbm_block<8, true>, k_mc_rbm_block<8, false>); break;
        }
        ANGPU_CHECK_LAUNCH(); count_launch();
        S.has_angles = true;
    } else {
        generic_mc(d, mc, S, acc_rej_dev);
        S.has_angles = false;
    }
}

// ---------------------------------------------------------------------------------------- PsiDeep

PsiDeep::PsiDeep(unsigned num_sites_, unsigned N_, const cplx* input_weights_, unsigned num_hidden, const unsigned* sizes,
                 const unsigned* conn, const cplx* biases, const unsigned* lhs_connections, const cplx* lhs_weights,
                 const cplx* final_weights_, cplx lp_) {
    kind = DEEP; num_sites = num_sites_; N = N_; words = words_for(N_); lp = lp_;
    // N == num_sites: spin basis.  N == 3 num_sites: the Pauli-string basis, one input unit per site and Pauli type
    // (PsiDeep.hpp:282-308, PauliString::network_unit_at) -- to be used with MonteCarloPaulis / ExactSummationPaulis.
    ANGPU_REQUIRE(N == num_sites || N == 3u * num_sites, "PsiDeep: N must be num_sites (spin basis) or 3 num_sites (Pauli-string basis)");
    pauli_sites = (N == 3u * num_sites && num_sites > 0u) ? num_sites : 0u;
    ANGPU_REQUIRE(N >= 1 && N <= 64u * MAXW, "PsiDeep: 1 <= N <= 256");
    ANGPU_REQUIRE(num_hidden >= 1 && num_hidden + 1 <= (unsigned)DEEP_MAX_LAYERS, "PsiDeep: 1..4 hidden layers");
    num_layers = num_hidden + 1;
    input_weights.assign(input_weights_, input_weights_ + N);
    layers.resize(num_layers);
    layers[0].size = N; width = N; P = N;
    size_t off_b = 0, off_w = 0; unsigned deep = 0;
    for(unsigned l = 1; l < num_layers; l++) {
        Layer& L = layers[l];
        L.size = sizes[l - 1]; L.conn = conn[l - 1];
        ANGPU_REQUIRE((size_t)L.size * L.conn % layers[l - 1].size == 0, "PsiDeep: size*conn must be a multiple of the previous layer size");
        width = std::max(width, L.size);
        const size_t nw = (size_t)L.size * L.conn;
        L.bias.assign(biases + off_b, biases + off_b + L.size);
        L.lhs_w.assign(lhs_weights + off_w, lhs_weights + off_w + nw);
        L.lhs_c.assign(lhs_connections + off_w, lhs_connections + off_w + nw);
        for(unsigned c : L.lhs_c) ANGPU_REQUIRE(c < layers[l - 1].size, "PsiDeep: connection index out of range");
        L.begin_params = P; P += L.size + (unsigned)nw;
        if(l > 1) { L.begin_deep = deep; deep += L.size; }
        off_b += L.size; off_w += nw;
    }
    num_deep = deep;
    final_weights.assign(final_weights_, final_weights_ + layers[num_layers - 1].size);
    compile_rhs();
    upload();
}
PsiDeep::PsiDeep(const PsiDeep& o) {
    kind = DEEP; N = o.N; words = o.words; P = o.P; lp = o.lp; pauli_sites = o.pauli_sites;
    num_sites = o.num_sites; num_layers = o.num_layers; width = o.width; num_deep = o.num_deep;
    input_weights = o.input_weights; final_weights = o.final_weights;
    layers.resize(num_layers);
    for(unsigned l = 0; l < num_layers; l++) {
        Layer& a = layers[l]; const Layer& b = o.layers[l];
        a.size = b.size; a.conn = b.conn; a.rhs_conn = b.rhs_conn; a.begin_params = b.begin_params; a.begin_deep = b.begin_deep;
        a.lhs_c = b.lhs_c; a.rhs_c = b.rhs_c; a.lhs_w = b.lhs_w; a.rhs_w = b.rhs_w; a.bias = b.bias;
    }
    upload();
}
// source/quantum_state/PsiDeep.cu:214-242
void PsiDeep::compile_rhs() {
    for(unsigned l = 0; l + 1 < num_layers; l++) {
        Layer& lo = layers[l]; const Layer& hi = layers[l + 1];
        lo.rhs_conn = hi.size * hi.conn / lo.size;
        lo.rhs_c.assign((size_t)lo.size * lo.rhs_conn, 0u);
        lo.rhs_w.assign((size_t)lo.size * lo.rhs_conn, cplx(0.0, 0.0));
        std::vector<unsigned> fill(lo.size, 0u);
        for(unsigned j = 0; j < hi.size; j++)
            for(unsigned i = 0; i < hi.conn; i++) {
                const unsigned lhs = hi.lhs_c[(size_t)i * hi.size + j];
                ANGPU_REQUIRE(fill[lhs] < lo.rhs_conn, "PsiDeep: connectivity is not regular (each unit must feed rhs_connectivity units)");
                lo.rhs_c[(size_t)lhs * lo.rhs_conn + fill[lhs]] = j;
                lo.rhs_w[(size_t)lhs * lo.rhs_conn + fill[lhs]] = hi.lhs_w[(size_t)i * hi.size + j];
                fill[lhs]++;
            }
    }
    layers[num_layers - 1].rhs_conn = 0;
}
void PsiDeep::upload() {
    for(auto& L : layers) {
        L.d_lhs_c.upload(L.lhs_c); L.d_rhs_c.upload(L.rhs_c);
        L.d_lhs_w.upload(L.lhs_w); L.d_rhs_w.upload(L.rhs_w); L.d_bias.upload(L.bias);
    }
    d_final.upload(final_weights);
    // eligibility of the block-per-chain sampler + its dense tables: layer l as wd[k][j] = weight of input k for unit j
    // (0 where unconnected), [N][64] for the first layer followed by [64][64] per deep layer
    block_sampler_ok = (num_layers == 3u || num_layers == 4u);
    for(unsigned l = 1; l < num_layers; l++) if(layers[l].size > (unsigned)DEEP_BLK_W) block_sampler_ok = false;
    if(block_sampler_ok) {
        std::vector<cplx> t((size_t)(N + (num_layers - 2u) * DEEP_BLK_W) * DEEP_BLK_W, cplx(0.0, 0.0));
        std::vector<unsigned char> seen(t.size(), 0);
        size_t base = 0;
        for(unsigned l = 1; l < num_layers && block_sampler_ok; l++) {
            const Layer& L = layers[l];
            for(unsigned i = 0; i < L.conn && block_sampler_ok; i++)
                for(unsigned j = 0; j < L.size; j++) {
                    const size_t at = base + (size_t)L.lhs_c[(size_t)i * L.size + j] * DEEP_BLK_W + j;
                    if(seen[at]) { block_sampler_ok = false; break; }      // a unit reading the same input twice
                    seen[at] = 1; t[at] = L.lhs_w[(size_t)i * L.size + j];
                }
            base += (size_t)(l == 1 ? N : (unsigned)DEEP_BLK_W) * DEEP_BLK_W;
        }
        if(block_sampler_ok) d_w1dense.upload(t);
    }
}
DeepDev PsiDeep::dev() const {
    DeepDev d{};
    d.N = N; d.words = words; d.P = P; d.num_layers = num_layers; d.width = width; d.num_deep = num_deep; d.lp = lp;
    for(unsigned l = 0; l < num_layers; l++) {
        const Layer& L = layers[l];
        d.L[l] = DeepLayerDev{L.size, L.conn, L.rhs_conn, L.begin_params, L.begin_deep,
                              L.d_lhs_c.p, L.d_rhs_c.p, L.d_lhs_w.p, L.d_rhs_w.p, L.d_bias.p};
    }
    d.final_w = d_final.p;
    return d;
}
// source/quantum_state/PsiDeep.cu:246-270
void PsiDeep::get_params(cplx* out) const {
    std::memcpy(out, input_weights.data(), sizeof(cplx) * N); out += N;
    for(unsigned l = 1; l < num_layers; l++) {
        const Layer& L = layers[l];
        std::memcpy(out, L.bias.data(), sizeof(cplx) * L.size); out += L.size;
        std::memcpy(out, L.lhs_w.data(), sizeof(cplx) * L.lhs_w.size()); out += L.lhs_w.size();
    }
}
// source/quantum_state/PsiDeep.cu:274-305
void PsiDeep::set_params(const cplx* in) {
    input_weights.assign(in, in + N); in += N;
    for(unsigned l = 1; l < num_layers; l++) {
        Layer& L = layers[l];
        L.bias.assign(in, in + L.size); in += L.size;
        const size_t nw = L.lhs_w.size();
        L.lhs_w.assign(in, in + nw); in += nw;
    }
    compile_rhs();
    upload();
}
void PsiDeep::log_psi(SampleSet& S, bool es_weights) { generic_log_psi(dev(), S, es_weights); }
void PsiDeep::eloc(const Operator& op, SampleSet& S) {
    ANGPU_REQUIRE(S.pauli_sites == pauli_sites, pauli_sites ? "PsiDeep on the Pauli-string basis (N = 3 num_sites): use MonteCarloPaulis / ExactSummationPaulis"
                                                            : "Pauli-string ensembles need a PsiDeep with N = 3 num_sites input units");
    if(pauli_sites) {
        ANGPU_REQUIRE(op.words == words_for(num_sites) && op.words <= (unsigned)PAULI_SITE_WORDS, "operator / wavefunction word count mismatch");
        ANGPU_REQUIRE(op.num_sites_touched <= num_sites, "operator acts on site " + std::to_string(op.num_sites_touched - 1) + " but the wavefunction has " + std::to_string(num_sites) + " sites");
        if(S.ns == 0) return;
        const DeepDev d = dev();
        const WarpCfg c = warp_cfg(d.payload_elems(), S.ns, d.block_scratch_bytes());
        set_smem(k_eloc_paulis<DeepDev>, c.smem);
        const PauliOpDev pop{op.num_strings, op.words, op.d_pcoef.p, op.d_pa.p, op.d_pb.p};
        k_eloc_paulis<DeepDev><<<c.grid, c.wpb * 32, c.smem, stream()>>>(d, pop, num_sites, S.conf.p, S.log_psi.p, S.ns, S.eloc.p);
        ANGPU_CHECK_LAUNCH(); count_launch();
        return;
    }
    require_operator_fits(op, N, words);
    if(S.ns == 0) return;
    const char* env_s = getenv("ANGPU_DEEP_ELOC");              // "generic" forces the warp-per-sample kernel
    const size_t dyn = (size_t)op.dev.num_groups * (sizeof(cplx) + sizeof(unsigned)) + 16;
    if(!block_sampler_ok || (env_s && std::string(env_s) == "generic") || op.dev.num_groups == 0 || dyn > 64 * 1024) { generic_eloc(dev(), op, S); return; }
    const unsigned grid = (unsigned)std::min<size_t>(S.ns, (size_t)ctx().num_sms * 8);
    if(num_layers == 3u) { set_smem(k_eloc_deep_block<1>, dyn); k_eloc_deep_block<1><<<grid, DEEP_BLK_T, dyn, stream()>>>(dev(), d_w1dense.p, op.dev, S.conf.p, S.ns, S.eloc.p); }
    else                 { set_smem(k_eloc_deep_block<2>, dyn); k_eloc_deep_block<2><<<grid, DEEP_BLK_T, dyn, stream()>>>(dev(), d_w1dense.p, op.dev, S.conf.p, S.ns, S.eloc.p); }
    ANGPU_CHECK_LAUNCH(); count_launch();
}
void PsiDeep::ok_rows(SampleSet& S, size_t s0, size_t cnt, cplx* out) { generic_ok(dev(), S, s0, cnt, out); }
void PsiDeep::mc_sample(const McParams& mc, SampleSet& S, unsigned long long* a) {
    ANGPU_REQUIRE(mc.pauli_sites == pauli_sites, pauli_sites ? "PsiDeep on the Pauli-string basis (N = 3 num_sites): use MonteCarloPaulis"
                                                             : "MonteCarloPaulis needs a PsiDeep with N = 3 num_sites input units");
    if(pauli_sites) {
        if(mc.num_chains_local == 0) return;
        const DeepDev d = dev();
        const size_t slice = warp_slice_bytes(d.payload_elems());
        const size_t budget = std::min<size_t>(ctx().smem_optin, 200 * 1024);
        ANGPU_REQUIRE(slice <= budget, "model scratch does not fit in shared memory");
        unsigned wpb = (unsigned)std::min<size_t>(4, budget / slice);
        while(wpb > 1 && wpb * slice > budget / 2) wpb--;
        set_smem(k_mc_paulis<DeepDev>, wpb * slice);
        k_mc_paulis<DeepDev><<<ceil_div(mc.num_chains_local, wpb), wpb * 32, wpb * slice, stream()>>>(d, mc, S.conf.p, S.log_psi.p, a);
        ANGPU_CHECK_LAUNCH(); count_launch();
        S.has_angles = false;
        return;
    }
    const char* env_s = getenv("ANGPU_DEEP_SAMPLER");          // "generic" forces the warp-per-chain kernel (tests, A/B timing)
    const bool force_generic = env_s && std::string(env_s) == "generic";
    if(!block_sampler_ok || force_generic) { generic_mc(dev(), mc, S, a); return; }
    if(mc.num_chains_local == 0) return;
    const unsigned grid = ceil_div(mc.num_chains_local, (unsigned)DEEP_BLK_NC);
    if(num_layers == 3u) k_mc_deep_block<1><<<grid, DEEP_BLK_T, 0, stream()>>>(dev(), d_w1dense.p, mc, S.conf.p, S.log_psi.p, a);
    else                 k_mc_deep_block<2><<<grid, DEEP_BLK_T, 0, stream()>>>(dev(), d_w1dense.p, mc, S.conf.p, S.log_psi.p, a);
    ANGPU_CHECK_LAUNCH(); count_launch();
}

// ---------------------------------------------------------------------------------------- PsiCNN

PsiCNN::PsiCNN(const unsigned* extent_, unsigned num_layers_, const unsigned* num_channels_, const unsigned* connectivity_,
               const unsigned* symmetry_classes, const cplx* params_, unsigned num_params, double final_factor_, cplx lp_) {
    kind = CNN; lp = lp_; final_factor = final_factor_; num_layers = num_layers_;
    for(int d = 0; d < 3; d++) extent[d] = extent_[d];
    N = extent[0] * extent[1] * extent[2]; words = words_for(N); P = num_params;
    ANGPU_REQUIRE(N >= 1 && N <= 64u * MAXW, "PsiCNN: 1 <= N <= 256");
    ANGPU_REQUIRE(num_layers >= 1 && num_layers <= (unsigned)CNN_MAX_LAYERS, "PsiCNN: 1..4 layers");
    num_channels.assign(num_channels_, num_channels_ + num_layers);
    connectivity.assign(connectivity_, connectivity_ + 3 * num_layers);
    sym.assign(symmetry_classes, symmetry_classes + N);
    params.assign(params_, params_ + P);
    build();
}
// source/quantum_state/PsiCNN.cpp:10-55 (parameter layout), detail/Convolve.hpp:108-141 (neighbours)
void PsiCNN::build() {
    // distinct symmetry-class labels must be 0..num_sym-1 for the weight layout sym*vol + c to be dense
    std::set<unsigned> classes(sym.begin(), sym.end());
    num_sym = (unsigned)classes.size();
    for(unsigned s : sym) ANGPU_REQUIRE(s < num_sym, "PsiCNN: symmetry classes must be labelled 0..num_classes-1");
    unsigned off = 0; num_angles = 0; maxch = 1;
    h_nbr.assign(num_layers, {}); h_inv.assign(num_layers, {});
    d_nbr.clear(); d_inv.clear(); d_nbr.resize(num_layers); d_inv.resize(num_layers);
    const unsigned page = extent[1] * extent[2];
    for(unsigned l = 0; l < num_layers; l++) {
        CnnLayerDev& L = layer_dev[l];
        L.nch = num_channels[l]; L.prev = l > 0 ? num_channels[l - 1] : 1u;
        ANGPU_REQUIRE(L.nch * L.prev <= (unsigned)CNN_MAX_LINKS, "PsiCNN: too many channel links in a layer");
        maxch = std::max(maxch, L.nch);
        const unsigned* c = &connectivity[3 * l];
        L.vol = c[0] * c[1] * c[2];
        L.begin_params = off;
        for(unsigned ci = 0; ci < L.prev; ci++) for(unsigned cj = 0; cj < L.nch; cj++) { L.link_begin[ci * L.nch + cj] = off; off += num_sym * L.vol; }
        L.num_params = off - L.begin_params;
        L.angle_off = num_angles; num_angles += L.nch * N;
        h_nbr[l].resize((size_t)N * L.vol); h_inv[l].resize((size_t)N * L.vol);
        for(unsigned x = 0; x < N; x++) {
            const unsigned pg = x / page, row = (x % page) / extent[2], col = (x % page) % extent[2];
            unsigned cidx = 0;
            for(unsigned k = 0; k < c[0]; k++) for(unsigned i = 0; i < c[1]; i++) for(unsigned j = 0; j < c[2]; j++, cidx++) {
                const unsigned y = ((pg + k) % extent[0]) * page + ((row + i) % extent[1]) * extent[2] + ((col + j) % extent[2]);
                h_nbr[l][(size_t)x * L.vol + cidx] = y;
            }
        }
        // inverse: for every (y, c) the unique x with nbr(x, c) == y
        for(unsigned x = 0; x < N; x++) for(unsigned cidx = 0; cidx < L.vol; cidx++) h_inv[l][(size_t)h_nbr[l][(size_t)x * L.vol + cidx] * L.vol + cidx] = x;
        d_nbr[l].upload(h_nbr[l]); d_inv[l].upload(h_inv[l]);
        L.nbr = d_nbr[l].p; L.inv = d_inv[l].p;
    }
    ANGPU_REQUIRE(off == P, "PsiCNN: parameter count does not match the layer description");
    d_sym.upload(sym); d_params.upload(params);
    // receptive cones for the incremental sampler: affected_0(p) = {x : p in nbr_0(x, .)}, affected_l = preimage of affected_{l-1}
    d_aff.clear(); d_aff_cnt.clear(); d_aff.resize(num_layers); d_aff_cnt.resize(num_layers);
    h_cone.assign(num_layers, std::vector<std::vector<unsigned>>(N));
    auto& cone = h_cone;
    gaff_hash = 0; gaff_groups = 0; d_gaff.clear();
    for(unsigned l = 0; l < num_layers; l++) {
        const unsigned vol = layer_dev[l].vol;
        aff_max[l] = 0;
        for(unsigned p = 0; p < N; p++) {
            std::set<unsigned> out;
            if(l == 0) for(unsigned c = 0; c < vol; c++) out.insert(h_inv[0][(size_t)p * vol + c]);
            else for(unsigned y : cone[l - 1][p]) for(unsigned c = 0; c < vol; c++) out.insert(h_inv[l][(size_t)y * vol + c]);
            cone[l][p].assign(out.begin(), out.end());
            aff_max[l] = std::max<unsigned>(aff_max[l], (unsigned)out.size());
        }
        std::vector<unsigned> flat((size_t)N * aff_max[l], 0u), cnt(N);
        for(unsigned p = 0; p < N; p++) {
            cnt[p] = (unsigned)cone[l][p].size();
            std::copy(cone[l][p].begin(), cone[l][p].end(), flat.begin() + (size_t)p * aff_max[l]);
        }
        d_aff[l].upload(flat); d_aff_cnt[l].upload(cnt);
    }
}
CnnDev PsiCNN::dev(bool keep_angles) const {
    CnnDev d{};
    d.keep_angles = keep_angles;
    d.N = N; d.words = words; d.P = P; d.num_layers = num_layers; d.num_sym = num_sym; d.num_angles = num_angles; d.maxch = maxch;
    d.final_factor = final_factor; d.lp = lp; d.sym = d_sym.p; d.params = d_params.p;
    for(unsigned l = 0; l < num_layers; l++) d.L[l] = layer_dev[l];
    return d;
}
// only O_k (back-propagation) needs the recorded pre-activations: the other kernels run with the smaller per-warp scratch
void PsiCNN::log_psi(SampleSet& S, bool es_weights) { generic_log_psi(dev(false), S, es_weights); }
void PsiCNN::eloc(const Operator& op, SampleSet& S) {
    require_operator_fits(op, N, words);
    if(S.ns == 0) return;
    const CnnDev d = dev(false);
    const size_t bb = d.block_scratch_bytes();
    const char* env = getenv("ANGPU_CNN_ELOC");                // "generic" forces one full forward per flip group
    if((env && std::string(env) == "generic") || bb == 0 || op.dev.num_groups == 0) { generic_eloc(d, op, S); return; }
    ANGPU_REQUIRE(op.words == words, "operator / wavefunction word count mismatch");
    // union of the receptive cones of every flip group's sites, per layer (cached per operator)
    const unsigned G = op.dev.num_groups;
    uint64_t key = 1469598103934665603ull;                      // FNV-1a over the flip masks: the cache is keyed by content
    for(uint64_t m : op.h_flip) { key ^= m; key *= 1099511628211ull; }
    key ^= G; key *= 1099511628211ull;
    if(gaff_hash != key || gaff_groups != G || d_gaff.size() != num_layers) {
        d_gaff.clear(); d_gaff_cnt.clear(); d_gaff.resize(num_layers); d_gaff_cnt.resize(num_layers);
        for(unsigned l = 0; l < num_layers; l++) {
            std::vector<std::vector<unsigned>> uni(G);
            gaff_max[l] = 0;
            for(unsigned g = 0; g < G; g++) {
                std::set<unsigned> u;
                for(unsigned w = 0; w < words; w++) {
                    uint64_t m = op.h_flip[(size_t)g * words + w];
                    while(m) {
                        const unsigned p = w * 64u + (unsigned)__builtin_ctzll(m);
                        ANGPU_REQUIRE(p < N, "operator acts on a site beyond the lattice");
                        u.insert(h_cone[l][p].begin(), h_cone[l][p].end());
                        m &= m - 1ull;
                    }
                }
                uni[g].assign(u.begin(), u.end());
                gaff_max[l] = std::max<unsigned>(gaff_max[l], (unsigned)u.size());
            }
            std::vector<unsigned> flat((size_t)G * gaff_max[l], 0u), cnt(G);
            for(unsigned g = 0; g < G; g++) { cnt[g] = (unsigned)uni[g].size(); std::copy(uni[g].begin(), uni[g].end(), flat.begin() + (size_t)g * gaff_max[l]); }
            d_gaff[l].upload(flat); d_gaff_cnt[l].upload(cnt);
        }
        gaff_hash = key; gaff_groups = G;
    }
    CnnIncDev inc{};
    inc.backup_elems = 0;
    for(unsigned l = 0; l < num_layers; l++) {
        inc.aff[l] = d_gaff[l].p; inc.aff_cnt[l] = d_gaff_cnt[l].p; inc.aff_max[l] = gaff_max[l];
        inc.backup_elems += layer_dev[l].nch * gaff_max[l];
    }
    const size_t slice = cnn_inc_slice_bytes(d, inc), cap = ctx().smem_optin;
    if(slice + bb > cap) { generic_eloc(d, op, S); return; }
    unsigned wpb = 1, best = 0;
    for(unsigned w = 1; w <= 4; w++) {
        if(w * slice + bb > cap) break;
        const unsigned resident = w * (unsigned)std::min<size_t>(32, (size_t)(228 * 1024) / (w * slice + bb + 1024));
        if(resident >= best) { best = resident; wpb = w; }
    }
    const size_t smem = wpb * slice + bb;
    const unsigned grid = (unsigned)std::min<size_t>((S.ns + wpb - 1) / wpb, (size_t)ctx().num_sms * 32);
    set_smem(k_eloc_cnn_inc, smem);
    k_eloc_cnn_inc<<<grid, wpb * 32, smem, stream()>>>(d, inc, op.dev, S.conf.p, S.ns, S.eloc.p);
    ANGPU_CHECK_LAUNCH(); count_launch();
}
void PsiCNN::ok_rows(SampleSet& S, size_t s0, size_t cnt, cplx* out) { generic_ok(dev(true), S, s0, cnt, out); }
void PsiCNN::mc_sample(const McParams& mc, SampleSet& S, unsigned long long* a) {
    const CnnDev d = dev(false);
    CnnIncDev inc{};
    inc.backup_elems = 0;
    for(unsigned l = 0; l < num_layers; l++) {
        inc.aff[l] = d_aff[l].p; inc.aff_cnt[l] = d_aff_cnt[l].p; inc.aff_max[l] = aff_max[l];
        inc.backup_elems += layer_dev[l].nch * aff_max[l];
    }
    const size_t slice = cnn_inc_slice_bytes(d, inc), bb = d.block_scratch_bytes();
    const size_t cap = ctx().smem_optin;
    const char* env = getenv("ANGPU_CNN_SAMPLER");             // "generic" forces the full-forward kernel (tests, A/B timing)
    if((env && std::string(env) == "generic") || bb == 0 || slice + bb > cap) { generic_mc(d, mc, S, a); return; }
    if(mc.num_chains_local == 0) return;
    // warps per block: the choice that keeps the most warps resident per SM (shared memory is the limit)
    unsigned wpb = 1, best = 0;
    for(unsigned w = 1; w <= 4; w++) {
        if(w * slice + bb > cap) break;
        const unsigned resident = w * (unsigned)std::min<size_t>(32, (size_t)(228 * 1024) / (w * slice + bb + 1024));
        if(resident >= best) { best = resident; wpb = w; }
    }
    const size_t smem = wpb * slice + bb;
    set_smem(k_mc_cnn_inc, smem);
    k_mc_cnn_inc<<<ceil_div(mc.num_chains_local, wpb), wpb * 32, smem, stream()>>>(d, inc, mc, S.conf.p, S.log_psi.p, a);
    ANGPU_CHECK_LAUNCH(); count_launch();
}

// ---------------------------------------------------------------------------------------- PsiClassical

PsiClassical::PsiClassical(unsigned num_sites, unsigned order_, unsigned num_ops_, const Operator* const* ops_, const cplx* params_,
                           unsigned num_own, const PsiCNN* ref_, cplx lp_) {
    kind = CLASSICAL; N = num_sites; words = words_for(N); order = order_; num_ops = num_ops_; lp = lp_;
    ANGPU_REQUIRE(order == 1u || order == 2u, "PsiClassical: order must be 1 or 2");
    ANGPU_REQUIRE(num_own == num_ops, "PsiClassical: one parameter per local operator");
    for(unsigned i = 0; i < num_ops; i++) {
        require_operator_fits(*ops_[i], N, words);
        ops.emplace_back(new Operator(*ops_[i]));
    }
    own_params.assign(params_, params_ + num_own);
    if(ref_) ref.reset(static_cast<PsiCNN*>(ref_->clone()));
    P = num_own + ((order > 1u && ref) ? ref->P : 0u);
    upload();
}
void PsiClassical::upload() {
    std::vector<OpDev> v;
    for(auto& o : ops) v.push_back(o->dev);
    d_ops.upload(v); d_params.upload(own_params);
}
ClassicalDev PsiClassical::dev() const {
    ClassicalDev d{};
    d.N = N; d.words = words; d.P = P; d.num_ops = num_ops; d.order = order; d.lp = lp;
    d.ops = d_ops.p; d.params = d_params.p; d.has_ref = (bool)ref;
    if(ref) d.ref = ref->dev();
    return d;
}
Psi* PsiClassical::clone() const {
    std::vector<const Operator*> o;
    for(auto& p : ops) o.push_back(p.get());
    return new PsiClassical(N, order, num_ops, o.data(), own_params.data(), (unsigned)own_params.size(), ref.get(), lp);
}
// source/quantum_state/PsiClassical.cu:36-70
void PsiClassical::get_params(cplx* out) const {
    std::memcpy(out, own_params.data(), sizeof(cplx) * own_params.size());
    if(order > 1u && ref) ref->get_params(out + own_params.size());
}
void PsiClassical::set_params(const cplx* in) {
    own_params.assign(in, in + own_params.size());
    d_params.upload(own_params);
    if(order > 1u && ref) ref->set_params(in + own_params.size());
}
void PsiClassical::log_psi(SampleSet& S, bool es_weights) { generic_log_psi(dev(), S, es_weights); }
void PsiClassical::eloc(const Operator& op, SampleSet& S) { require_operator_fits(op, N, words); generic_eloc(dev(), op, S); }
void PsiClassical::ok_rows(SampleSet& S, size_t s0, size_t cnt, cplx* out) { generic_ok(dev(), S, s0, cnt, out); }
void PsiClassical::mc_sample(const McParams& mc, SampleSet& S, unsigned long long* a) { generic_mc(dev(), mc, S, a); }

} // namespace angpu

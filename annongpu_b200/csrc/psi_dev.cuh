// Device-side wavefunction models: by-value views + warp-cooperative evaluation.
//
// One warp evaluates one configuration; hidden units / lattice sites are spread over the 32 lanes, the
// per-configuration scratch ("payload": cached first-layer angles, activations) lives in that warp's slice
// of shared memory, and warp shuffles replace the reference's block barriers + M-way contended shared
// atomics (include/quantum_state/PsiRBM.hpp:100-104, include/cuda_kernel_defines.h:9-15).
//
// Every model exposes the same duck-typed interface:
//   payload_elems()                      number of cplx in the per-warp scratch
//   init(conf, pl)                       init_payload
//   log_psi(conf, pl) -> cplx            log_psi_s     (warp-uniform result)
//   update(old, new, pl)                 update_input_units
//   O_k(conf, pl, row)                   foreach_O_k, written (not accumulated) as one dense row of P entries
// `conf` points at `words` uint64 in shared memory; all 32 lanes call every function together.
#pragma once
#include "common.cuh"
#include "operator.hpp"

namespace angpu {

// ------------------------------------------------------------------------------------------ PsiRBM
// log psi = log_prefactor + final_weight * sum_j lc0(theta_j), theta = W^T s, W[N][M] row-major; no biases;
// parameters = W only (include/quantum_state/PsiRBM.hpp:42-195).
struct RbmDev {
    unsigned N, M, words, P;
    cplx     lp, fw;
    const cplx* W;     // [N][M]
    float    c2fw;     // (float)(2 Re fw): scale of the fp32-screened acceptance test (rbm_sampler.cuh)

    __host__ __device__ unsigned payload_elems() const { return M; }
    __host__ __device__ unsigned block_scratch_bytes() const { return 0u; }
#ifdef __CUDACC__
    __device__ const unsigned char* stage(unsigned char*) const { return nullptr; }
    // compute_angles, PsiRBM.hpp:71-80
    __device__ void init(const uint64_t* conf, cplx* pl, const unsigned char* = nullptr) const {
        const unsigned lane = threadIdx.x & 31u;
        for(unsigned j = lane; j < M; j += 32u) {
            cplx a(0.0, 0.0);
            for(unsigned i = 0; i < N; i++) a += spin_at(conf, i) * W[i * M + j];
            pl[j] = a;
        }
        __syncwarp();
    }
    // forward_pass + log_psi_s, PsiRBM.hpp:92-119
    __device__ cplx log_psi(const uint64_t*, cplx* pl, const unsigned char* = nullptr) const {
        const unsigned lane = threadIdx.x & 31u;
        cplx acc(0.0, 0.0);
        for(unsigned j = lane; j < M; j += 32u) acc += act_lc(pl[j], 0u);
        acc = warp_sum(acc);
        return lp + fw * acc;
    }
    // update_input_units, PsiRBM.hpp:122-157
    __device__ void update(const uint64_t* oldc, const uint64_t* newc, cplx* pl, const unsigned char* = nullptr) const {
        const unsigned lane = threadIdx.x & 31u;
        for(unsigned w = 0; w < words; w++) {
            uint64_t diff = oldc[w] ^ newc[w];
            while(diff) {
                const unsigned pos = w * 64u + (unsigned)__ffsll((long long)diff) - 1u;
                const double delta = spin_at(newc, pos) - spin_at(oldc, pos);
                for(unsigned j = lane; j < M; j += 32u) pl[j] += delta * W[pos * M + j];
                diff &= diff - 1ull;
            }
        }
        __syncwarp();
    }
    // foreach_O_k, PsiRBM.hpp:161-176: O_{i*M+j} = final_weight * th0(theta_j) * s_i
    __device__ void O_k(const uint64_t* conf, cplx* pl, cplx* row, const unsigned char* blk = nullptr) const {
        const unsigned lane = threadIdx.x & 31u;
        for(unsigned j = lane; j < M; j += 32u) {
            const cplx t = fw * act_th(pl[j], 0u);
            for(unsigned i = 0; i < N; i++) row[i * M + j] = spin_at(conf, i) * t;
        }
    }
#endif
};

// ------------------------------------------------------------------------------------------ PsiDeep
// Sparse-connectivity feed-forward net with biases (include/quantum_state/PsiDeep.hpp:71-471).
// layers[0] = input spins; hidden layer l has `size` units, each reading `conn` units of layer l-1.
// Parameter order: [input_weights (N)] then per hidden layer [biases][lhs_weights (conn x size)]
// (source/quantum_state/PsiDeep.cu:246-270).  Quirk kept: input_weights are parameters with O_k = s_i
// although log psi ignores them (PsiDeep.hpp:260-262, 353-362).
constexpr int DEEP_MAX_LAYERS = 5;   // input + up to 4 hidden (reference: max_layers = 4 incl. input)
struct DeepLayerDev {
    unsigned size, conn, rhs_conn, begin_params, begin_deep;
    const unsigned* lhs_c;   // [conn][size]
    const unsigned* rhs_c;   // [size][rhs_conn]
    const cplx*     lhs_w;   // [conn][size]
    const cplx*     rhs_w;   // [size][rhs_conn]
    const cplx*     bias;    // [size]
};
struct DeepDev {
    unsigned N, words, P, num_layers, width, num_deep;
    cplx     lp;
    DeepLayerDev L[DEEP_MAX_LAYERS];
    const cplx* final_w;

    // scratch: angles[L1.size] | act[width] | tmp[width] | deep[num_deep]
    __host__ __device__ unsigned payload_elems() const { return L[1].size + 2u * width + num_deep; }
    __host__ __device__ unsigned block_scratch_bytes() const { return 0u; }
#ifdef __CUDACC__
    __device__ const unsigned char* stage(unsigned char*) const { return nullptr; }
    // compute_angles, PsiDeep.hpp:140-157
    __device__ void init(const uint64_t* conf, cplx* pl, const unsigned char* = nullptr) const {
        const unsigned lane = threadIdx.x & 31u;
        const DeepLayerDev& l1 = L[1];
        for(unsigned j = lane; j < l1.size; j += 32u) {
            cplx a(0.0, 0.0);
            for(unsigned i = 0; i < l1.conn; i++) a += spin_at(conf, l1.lhs_c[i * l1.size + j]) * l1.lhs_w[i * l1.size + j];
            pl[j] = a + l1.bias[j];
        }
        __syncwarp();
    }
    // forward_pass, PsiDeep.hpp:173-215; leaves last-layer activations in act[], deep angles in deep[]
    __device__ cplx forward(cplx* pl) const {
        const unsigned lane = threadIdx.x & 31u;
        cplx* angles = pl; cplx* act = pl + L[1].size; cplx* tmp = act + width; cplx* deep = tmp + width;
        for(unsigned i = lane; i < L[1].size; i += 32u) act[i] = act_lc(angles[i], 0u);
        __syncwarp();
        for(unsigned l = 2; l < num_layers; l++) {
            const DeepLayerDev& ly = L[l];
            for(unsigned j = lane; j < ly.size; j += 32u) {
                cplx a(0.0, 0.0);
                for(unsigned i = 0; i < ly.conn; i++) cfma(a, ly.lhs_w[i * ly.size + j], act[ly.lhs_c[i * ly.size + j]]);
                a += ly.bias[j];
                deep[ly.begin_deep + j] = a;
                tmp[j] = act_lc(a, l - 1u);
            }
            __syncwarp();
            for(unsigned j = lane; j < ly.size; j += 32u) act[j] = tmp[j];
            __syncwarp();
        }
        cplx r(0.0, 0.0);
        const unsigned nf = L[num_layers - 1u].size;
        for(unsigned j = lane; j < nf; j += 32u) cfma(r, act[j], final_w[j]);
        return warp_sum(r);
    }
    __device__ cplx log_psi(const uint64_t*, cplx* pl, const unsigned char* = nullptr) const { return lp + forward(pl); }
    // update_input_units / update_angles, PsiDeep.hpp:269-280, 311-343
    __device__ void update(const uint64_t* oldc, const uint64_t* newc, cplx* pl, const unsigned char* = nullptr) const {
        const unsigned lane = threadIdx.x & 31u;
        const DeepLayerDev& l0 = L[0];
        for(unsigned w = 0; w < words; w++) {
            uint64_t diff = oldc[w] ^ newc[w];
            while(diff) {
                const unsigned pos = w * 64u + (unsigned)__ffsll((long long)diff) - 1u;
                const double delta = spin_at(newc, pos) - spin_at(oldc, pos);
                for(unsigned j = lane; j < l0.rhs_conn; j += 32u)
                    pl[l0.rhs_c[pos * l0.rhs_conn + j]] += delta * l0.rhs_w[pos * l0.rhs_conn + j];
                __syncwarp();
                diff &= diff - 1ull;
            }
        }
    }
    // foreach_O_k (back-propagation), PsiDeep.hpp:347-445
    __device__ void O_k(const uint64_t* conf, cplx* pl, cplx* row, const unsigned char* blk = nullptr) const {
        const unsigned lane = threadIdx.x & 31u;
        cplx* angles = pl; cplx* act = pl + L[1].size; cplx* tmp = act + width; cplx* deep = tmp + width;
        for(unsigned i = lane; i < N; i += 32u) row[i] = cplx(spin_at(conf, i), 0.0);
        init(conf, pl);
        (void)forward(pl);
        __syncwarp();
        for(int l = (int)num_layers - 1; l > 0; l--) {
            const DeepLayerDev& ly = L[l];
            if(l == (int)num_layers - 1) {
                for(unsigned j = lane; j < ly.size; j += 32u)
                    tmp[j] = final_w[j] * (num_layers == 2u ? act_th(angles[j], 0u) : act_th(deep[ly.begin_deep + j], num_layers - 2u));
            } else {
                for(unsigned i = lane; i < ly.size; i += 32u) {
                    cplx u(0.0, 0.0);
                    for(unsigned j = 0; j < ly.rhs_conn; j++) cfma(u, ly.rhs_w[i * ly.rhs_conn + j], act[ly.rhs_c[i * ly.rhs_conn + j]]);
                    tmp[i] = u * (l == 1 ? act_th(angles[i], 0u) : act_th(deep[ly.begin_deep + i], (unsigned)l - 1u));
                }
            }
            __syncwarp();
            for(unsigned j = lane; j < ly.size; j += 32u) act[j] = tmp[j];
            __syncwarp();
            for(unsigned j = lane; j < ly.size; j += 32u) {
                const cplx d = act[j];
                row[ly.begin_params + j] = d;
                for(unsigned i = 0; i < ly.conn; i++) {
                    const unsigned lhs = ly.lhs_c[i * ly.size + j];
                    const cplx in = (l == 1) ? cplx(spin_at(conf, lhs), 0.0)
                                  : (l == 2) ? act_lc(angles[lhs], 0u)
                                             : act_lc(deep[L[l - 1].begin_deep + lhs], (unsigned)l - 1u);
                    row[ly.begin_params + ly.size + i * ly.size + j] = d * in;
                }
            }
            __syncwarp();
        }
    }
#endif
};

// ------------------------------------------------------------------------------------------ PsiCNN
// Periodic cross-correlation network, <= 3 lattice dimensions, symmetry-class weight sharing, no biases,
// no incremental update (include/quantum_state/PsiCNN.hpp:33-286, detail/Convolve.hpp:108-173).
// out_cj[x] = sum_ci sum_c w[ci,cj][sym(x)*vol + c] * in_ci[nbr(x, c)], nbr = forward periodic shift.
// The neighbour table nbr[l][x*vol + c] and its inverse inv[l][y*vol + c] (the x with nbr(x,c) = y) are
// precomputed on the host, so the kernels do no div/mod.
constexpr int CNN_MAX_LAYERS = 4;
constexpr int CNN_MAX_LINKS = 64;
constexpr int CNN_MAXCH = 8;          // channels per layer (the reference: max_channels_per_layer = 6)
struct CnnLayerDev {
    unsigned nch, prev, vol, angle_off, begin_params, num_params;
    unsigned link_begin[CNN_MAX_LINKS];   // [ci * nch + cj] -> offset into params
    const unsigned* nbr;                  // [N][vol]
    const unsigned* inv;                  // [N][vol]
};
struct CnnDev {
    unsigned N, words, P, num_layers, num_sym, num_angles, maxch;
    bool     keep_angles;      // pre-activations are recorded only for O_k (back-propagation); samplers / E_loc drop them
    double   final_factor;
    cplx     lp;
    const unsigned* sym;       // [N]
    const cplx*     params;    // [P]
    CnnLayerDev     L[CNN_MAX_LAYERS];

    // scratch: in[maxch*N] | out[maxch*N] | angles[num_angles] (the last only when keep_angles: 14.4 of 24 KB per warp
    // at C3, i.e. 2.5x the resident warps for the sampler and E_loc without it)
    __host__ __device__ unsigned payload_elems() const { return 2u * maxch * N + (keep_angles ? num_angles : 0u); }
    // the weights are staged once per block in shared memory (broadcast LDS instead of L1 round trips) when they fit
    __host__ __device__ unsigned block_scratch_bytes() const { return (P <= 1024u) ? P * (unsigned)sizeof(cplx) : 0u; }
#ifdef __CUDACC__
    // copies the weights into the block's shared scratch (all threads of the block call this); returns the staged copy
    __device__ const unsigned char* stage(unsigned char* blk) const {
        if(!block_scratch_bytes()) return nullptr;
        // staged order per layer: [symmetry class][kernel offset c][input channel ci][output channel cj], i.e. the order
        // in which layer_forward consumes them (one running pointer, no index arithmetic in the inner loop)
        cplx* w = reinterpret_cast<cplx*>(blk);
        for(unsigned l = 0; l < num_layers; l++) {
            const CnnLayerDev& ly = L[l];
            for(unsigned e = threadIdx.x; e < ly.num_params; e += blockDim.x) {
                const unsigned cj = e % ly.nch, ci = (e / ly.nch) % ly.prev, c = (e / (ly.nch * ly.prev)) % ly.vol, sc = e / (ly.nch * ly.prev * ly.vol);
                w[ly.begin_params + e] = params[ly.link_begin[ci * ly.nch + cj] + sc * ly.vol + c];
            }
        }
        __syncthreads();
        return blk;
    }
    __device__ void init(const uint64_t*, cplx*, const unsigned char* = nullptr) const {}
    // All NCH outputs of layer l at site x from staged weights (see stage()): the one routine behind the full forward pass
    // and the incremental sampler (cnn_kernels.cuh), so both produce bit-identical activations.  PREV = input channels
    // (compile time, 0 = run-time loop); FIRST = layer 0, whose inputs are the real spins: only the two FMAs with a
    // non-zero factor are issued (the omitted ones add w * 0).  in_real (optional, FIRST only): spins as doubles.
    template<int NCH, int PREV, bool FIRST>
    __device__ __forceinline__ void site_outputs(const CnnLayerDev& ly, unsigned l, const cplx* __restrict__ wgt, const cplx* in,
                                                 const double* in_real, unsigned x, cplx* out, cplx* angles) const {
        const unsigned* nb = ly.nbr + x * ly.vol;
        const unsigned prev = PREV ? (unsigned)PREV : ly.prev;
        const cplx* wq = wgt + ly.begin_params + (size_t)sym[x] * ly.vol * prev * NCH;
        cplx acc[NCH];
        #pragma unroll
        for(int cj = 0; cj < NCH; cj++) acc[cj] = cplx(0.0, 0.0);
        #pragma unroll 3
        for(unsigned c = 0; c < ly.vol; c++) {
            const unsigned src = nb[c];
            if(FIRST) {
                const double sv = in_real ? in_real[src] : in[src].re;
                #pragma unroll
                for(int cj = 0; cj < NCH; cj++) { acc[cj].re = fma(wq[cj].re, sv, acc[cj].re); acc[cj].im = fma(wq[cj].im, sv, acc[cj].im); }
                wq += NCH;
            } else if(PREV) {
                #pragma unroll
                for(int ci = 0; ci < (PREV ? PREV : 1); ci++) {
                    const cplx sv = in[(unsigned)ci * N + src];
                    #pragma unroll
                    for(int cj = 0; cj < NCH; cj++) cfma(acc[cj], wq[cj], sv);
                    wq += NCH;
                }
            } else {
                for(unsigned ci = 0; ci < prev; ci++) {
                    const cplx sv = in[ci * N + src];
                    #pragma unroll
                    for(int cj = 0; cj < NCH; cj++) cfma(acc[cj], wq[cj], sv);
                    wq += NCH;
                }
            }
        }
        #pragma unroll
        for(int cj = 0; cj < NCH; cj++) {
            if(angles) angles[ly.angle_off + (unsigned)cj * N + x] = acc[cj];
            out[(unsigned)cj * N + x] = act_lc(acc[cj], l);
        }
    }
    template<int NCH>
    __device__ __forceinline__ void site_outputs_nch(const CnnLayerDev& ly, unsigned l, const cplx* __restrict__ wgt, const cplx* in,
                                                     const double* in_real, unsigned x, cplx* out, cplx* angles) const {
        if(l == 0u) { site_outputs<NCH, 1, true>(ly, l, wgt, in, in_real, x, out, angles); return; }
        if(NCH <= 4) {
            switch(ly.prev) {
                case 1: site_outputs<NCH, 1, false>(ly, l, wgt, in, in_real, x, out, angles); return;
                case 2: site_outputs<NCH, 2, false>(ly, l, wgt, in, in_real, x, out, angles); return;
                case 3: site_outputs<NCH, 3, false>(ly, l, wgt, in, in_real, x, out, angles); return;
                case 4: site_outputs<NCH, 4, false>(ly, l, wgt, in, in_real, x, out, angles); return;
                default: break;
            }
        }
        site_outputs<NCH, 0, false>(ly, l, wgt, in, in_real, x, out, angles);
    }
    __device__ __forceinline__ void site_outputs_any(const CnnLayerDev& ly, unsigned l, const cplx* __restrict__ wgt, const cplx* in,
                                                     const double* in_real, unsigned x, cplx* out, cplx* angles) const {
        switch(ly.nch) {                 // compile-time channel count: no predicated-off FP64 work
            case 1: site_outputs_nch<1>(ly, l, wgt, in, in_real, x, out, angles); break;
            case 2: site_outputs_nch<2>(ly, l, wgt, in, in_real, x, out, angles); break;
            case 3: site_outputs_nch<3>(ly, l, wgt, in, in_real, x, out, angles); break;
            case 4: site_outputs_nch<4>(ly, l, wgt, in, in_real, x, out, angles); break;
            case 5: site_outputs_nch<5>(ly, l, wgt, in, in_real, x, out, angles); break;
            case 6: site_outputs_nch<6>(ly, l, wgt, in, in_real, x, out, angles); break;
            case 7: site_outputs_nch<7>(ly, l, wgt, in, in_real, x, out, angles); break;
            default: site_outputs_nch<8>(ly, l, wgt, in, in_real, x, out, angles); break;
        }
    }
    // un-staged weights (P > 1024): indexed through link_begin
    template<int NCH>
    __device__ __forceinline__ void layer_forward_unstaged(const CnnLayerDev& ly, unsigned l, const cplx* __restrict__ wgt,
                                                           const cplx* in, cplx* out, cplx* angles) const {
        const unsigned lane = threadIdx.x & 31u;
        for(unsigned x = lane; x < N; x += 32u) {
            const unsigned* nb = ly.nbr + x * ly.vol;
            const unsigned wo = sym[x] * ly.vol;
            cplx acc[NCH];
            #pragma unroll
            for(int cj = 0; cj < NCH; cj++) acc[cj] = cplx(0.0, 0.0);
            for(unsigned c = 0; c < ly.vol; c++) {
                const unsigned src_idx = nb[c];
                for(unsigned ci = 0; ci < ly.prev; ci++) {
                    const cplx sv = in[ci * N + src_idx];
                    const unsigned* lb = ly.link_begin + ci * NCH;
                    #pragma unroll
                    for(int cj = 0; cj < NCH; cj++) cfma(acc[cj], wgt[lb[cj] + wo + c], sv);
                }
            }
            #pragma unroll
            for(int cj = 0; cj < NCH; cj++) {
                if(angles) angles[ly.angle_off + (unsigned)cj * N + x] = acc[cj];
                out[(unsigned)cj * N + x] = act_lc(acc[cj], l);
            }
        }
    }
    // forward_pass, PsiCNN.hpp:99-160 (angles recorded into the warp's scratch when keep_angles).  One lane per lattice site x: the
    // vol neighbour indices are loaded once, every input value is read once from shared memory and used for ALL output
    // channels (register accumulators), so the inner loop is FP64-bound instead of load-bound.
    __device__ cplx forward(const uint64_t* conf, cplx* pl, const unsigned char* blk) const {
        const unsigned lane = threadIdx.x & 31u;
        const cplx* __restrict__ wgt = blk ? reinterpret_cast<const cplx*>(blk) : params;
        cplx* in = pl; cplx* out = pl + maxch * N; cplx* angles = keep_angles ? out + maxch * N : nullptr;
        for(unsigned j = lane; j < N; j += 32u) in[j] = cplx(spin_at(conf, j), 0.0);
        __syncwarp();
        cplx result(0.0, 0.0);
        for(unsigned l = 0; l < num_layers; l++) {
            const CnnLayerDev& ly = L[l];
            if(blk) {
                for(unsigned x = lane; x < N; x += 32u) site_outputs_any(ly, l, wgt, in, nullptr, x, out, angles);
            } else {
                switch(ly.nch) {
                    case 1: layer_forward_unstaged<1>(ly, l, wgt, in, out, angles); break;
                    case 2: layer_forward_unstaged<2>(ly, l, wgt, in, out, angles); break;
                    case 3: layer_forward_unstaged<3>(ly, l, wgt, in, out, angles); break;
                    case 4: layer_forward_unstaged<4>(ly, l, wgt, in, out, angles); break;
                    case 5: layer_forward_unstaged<5>(ly, l, wgt, in, out, angles); break;
                    case 6: layer_forward_unstaged<6>(ly, l, wgt, in, out, angles); break;
                    case 7: layer_forward_unstaged<7>(ly, l, wgt, in, out, angles); break;
                    default: layer_forward_unstaged<8>(ly, l, wgt, in, out, angles); break;
                }
            }
            __syncwarp();
            if(l + 1u < num_layers) {
                for(unsigned idx = lane; idx < ly.nch * N; idx += 32u) in[idx] = out[idx];
                __syncwarp();
            } else {
                for(unsigned idx = lane; idx < ly.nch * N; idx += 32u) result += out[idx];
            }
        }
        return final_factor * warp_sum(result);
    }
    __device__ cplx log_psi(const uint64_t* conf, cplx* pl, const unsigned char* blk = nullptr) const { return lp + forward(conf, pl, blk); }
    __device__ void update(const uint64_t*, const uint64_t*, cplx*, const unsigned char* = nullptr) const {}   // PsiCNN.hpp:177-181
    // foreach_O_k, PsiCNN.hpp:185-266, in gather form: each parameter is produced once (the reference emits
    // a k several times and its consumers accumulate atomically, Appendix A.9).
    __device__ void O_k(const uint64_t* conf, cplx* pl, cplx* row, const unsigned char* blk = nullptr) const {
        const unsigned lane = threadIdx.x & 31u;
        cplx* delta = pl; cplx* back = pl + maxch * N; cplx* angles = back + maxch * N;
        (void)forward(conf, pl, blk);
        __syncwarp();
        const unsigned lastc = L[num_layers - 1u].nch;
        for(unsigned idx = lane; idx < lastc * N; idx += 32u) back[idx] = cplx(final_factor, 0.0);
        __syncwarp();
        for(int l = (int)num_layers - 1; l >= 0; l--) {
            const CnnLayerDev& ly = L[l];
            // delta at this layer's pre-activations
            for(unsigned idx = lane; idx < ly.nch * N; idx += 32u) delta[idx] = back[idx] * act_th(angles[ly.angle_off + idx], (unsigned)l);
            __syncwarp();
            // parameter gradients: one lane per (link, symmetry class, kernel offset)
            const unsigned per_link = num_sym * ly.vol;
            for(unsigned k = lane; k < ly.num_params; k += 32u) {
                const unsigned link = k / per_link, r = k - link * per_link;
                const unsigned s = r / ly.vol, c = r - s * ly.vol;
                const unsigned ci = link / ly.nch, cj = link - ci * ly.nch;
                cplx g(0.0, 0.0);
                for(unsigned x = 0; x < N; x++) {
                    if(sym[x] != s) continue;
                    const unsigned src = ly.nbr[x * ly.vol + c];
                    const cplx in = (l == 0) ? cplx(spin_at(conf, src), 0.0)
                                             : act_lc(angles[L[l - 1].angle_off + ci * N + src], (unsigned)l - 1u);
                    cfma(g, delta[cj * N + x], in);
                }
                row[ly.link_begin[link] + r] = g;
            }
            // back-propagate to the previous layer's activations
            if(l > 0) {
                __syncwarp();
                for(unsigned idx = lane; idx < ly.prev * N; idx += 32u) {
                    const unsigned ci = idx / N, y = idx - ci * N;
                    cplx acc(0.0, 0.0);
                    for(unsigned cj = 0; cj < ly.nch; cj++) {
                        const cplx* w = params + ly.link_begin[ci * ly.nch + cj];
                        for(unsigned c = 0; c < ly.vol; c++) {
                            const unsigned x = ly.inv[y * ly.vol + c];
                            cfma(acc, w[sym[x] * ly.vol + c], delta[cj * N + x]);
                        }
                    }
                    back[idx] = acc;
                }
            }
            __syncwarp();
        }
    }
#endif
};

// ------------------------------------------------------------------------------------------ PsiClassical
// log psi = log_prefactor + sum_n params[n] * fast_local_energy(H_local[n], s) (+ log psi_ref(s) for order 2);
// O_k[n] = that energy, followed by psi_ref's O_k for order 2 (include/quantum_state/PsiClassical.hpp:48-160).
// psi_ref is PsiFullyPolarized (log psi_ref = 0, PsiFullyPolarized.hpp:41-49) or a PsiCNN.
struct ClassicalDev {
    unsigned N, words, P, num_ops, order;
    cplx     lp;
    const OpDev* ops;      // [num_ops] (device array of views)
    const cplx*  params;   // [num_ops]
    bool     has_ref;
    CnnDev   ref;

    __host__ __device__ unsigned payload_elems() const { return (order > 1u && has_ref) ? ref.payload_elems() : 1u; }
    __host__ __device__ unsigned block_scratch_bytes() const { return 0u; }
#ifdef __CUDACC__
    __device__ const unsigned char* stage(unsigned char*) const { return nullptr; }
    __device__ void init(const uint64_t*, cplx*, const unsigned char* = nullptr) const {}
    __device__ cplx log_psi(const uint64_t* conf, cplx* pl, const unsigned char* = nullptr) const {
        const unsigned lane = threadIdx.x & 31u;
        cplx acc(0.0, 0.0);
        for(unsigned n = lane; n < num_ops; n += 32u) acc += params[n] * fast_local_energy_serial(ops[n], conf);
        cplx r = lp + warp_sum(acc);
        if(order > 1u && has_ref) r += ref.log_psi(conf, pl);
        return r;
    }
    __device__ void update(const uint64_t*, const uint64_t*, cplx*, const unsigned char* = nullptr) const {}
    __device__ void O_k(const uint64_t* conf, cplx* pl, cplx* row, const unsigned char* blk = nullptr) const {
        const unsigned lane = threadIdx.x & 31u;
        for(unsigned n = lane; n < num_ops; n += 32u) row[n] = fast_local_energy_serial(ops[n], conf);
        if(order > 1u && has_ref) ref.O_k(conf, pl, row + num_ops);
    }
#endif
};

} // namespace angpu

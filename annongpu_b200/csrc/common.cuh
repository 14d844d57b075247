// Shared device/host helpers for the sm_100a VMC kernels.
//
// Arithmetic is complex fp64 throughout, as in the reference (include/types.h:25).  `cplx` is a
// plain (re, im) pair with the unscaled multiply the reference's cuda_complex.hpp uses.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace angpu {

constexpr int MAXW = 4;            // 64-bit words per configuration: N <= 256
constexpr unsigned FULL = 0xffffffffu;

struct __align__(16) cplx {
    double re, im;
    __host__ __device__ cplx() {}
    __host__ __device__ constexpr cplx(double r, double i = 0.0) : re(r), im(i) {}
};

__host__ __device__ __forceinline__ cplx operator+(cplx a, cplx b) { return cplx(a.re + b.re, a.im + b.im); }
__host__ __device__ __forceinline__ cplx operator-(cplx a, cplx b) { return cplx(a.re - b.re, a.im - b.im); }
__host__ __device__ __forceinline__ cplx operator-(cplx a) { return cplx(-a.re, -a.im); }
__host__ __device__ __forceinline__ cplx operator*(cplx a, cplx b) { return cplx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__host__ __device__ __forceinline__ cplx operator*(double a, cplx b) { return cplx(a * b.re, a * b.im); }
__host__ __device__ __forceinline__ cplx operator*(cplx b, double a) { return cplx(a * b.re, a * b.im); }
__host__ __device__ __forceinline__ cplx& operator+=(cplx& a, cplx b) { a.re += b.re; a.im += b.im; return a; }
__host__ __device__ __forceinline__ cplx& operator-=(cplx& a, cplx b) { a.re -= b.re; a.im -= b.im; return a; }
__host__ __device__ __forceinline__ cplx conj(cplx a) { return cplx(a.re, -a.im); }
__host__ __device__ __forceinline__ double abs2(cplx a) { return a.re * a.re + a.im * a.im; }
// a += b * c
__host__ __device__ __forceinline__ void cfma(cplx& a, cplx b, cplx c) {
    a.re = fma(b.re, c.re, fma(-b.im, c.im, a.re));
    a.im = fma(b.re, c.im, fma(b.im, c.re, a.im));
}
__device__ __forceinline__ cplx cexp(cplx z) {
    double s, c;
    sincos(z.im, &s, &c);
    const double e = exp(z.re);
    return cplx(e * c, e * s);
}

// Activation polynomials of the reference (include/quantum_state/psi_functions.hpp:43-48, 106-115):
//   lc(z, 0) = z^2/2 - z^4/12 + z^6/45        lc(z, >0) = th(z, 0) = z - z^3/3 + 2 z^5/15
//   th(z, >0) = 1 - z^2 + 2 z^4/3
__host__ __device__ __forceinline__ cplx act_lc(cplx z, unsigned layer) {
    const cplx z2 = z * z, z4 = z2 * z2;
    if(layer == 0u) return 0.5 * z2 - (1.0 / 12.0) * z4 + (1.0 / 45.0) * (z4 * z2);
    return z - (1.0 / 3.0) * (z2 * z) + (2.0 / 15.0) * (z4 * z);
}
__host__ __device__ __forceinline__ cplx act_th(cplx z, unsigned layer) {
    const cplx z2 = z * z, z4 = z2 * z2;
    if(layer == 0u) return z - (1.0 / 3.0) * (z2 * z) + (2.0 / 15.0) * (z4 * z);
    return cplx(1.0, 0.0) - z2 + (2.0 / 3.0) * z4;
}

// spin value of site i: bit set <=> +1 (include/basis/Spins.h:104-108)
__host__ __device__ __forceinline__ double spin_at(const uint64_t* conf, unsigned i) {
    return ((conf[i >> 6] >> (i & 63u)) & 1ull) ? 1.0 : -1.0;
}

#ifdef __CUDACC__
// read-only 16-byte load of a complex number
__device__ __forceinline__ cplx ldg(const cplx* p) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return cplx(v.x, v.y);
}
__device__ __forceinline__ double warp_sum(double v) {
    #pragma unroll
    for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ cplx warp_sum(cplx v) {
    #pragma unroll
    for(int o = 16; o > 0; o >>= 1) {
        v.re += __shfl_xor_sync(FULL, v.re, o);
        v.im += __shfl_xor_sync(FULL, v.im, o);
    }
    return v;
}

// Philox4x32-10 (Salmon et al., SC'11).  Counter layout shared with the CPU oracle
// (oracle/port/vmc_port.c): ctr = (step_lo, step_hi, chain, (call << 1) | tag), key = (seed_lo, seed_hi);
// tag 0 = initial configuration (step = word index), tag 1 = Metropolis proposals.
// Replaces the reference's per-chain XORWOW state (source/RNGStates.cu:13-19): no state to load/store,
// streams are reproducible for any chain->GPU assignment.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1, uint32_t out[4]) {
    #pragma unroll
    for(int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// uniform in (0, 1] from 53 random bits (the reference's GPU path draws curand_uniform in (0,1], include/random.h:47-54)
__host__ __device__ __forceinline__ double u01_from_bits(uint32_t lo, uint32_t hi) {
    return (double)((((uint64_t)lo | ((uint64_t)hi << 32)) >> 11) + 1ull) * 0x1.0p-53;
}
#endif

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

#define ANGPU_CUDA(cmd) do { cudaError_t e_ = (cmd); if(e_ != cudaSuccess) \
    throw ::angpu::Error(std::string(#cmd) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); } while(0)
#define ANGPU_CHECK_LAUNCH() ANGPU_CUDA(cudaGetLastError())
#define ANGPU_REQUIRE(cond, msg) do { if(!(cond)) throw ::angpu::Error(std::string(msg) + " [" #cond "]"); } while(0)

inline unsigned words_for(unsigned n) { return (n + 63u) / 64u; }

} // namespace angpu
